"""scikit-learn-shaped shim: ``from <pkg>.compat import KDTree`` (choosePose.py:6,21-22).

Only the subset the reference uses: ``KDTree(X, leaf_size=2).query(Q, k=1)`` ->
``(dist (nq,1) float64, idx (nq,1) int64)``.  There is no tree: the "build" repacks X into
the SoA planes once and every query is a brute-force K2 launch.
"""
from __future__ import annotations

import numpy as np

from . import api


class KDTree:
    def __init__(self, X, leaf_size=40, metric="minkowski", **kwargs):
        if metric not in ("minkowski", "euclidean", "l2"):
            raise ValueError("only the Euclidean metric is supported")
        self._dev = api._device()
        X = np.asarray(X) if not hasattr(X, "device") else X
        self._pts = api._points(X, self._dev)
        if self._pts.dim() != 2:
            raise ValueError("X must be [N, 3]")
        self.data = self._pts
        self._cen = api.centroid_of(self._pts, self._dev)
        self._soa = api.prepare_cloud(self._pts, centroid=self._cen,
                                      perm=api.spatial_order(self._pts, self._dev), stage_centroids=True,
                                      device=self._dev)

    def query(self, X, k=1, return_distance=True):
        if k != 1:
            raise NotImplementedError("only k=1 is used by the reference (choosePose.py:22)")
        qp = api._points(X, self._dev)
        q = api.prepare_cloud(qp, centroid=self._cen, perm=api.spatial_order(qp, self._dev),
                              device=self._dev)
        res = api.nearest_neighbors_soa(q, self._soa, return_index=True)
        idx = res.idx[0].to("cpu").numpy().astype(np.int64).reshape(-1, 1)
        if not return_distance:
            return idx
        dist = res.dist[0].to("cpu").numpy().reshape(-1, 1)
        return dist, idx
