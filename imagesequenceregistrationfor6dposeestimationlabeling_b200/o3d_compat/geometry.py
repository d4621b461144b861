"""o3d.geometry.PointCloud subset, backed by the CUDA library."""
from __future__ import annotations

import numpy as np

from .. import api
from .utility import Vector3dVector


class _Distances(np.ndarray):
    """(N,) float64 ndarray subclass so that np.asarray(result) is free."""


class PointCloud:
    def __init__(self, points=None):
        self._points = Vector3dVector(points) if points is not None else Vector3dVector()
        self.colors = Vector3dVector()

    # -- points property: accepts Vector3dVector or any N x 3 array ----------------------
    @property
    def points(self):
        return self._points

    @points.setter
    def points(self, value):
        self._points = value if isinstance(value, Vector3dVector) else Vector3dVector(value)

    def _np(self) -> np.ndarray:
        return np.asarray(self._points)

    def __len__(self):
        return len(self._points)

    def has_points(self) -> bool:
        return len(self._points) > 0

    # -- icp.py:22,110 --------------------------------------------------------------
    def transform(self, transformation):
        """In-place p <- T[:3,:3] p + T[:3,3]; returns self (like Open3D).

        Kept in float64 on the host copy so repeated transforms do not accumulate FP32
        rounding; the device sees the cloud only when a distance / ICP call needs it."""
        T = np.asarray(transformation, dtype=np.float64)
        if T.shape != (4, 4):
            raise RuntimeError("transform expects a 4x4 matrix")
        self._points = Vector3dVector(self._np() @ T[:3, :3].T + T[:3, 3])
        return self

    def paint_uniform_color(self, color):
        self.colors = Vector3dVector(np.tile(np.asarray(color, dtype=np.float64), (len(self), 1)))
        return self

    def __add__(self, other):
        """Concatenate points (icp.py:111)."""
        out = PointCloud()
        out._points = Vector3dVector(np.concatenate([self._np(), other._np()], axis=0))
        return out

    def __iadd__(self, other):
        self._points = Vector3dVector(np.concatenate([self._np(), other._np()], axis=0))
        return self

    def __copy__(self):
        return PointCloud(self._np().copy())

    def __deepcopy__(self, memo):
        c = PointCloud(self._np().copy())
        c.colors = Vector3dVector(np.asarray(self.colors).copy())
        return c

    # -- verfication.py:97,99; icp.py:113,115 -------------------------------------------
    def compute_point_cloud_distance(self, target):
        """Distance of every point to its nearest neighbour in `target` (K2 on the GPU)."""
        if len(self) == 0:
            return np.zeros(0).view(_Distances)
        if len(target) == 0:
            return np.zeros(len(self)).view(_Distances)
        d = api.point_cloud_distance(self._np(), target._np())
        return d.cpu().numpy().view(_Distances)

    # -- generateCors.py:254-258, trainPose.py:343-347 ------------------------------------
    def remove_radius_outlier(self, nb_points, radius, print_progress=False):
        """(filtered cloud, list of kept indices): a point stays iff more than `nb_points`
        points (itself included) lie within `radius` of it (strict, float64 decision)."""
        if len(self) == 0:
            return PointCloud(), []
        cnt = api.radius_neighbor_count(self._np(), float(radius)).cpu().numpy()
        ind = np.nonzero(cnt > int(nb_points))[0]
        return self.select_by_index(ind), [int(i) for i in ind]

    def select_by_index(self, indices, invert=False):
        idx = np.asarray(indices, dtype=np.int64)
        if invert:
            keep = np.ones(len(self), dtype=bool)
            keep[idx] = False
            idx = np.nonzero(keep)[0]
        out = PointCloud()
        out._points = Vector3dVector(self._np()[idx])
        return out

    def __repr__(self):
        return f"PointCloud with {len(self)} points."
