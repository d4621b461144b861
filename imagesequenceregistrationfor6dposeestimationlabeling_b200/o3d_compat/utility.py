"""o3d.utility subset."""
from __future__ import annotations

import numpy as np


class Vector3dVector:
    """N x 3 float64 host array with the sequence protocol np.asarray() expects."""

    def __init__(self, data=None):
        a = np.zeros((0, 3)) if data is None else np.array(data, dtype=np.float64)
        if a.ndim != 2 or a.shape[1] != 3:
            raise RuntimeError("Vector3dVector expects an N x 3 array")
        self._a = np.ascontiguousarray(a)

    def __array__(self, dtype=None, copy=None):
        return self._a if dtype is None else self._a.astype(dtype)

    def __len__(self):
        return len(self._a)

    def __getitem__(self, i):
        return self._a[i]


class DoubleVector(list):
    """What compute_point_cloud_distance returns; np.asarray() gives (N,) float64."""

    def __array__(self, dtype=None, copy=None):
        return np.array(list(self), dtype=np.float64 if dtype is None else dtype)
