"""o3d.geometry.PointCloud subset, backed by the CUDA library."""
from __future__ import annotations

import numpy as np

from .. import api
from .utility import Vector3dVector


class _Distances(np.ndarray):
    """(N,) float64 ndarray subclass so that np.asarray(result) is free."""


#: clouds of at least this many points are transformed on the device (isr_transform_points_f64)
DEVICE_TRANSFORM_MIN_POINTS = 50_000


class PointCloud:
    """Coordinates are float64, like Open3D's.  The cloud has a host copy (numpy, what `.points`
    shows) and, once a device call has needed it, a float64 device copy; each is refreshed from
    the other only when it is stale, so a transform followed by a distance or ICP call on a
    large cloud never round-trips through the host."""

    def __init__(self, points=None):
        self._points = Vector3dVector(points) if points is not None else Vector3dVector()
        self._dev = None          # torch float64 [N,3] on the device, or None
        self._host_stale = False  # the device copy is newer than self._points
        self.colors = Vector3dVector()

    # -- points property: accepts Vector3dVector or any N x 3 array ----------------------
    @property
    def points(self):
        self._sync_host()
        return self._points

    @points.setter
    def points(self, value):
        self._points = value if isinstance(value, Vector3dVector) else Vector3dVector(value)
        self._dev, self._host_stale = None, False

    def _sync_host(self):
        if self._host_stale:
            self._points = Vector3dVector(self._dev.cpu().numpy())
            self._host_stale = False

    def _np(self) -> np.ndarray:
        self._sync_host()
        return np.asarray(self._points)

    def _device_points(self):
        """float64 [N,3] on the current CUDA device (uploaded once)."""
        import torch

        if self._dev is None:
            self._dev = torch.from_numpy(np.ascontiguousarray(self._np(), dtype=np.float64)).to(api._device())
        return self._dev

    def _for_kernel(self):
        """What a distance / ICP call should be handed: the device copy if there is one."""
        return self._dev if self._dev is not None else self._np()

    def __len__(self):
        return len(self._points) if not self._host_stale else int(self._dev.shape[0])

    def has_points(self) -> bool:
        return len(self) > 0

    # -- icp.py:22,110 --------------------------------------------------------------
    def transform(self, transformation):
        """In-place p <- T[:3,:3] p + T[:3,3] in float64; returns self (like Open3D).  Large clouds
        (and clouds that already live on the device) are transformed there
        (isr_transform_points_f64: K1 with float64 output); small host clouds with numpy."""
        T = np.asarray(transformation, dtype=np.float64)
        if T.shape != (4, 4):
            raise RuntimeError("transform expects a 4x4 matrix")
        import torch

        if (self._dev is not None or len(self) >= DEVICE_TRANSFORM_MIN_POINTS) and torch.cuda.is_available():
            d = self._device_points()
            api.transform_points_f64(d, T, out=d)
            self._host_stale = True
        else:
            self._points = Vector3dVector(self._np() @ T[:3, :3].T + T[:3, 3])
        return self

    def paint_uniform_color(self, color):
        self.colors = Vector3dVector(np.tile(np.asarray(color, dtype=np.float64), (len(self), 1)))
        return self

    def __add__(self, other):
        """Concatenate points (icp.py:111)."""
        out = PointCloud()
        if self._dev is not None or other._dev is not None:   # stay on the device
            import torch

            out._dev = torch.cat([self._device_points(), other._device_points()], dim=0)
            out._host_stale = True
            return out
        out._points = Vector3dVector(np.concatenate([self._np(), other._np()], axis=0))
        return out

    def __iadd__(self, other):
        self._points = Vector3dVector(np.concatenate([self._np(), other._np()], axis=0))
        self._dev, self._host_stale = None, False
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
        d = api.point_cloud_distance(self._for_kernel(), target._for_kernel())
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
