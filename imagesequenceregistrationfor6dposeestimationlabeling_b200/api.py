"""Batched tensor-in / tensor-out API over libisr (hand-written sm_100a kernels).

Every function takes numpy arrays or torch tensors (host or device), moves them to the
current CUDA device if needed, and launches on torch's current stream.  Nothing here
computes on the CPU: without a GPU and csrc/libisr.so these functions raise.

Reference mapping (file:line into the reference repository):
  transform_points    pc.dot(R.T)+t            verfication.py:83-85, icp.py:68; PointCloud.transform icp.py:110
  nearest_neighbors   KD-tree 1-NN             verfication.py:97,99; icp.py:97-103; choosePose.py:21-22
  chamfer_distance    Chamfer                  verfication.py:97-101; icp.py:113-117
  adds                ADDS                     choosePose.py:20-22
  verify_poses        candidate loop + argmin  verfication.py:61-108; choosePose.py:124-138
  evaluate_registration / icp                  icp.py:96-103
  multistart_icp      symmetry-seeded starts   README.md:42-46 (config 5 of BASELINE.json)
  icp_refine_pose     (R, t, loss) convention  pose_refine.py:21-22,101-104 (alias: refine_pose)
"""
from __future__ import annotations

import ctypes
import dataclasses
import functools
import inspect
from typing import Optional, Sequence

import numpy as np
import torch

from . import _lib


# --------------------------------------------------------------------------------------
# plumbing
# --------------------------------------------------------------------------------------
def _device(device=None) -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError("CUDA device required: this package has no CPU fallback")
    if device is None:
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device(device)


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _on_device(fn):
    """libisr launches on the device that is CURRENT in the calling thread, on torch's current
    stream of that device.  A public function that is handed `device=` (or tensors living on
    another GPU) therefore runs its whole body under ``torch.cuda.device(...)``: allocations,
    stream and kernels all belong to the same GPU."""
    sig = inspect.signature(fn)

    @functools.wraps(fn)
    def wrapper(*args, **kwargs):
        dev = sig.bind(*args, **kwargs).arguments.get("device")
        if dev is None:
            for a in list(args) + list(kwargs.values()):
                t = a.data if isinstance(a, SoaCloud) else a
                if isinstance(t, torch.Tensor) and t.is_cuda:
                    dev = t.device
                    break
        if dev is None or not torch.cuda.is_available():
            return fn(*args, **kwargs)
        dev = torch.device(dev)
        if dev.type != "cuda":
            raise ValueError(f"device must be a CUDA device, got {dev}")
        with torch.cuda.device(dev):
            return fn(*args, **kwargs)

    return wrapper


_NP_OF = {torch.float32: np.float32, torch.float64: np.float64, torch.uint8: np.uint8,
          torch.int32: np.int32, torch.int64: np.int64}


def _to_dev(x, dtype, device) -> torch.Tensor:
    if isinstance(x, torch.Tensor):
        t = x
    else:
        t = torch.from_numpy(np.ascontiguousarray(np.asarray(x), dtype=_NP_OF[dtype]))
    if t.device != device:
        # convert on the side that avoids an extra host pass: dtype first for host tensors
        if t.dtype != dtype:
            t = t.to(dtype)
        t = t.to(device, non_blocking=True)
    elif t.dtype != dtype:
        t = t.to(dtype)
    return t.contiguous()


def _points(x, device) -> torch.Tensor:
    t = _to_dev(x, torch.float32, device)
    if t.dim() < 2 or t.shape[-1] != 3:
        raise ValueError(f"points must have shape [..., N, 3], got {tuple(t.shape)}")
    return t


def _points_hilo(x, device):
    """float64 input -> (float32 hi, float32 lo) with hi + lo == x to ~2^-48; float32 input ->
    (x, None).  The library's point type is float32; the lo part lets float64 callers
    (icp.py:68 builds a float64 source) lose nothing."""
    is64 = (isinstance(x, torch.Tensor) and x.dtype == torch.float64) or (
        not isinstance(x, torch.Tensor) and np.asarray(x).dtype == np.float64)
    if not is64:
        return _points(x, device), None
    t = x if isinstance(x, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(x))
    hi = t.to(torch.float32)
    lo = (t - hi.to(torch.float64)).to(torch.float32)
    return _points(hi, device), _points(lo, device)


def _poses(x, device) -> torch.Tensor:
    t = _to_dev(x, torch.float64, device)
    if t.dim() == 2:
        t = t[None]
    if t.dim() != 3 or t.shape[-2:] != (4, 4):
        raise ValueError(f"poses must have shape [B, 4, 4], got {tuple(t.shape)}")
    return t.contiguous()


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _workspace(nbytes: int, device) -> torch.Tensor:
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)


def set_nn_pruning(on: bool) -> None:
    """Process-wide switch of the nearest-neighbour kernel's tile pruning (default on).
    Off = exhaustive brute force over every (query, target) pair; results are identical."""
    _lib.check(_lib.load().isr_set_nn_pruning(1 if on else 0))


def get_nn_pruning() -> bool:
    return bool(_lib.load().isr_get_nn_pruning())


def pose_from_Rt(R, t) -> np.ndarray:
    """4x4 float64 pose (column-vector convention) from a 3x3 rotation and a translation."""
    T = np.eye(4)
    T[:3, :3] = np.asarray(R, dtype=np.float64).reshape(3, 3)
    T[:3, 3] = np.asarray(t, dtype=np.float64).reshape(3)
    return T


# --------------------------------------------------------------------------------------
# K1 -- transform
# --------------------------------------------------------------------------------------
@_on_device
def transform_points(points, poses, device=None) -> torch.Tensor:
    """out[b] = points @ poses[b,:3,:3].T + poses[b,:3,3]  ->  float32 [B, N, 3] (FP64 math)."""
    device = _device(device)
    pts = _points(points, device)
    if pts.dim() != 2:
        raise ValueError("transform_points expects points of shape [N, 3]")
    P = _poses(poses, device)
    n, b = pts.shape[0], P.shape[0]
    out = torch.empty((b, n, 3), dtype=torch.float32, device=device)
    lib = _lib.load()
    # grid.y carries the batch: chunk at 65535
    for b0 in range(0, b, 65535):
        bc = min(65535, b - b0)
        _lib.check(lib.isr_transform_points(_ptr(pts), n, _ptr(P[b0:]), bc, _ptr(out[b0:]), _stream()))
    return out


@_on_device
def transform_points_f64(points: torch.Tensor, pose, out: Optional[torch.Tensor] = None, device=None) -> torch.Tensor:
    """float64 [N,3] device tensor -> pose . points in float64 (in place when out is points):
    Open3D's PointCloud.transform (icp.py:22,110) for a cloud that lives on the device."""
    device = _device(device)
    pts = _to_dev(points, torch.float64, device)
    if pts.dim() != 2 or pts.shape[1] != 3:
        raise ValueError("transform_points_f64 expects points of shape [N, 3]")
    P = _poses(pose, device)
    out = torch.empty_like(pts) if out is None else out
    _lib.check(_lib.load().isr_transform_points_f64(_ptr(pts), pts.shape[0], _ptr(P), _ptr(out), _stream()))
    return out


@dataclasses.dataclass
class SoaCloud:
    """A cloud (or batch of clouds) in the padded plane layout K2 streams.

    3 planes (x, y, z): the direct-difference kernel.  7 planes (hi xyz, |hi|^2, lo xyz,
    centred): the production filtered-exact kernel; `centroid` is the centre it was
    prepared with (both clouds of a search must share it)."""
    data: torch.Tensor
    n: int
    centroid: Optional[torch.Tensor] = None
    perm: Optional[torch.Tensor] = None      # int32 [n]: stored position -> original index
    stage_c: Optional[torch.Tensor] = None   # float32 [B, npad/1024 + chunks, 4] stage, then chunk spheres (c, r)
    sub_c: Optional[torch.Tensor] = None     # float32 [B, npad/64, 4] sub-tile spheres
    sub_box: Optional[torch.Tensor] = None   # int32 [B, npad/64] packed half-extents of the sub-tile boxes

    @property
    def planes(self) -> int:
        return self.data.shape[1]

    def descriptor(self, batched: bool) -> "_lib.IsrCloud":
        """ctypes ``IsrCloud`` for this cloud (7-plane clouds only)."""
        return _lib.IsrCloud(self.data.data_ptr(), self.n, self.npad,
                             7 * self.npad if batched else 0, _ptr(self.stage_c), _ptr(self.perm),
                             _ptr(self.sub_c), None, _ptr(self.sub_box))

    @property
    def npad(self) -> int:
        return self.data.shape[-1]

    @property
    def batch(self) -> int:
        return self.data.shape[0]


@_on_device
def pack_soa(points, poses=None, device=None) -> SoaCloud:
    """Repack [N,3] (optionally transformed by each of poses [B,4,4]) into SoA planes."""
    device = _device(device)
    pts = _points(points, device)
    if pts.dim() != 2:
        raise ValueError("pack_soa expects points of shape [N, 3]")
    n = pts.shape[0]
    npad = _lib.soa_padded_len(n)
    lib = _lib.load()
    if poses is None:
        out = torch.empty((1, 3, npad), dtype=torch.float32, device=device)
        _lib.check(lib.isr_transform_points_soa(_ptr(pts), n, None, 16, 1, _ptr(out), npad, None, 0,
                                                _stream()))
        return SoaCloud(out, n)
    P = _poses(poses, device)
    b = P.shape[0]
    out = torch.empty((b, 3, npad), dtype=torch.float32, device=device)
    for b0 in range(0, b, 65535):
        bc = min(65535, b - b0)
        _lib.check(lib.isr_transform_points_soa(_ptr(pts), n, _ptr(P[b0:]), 16, bc, _ptr(out[b0:]),
                                                npad, None, 0, _stream()))
    return SoaCloud(out, n)


@_on_device
def centroid_of(points, device=None) -> torch.Tensor:
    """FP64 centroid [3] of an [N,3] cloud, on the device (deterministic order)."""
    device = _device(device)
    pts = _points(points, device)
    out = torch.zeros((3,), dtype=torch.float64, device=device)
    if pts.shape[0] > 0:
        _lib.check(_lib.load().isr_centroid(_ptr(pts), pts.shape[0], _ptr(out), _stream()))
    return out


@_on_device
def spatial_order(points, device=None) -> torch.Tensor:
    """int32 [N] Hilbert-curve order of an [N,3] cloud (perm[i] = original index of the i-th stored point)."""
    device = _device(device)
    pts = _points(points, device)
    n = pts.shape[0]
    perm = torch.empty((n,), dtype=torch.int32, device=device)
    if n > 0:
        lib = _lib.load()
        ws = _workspace(lib.isr_spatial_order_workspace_bytes(n), device)
        _lib.check(lib.isr_spatial_order(_ptr(pts), n, _ptr(perm), _ptr(ws), ws.numel(), _stream()))
    return perm


@_on_device
def prepare_cloud(points, poses=None, centroid=None, centre_poses=None, perm=None,
                  stage_centroids: bool = False, device=None) -> SoaCloud:
    """[N,3] (optionally transformed by poses [B,4,4]) -> centred hi/lo SoA7 planes [B,7,npad].
    The centre of batch item b is centre_poses[b] . centroid (centre_poses None: centroid).
    float64 input keeps its precision (hi/lo split).  `perm` (from spatial_order) stores the
    points in that (Hilbert-curve) order; `stage_centroids` adds the bounding spheres of the 1024-point
    stages and 64-point sub-tiles that a target needs for nearest-stage-first scanning and
    tile pruning."""
    device = _device(device)
    pts, pts_lo = _points_hilo(points, device)
    if pts.dim() != 2:
        raise ValueError("prepare_cloud expects points of shape [N, 3]")
    n = pts.shape[0]
    npad = _lib.soa_padded_len(n)
    lib = _lib.load()
    P = None if poses is None else _poses(poses, device)
    C = None if centre_poses is None else _poses(centre_poses, device)
    b = 1 if P is None else P.shape[0]
    if C is not None and C.shape[0] != b:
        raise ValueError("centre_poses must match poses in batch size")
    cen = None if centroid is None else _to_dev(centroid, torch.float64, device)
    out = torch.empty((b, 7, npad), dtype=torch.float32, device=device)
    for b0 in range(0, b, 65535):
        bc = min(65535, b - b0)
        _lib.check(lib.isr_prepare_cloud(
            _ptr(pts), _ptr(pts_lo), _ptr(perm), n, None if P is None else _ptr(P[b0:]), 16,
            None if C is None else _ptr(C[b0:]), 16, _ptr(cen), bc, _ptr(out[b0:]), npad, None, 0,
            _stream()))
    sc = sub = box = None
    if stage_centroids:
        sc = torch.empty((b, lib.isr_stage_sphere_count(npad), 4), dtype=torch.float32, device=device)
        sub = torch.empty((b, npad // _lib.ISR_SUB_TILE, 4), dtype=torch.float32, device=device)
        box = torch.empty((b, npad // _lib.ISR_SUB_TILE), dtype=torch.int32, device=device)
        for b0 in range(0, b, 65535):
            bc = min(65535, b - b0)
            _lib.check(lib.isr_tile_spheres(_ptr(out[b0:]), n, npad, 7 * npad, bc, _ptr(sc[b0:]),
                                            _ptr(sub[b0:]), _ptr(box[b0:]), _stream()))
    return SoaCloud(out, n, cen, perm, sc, sub, box)


def _pack_batched(points, device) -> SoaCloud:
    """[N,3] or [B,N,3] -> SoaCloud with batch 1 or B (no transform)."""
    pts = _points(points, device)
    if pts.dim() == 2:
        return pack_soa(pts, device=device)
    if pts.dim() != 3:
        raise ValueError("points must be [N,3] or [B,N,3]")
    b, n = pts.shape[0], pts.shape[1]
    npad = _lib.soa_padded_len(n)
    out = torch.empty((b, 3, npad), dtype=torch.float32, device=device)
    lib = _lib.load()
    for k in range(b):
        _lib.check(lib.isr_transform_points_soa(_ptr(pts[k]), n, None, 16, 1, _ptr(out[k]), npad,
                                                None, 0, _stream()))
    return SoaCloud(out, n)


# --------------------------------------------------------------------------------------
# K2 -- nearest neighbour, Chamfer, ADD-S
# --------------------------------------------------------------------------------------
@dataclasses.dataclass
class NNResult:
    d2: torch.Tensor            # float32 [B, Nq] squared distance (kernel arithmetic, bit-exact)
    idx: Optional[torch.Tensor]  # int32 [B, Nq] nearest target index (lowest index on ties)

    @property
    def dist(self) -> torch.Tensor:
        """Euclidean distances, float64 (what compute_point_cloud_distance returns)."""
        return torch.sqrt(self.d2.to(torch.float64))


@_on_device
def nearest_neighbors_soa(q: SoaCloud, t: SoaCloud, return_index: bool = True,
                          use_lo: bool = True) -> NNResult:
    """K2 on prepared clouds.  7-plane clouds run the production filtered-exact kernel
    (isr_nn2), 3-plane clouds the direct-difference kernel (isr_nn_soa)."""
    if t.n < 1:
        raise ValueError("nearest_neighbors: empty target cloud")
    if q.planes != t.planes:
        raise ValueError("query and target were prepared for different kernels")
    qb, tb = q.batch, t.batch
    batch = max(qb, tb)
    if qb not in (1, batch) or tb not in (1, batch):
        raise ValueError(f"batch mismatch: query {qb}, target {tb}")
    device = q.data.device
    d2 = torch.empty((batch, q.n), dtype=torch.float32, device=device)
    idx = torch.empty((batch, q.n), dtype=torch.int32, device=device) if return_index else None
    if q.n == 0:
        return NNResult(d2, idx)
    lib = _lib.load()
    pl = q.planes
    for b0 in range(0, batch, 65535):
        bc = min(65535, batch - b0)
        qd = q.data[b0:] if qb > 1 else q.data
        td = t.data[b0:] if tb > 1 else t.data
        if pl == 7:
            ws = _workspace(lib.isr_nn2_workspace_bytes(q.n, t.n, bc), device)
            qdesc = _lib.IsrCloud(qd.data_ptr(), q.n, q.npad, pl * q.npad if qb > 1 else 0, None,
                                  _ptr(q.perm), None, None, None)
            tsc = None if t.stage_c is None else (t.stage_c[b0:] if tb > 1 else t.stage_c)
            tsub = None if t.sub_c is None else (t.sub_c[b0:] if tb > 1 else t.sub_c)
            tbox = None if t.sub_box is None else (t.sub_box[b0:] if tb > 1 else t.sub_box)
            tdesc = _lib.IsrCloud(td.data_ptr(), t.n, t.npad, pl * t.npad if tb > 1 else 0,
                                  _ptr(tsc), _ptr(t.perm), _ptr(tsub), None, _ptr(tbox))
            _lib.check(lib.isr_nn2(
                ctypes.byref(qdesc), ctypes.byref(tdesc), bc, 1 if use_lo else 0, _ptr(d2[b0:]),
                _ptr(idx[b0:]) if idx is not None else None, None, 0, _ptr(ws), ws.numel(),
                _stream()))
        else:
            ws = _workspace(lib.isr_nn_workspace_bytes(q.n, t.n, bc), device)
            _lib.check(lib.isr_nn_soa(
                _ptr(qd), q.n, q.npad, pl * q.npad if qb > 1 else 0,
                _ptr(td), t.n, t.npad, pl * t.npad if tb > 1 else 0,
                bc, _ptr(d2[b0:]), _ptr(idx[b0:]) if idx is not None else None, None, 0,
                _ptr(ws), ws.numel(), _stream()))
    return NNResult(d2, idx)


@_on_device
def nearest_neighbors(query, target, return_index: bool = True, mode: str = "exact",
                      device=None) -> NNResult:
    """Brute-force 1-NN of query [Nq,3] / [B,Nq,3] in target [Nt,3] / [B,Nt,3].

    mode='exact' (default): FP32 filter + FP64 resolve; the index is the float64 argmin,
    d2 the float64 squared distance rounded to float32.  mode='direct': FP32
    direct-difference kernel; d2/idx follow fma(dz,dz,fma(dy,dy,dx*dx)) bit for bit."""
    if mode not in ("exact", "direct"):
        raise ValueError("mode must be 'exact' or 'direct'")
    device = _device(device)
    qp, tp = _points(query, device), _points(target, device)
    single = qp.dim() == 2 and tp.dim() == 2
    if mode == "direct":
        res = nearest_neighbors_soa(_pack_batched(qp, device), _pack_batched(tp, device), return_index)
    elif tp.dim() == 2:
        if tp.shape[0] < 1:
            raise ValueError("nearest_neighbors: empty target cloud")
        cen = centroid_of(tp, device)
        # (prepare_cloud gets the ORIGINAL arrays: float64 clouds are split into hi/lo pairs)
        t7 = prepare_cloud(target, centroid=cen, perm=spatial_order(tp, device), stage_centroids=True,
                           device=device)
        if qp.dim() == 2:
            q7 = prepare_cloud(query, centroid=cen, perm=spatial_order(qp, device), device=device)
            res = nearest_neighbors_soa(q7, t7, return_index)
        else:
            # a batch of different query clouds against one target: each gets its own curve
            # order (the pruned search relies on it), so they are searched one by one
            outs = [nearest_neighbors_soa(
                prepare_cloud(query[k], centroid=cen, perm=spatial_order(qp[k], device), device=device),
                t7, return_index) for k in range(qp.shape[0])]
            res = NNResult(torch.cat([o.d2 for o in outs]),
                           torch.cat([o.idx for o in outs]) if return_index else None)
    else:
        outs = []
        for k in range(tp.shape[0]):
            qk = query if qp.dim() == 2 else query[k]
            outs.append(nearest_neighbors(qk, target[k], return_index, mode, device))
        return NNResult(torch.stack([o.d2 for o in outs]),
                        torch.stack([o.idx for o in outs]) if return_index else None)
    if single:
        res = NNResult(res.d2[0], None if res.idx is None else res.idx[0])
    return res


@_on_device
def radius_neighbor_count(points, radius: float, target=None, device=None) -> torch.Tensor:
    """int32 [N]: for every point, the number of `target` points (default: the cloud itself,
    the point included) with d^2 < radius^2 -- strict, decided in float64.  The count behind
    Open3D's remove_radius_outlier (generateCors.py:254-258, trainPose.py:343-347)."""
    device = _device(device)
    if not radius > 0:
        raise ValueError("radius must be positive")
    qp, qlo = _points_hilo(points, device)
    if qp.dim() != 2:
        raise ValueError("radius_neighbor_count expects points of shape [N, 3]")
    out = torch.zeros((qp.shape[0],), dtype=torch.int32, device=device)
    if qp.shape[0] == 0:
        return out
    if target is None:
        cen = centroid_of(qp, device)
        q7 = t7 = prepare_cloud(points, centroid=cen, perm=spatial_order(qp, device), stage_centroids=True,
                                device=device)
    else:
        tp = _points(target, device)
        if tp.shape[0] == 0:
            return out
        cen = centroid_of(tp, device)
        t7 = prepare_cloud(target, centroid=cen, perm=spatial_order(tp, device), stage_centroids=True,
                           device=device)
        q7 = prepare_cloud(points, centroid=cen, perm=spatial_order(qp, device), device=device)
    qd, tdesc = q7.descriptor(batched=False), t7.descriptor(batched=False)
    _lib.check(_lib.load().isr_radius_count(ctypes.byref(qd), ctypes.byref(tdesc), float(radius),
                                            _ptr(out), _stream()))
    return out


@_on_device
def estimate_normals(points, neighborhood_size: int = 50, disambiguate_directions: bool = True,
                     device=None) -> torch.Tensor:
    """float32 [N,3]: pytorch3d.ops.estimate_pointcloud_normals(points[None], neighborhood_size,
    disambiguate_directions)[0] (generateCors.py:200-215 calls it with 400 neighbours on 1000
    points and negates the result): k-nearest-neighbour PCA, smallest-eigenvalue eigenvector,
    majority-side direction (isr_knn_normals)."""
    device = _device(device)
    pts = _points(points, device)
    if pts.dim() != 2:
        raise ValueError("estimate_normals expects points of shape [N, 3]")
    n = pts.shape[0]
    k = int(neighborhood_size)
    if not 1 <= k <= n:
        raise ValueError(f"neighborhood_size {k} outside 1..{n}")
    out = torch.empty((n, 3), dtype=torch.float32, device=device)
    _lib.check(_lib.load().isr_knn_normals(_ptr(pts), n, k, 1 if disambiguate_directions else 0, _ptr(out),
                                           _stream()))
    return out


def _mean_sqrt(d2: torch.Tensor) -> torch.Tensor:
    """FP64 mean of sqrt(d2) per row, deterministic order."""
    b, n = d2.shape
    out = torch.empty((b,), dtype=torch.float64, device=d2.device)
    _lib.check(_lib.load().isr_mean_sqrt(_ptr(d2), n, b, _ptr(out), _stream()))
    return out


@_on_device
def point_cloud_distance(source, target, device=None) -> torch.Tensor:
    """Open3D compute_point_cloud_distance: float64 [N] distances; empty target -> zeros."""
    device = _device(device)
    src = _points(source, device)
    tgt = _points(target, device)
    if tgt.shape[0] == 0:
        return torch.zeros((src.shape[0],), dtype=torch.float64, device=device)
    if src.shape[0] == 0:
        return torch.zeros((0,), dtype=torch.float64, device=device)
    # (the ORIGINAL arrays go to the search: float64 clouds keep their precision as hi/lo pairs)
    return nearest_neighbors(source, target, return_index=False, device=device).dist


@_on_device
def chamfer_distance(a, b, device=None) -> torch.Tensor:
    """(mean d(a->b) + mean d(b->a)) / 2, unsquared, float64 scalar tensor
    (verfication.py:97-101).  Accepts [N,3] or batched [B,N,3]."""
    device = _device(device)
    pa, pb = _points(a, device), _points(b, device)
    single = pa.dim() == 2 and pb.dim() == 2
    if not single:
        if pa.dim() != 3 or pb.dim() != 3 or pa.shape[0] != pb.shape[0]:
            raise ValueError("batched chamfer_distance expects [B,N,3] and [B,M,3]")
        return torch.stack([chamfer_distance(a[k], b[k], device) for k in range(pa.shape[0])])
    if pa.shape[0] == 0 or pb.shape[0] == 0:
        raise ValueError("chamfer_distance: empty cloud")
    cen = centroid_of(pb, device)
    # (float64 inputs -- icp.py:110-113 builds a float64 merged cloud -- keep their precision)
    A = prepare_cloud(a, centroid=cen, perm=spatial_order(pa, device), stage_centroids=True,
                      device=device)
    B = prepare_cloud(b, centroid=cen, perm=spatial_order(pb, device), stage_centroids=True,
                      device=device)
    ab = _mean_sqrt(nearest_neighbors_soa(A, B, return_index=False).d2)
    ba = _mean_sqrt(nearest_neighbors_soa(B, A, return_index=False).d2)
    return ((ab + ba) / 2)[0]


@dataclasses.dataclass
class VerifyResult:
    losses: torch.Tensor   # float64 [B] on device; +inf for invalid candidates
    best: torch.Tensor     # int64 [2] on device: {argmin index, bits of the float64 loss}

    @property
    def best_index(self) -> int:
        return int(self.best[0].item())

    @property
    def best_loss(self) -> float:
        return float(self.best[1:2].view(torch.float64).item())


@_on_device
def verify_poses(cloud_q, poses_q, poses_t, cloud_t=None, mode: str = "chamfer",
                 valid_mask=None, device=None) -> VerifyResult:
    """Score B candidate poses and select the first minimum, entirely on the device.

    Candidate k compares ``poses_q[k] . cloud_q`` with ``poses_t[k] . cloud_t``
    (cloud_t defaults to cloud_q).  mode='chamfer': bidirectional mean distance / 2
    (verfication.py:97-101); mode='adds': one-directional mean distance query->target
    (choosePose.py:20-22).
    """
    if mode not in ("chamfer", "adds"):
        raise ValueError("mode must be 'chamfer' or 'adds'")
    device = _device(device)
    cq = _points(cloud_q, device)
    ct = cq if cloud_t is None else _points(cloud_t, device)
    Pq = _poses(poses_q, device)
    Pt = _poses(poses_t, device)
    if Pq.shape[0] != Pt.shape[0]:
        raise ValueError("poses_q and poses_t must have the same batch size")
    b = Pq.shape[0]
    if b == 0:
        raise ValueError("verify_poses: empty candidate list")
    if cq.shape[0] == 0 or ct.shape[0] == 0:
        raise ValueError("verify_poses: empty cloud")
    valid = None
    if valid_mask is not None:
        valid = _to_dev(np.asarray(valid_mask).astype(np.uint8) if not isinstance(
            valid_mask, torch.Tensor) else valid_mask.to(torch.uint8), torch.uint8, device)
        if valid.numel() != b:
            raise ValueError("valid_mask must have one entry per candidate")
    bidir = 1 if mode == "chamfer" else 0
    lib = _lib.load()
    ws = _workspace(lib.isr_verify_workspace_bytes(cq.shape[0], ct.shape[0], b, bidir), device)
    losses = torch.empty((b,), dtype=torch.float64, device=device)
    best = torch.empty((2,), dtype=torch.int64, device=device)
    _lib.check(lib.isr_verify_poses(_ptr(cq), cq.shape[0], _ptr(ct), ct.shape[0], _ptr(Pq), _ptr(Pt),
                                    _ptr(valid), b, bidir, _ptr(losses), _ptr(best), _ptr(ws),
                                    ws.numel(), _stream()))
    return VerifyResult(losses, best)


@_on_device
def adds(verts, gtR, gtT, R, T, surface_points, device=None) -> torch.Tensor:
    """Batched ADDS (choosePose.py:20-22): mean 1-NN distance from verts.gtR^T+gtT to
    surface.R^T+T.  gtR/R may be [3,3] or [B,3,3]; returns float64 [B] (or scalar)."""
    gtR = np.asarray(gtR, dtype=np.float64)
    single = gtR.ndim == 2
    gtR = gtR.reshape(-1, 3, 3)
    R = np.asarray(R, dtype=np.float64).reshape(-1, 3, 3)
    gtT = np.asarray(gtT, dtype=np.float64).reshape(-1, 3)
    T = np.asarray(T, dtype=np.float64).reshape(-1, 3)
    b = len(gtR)
    Pq = np.tile(np.eye(4), (b, 1, 1))
    Pt = np.tile(np.eye(4), (b, 1, 1))
    Pq[:, :3, :3], Pq[:, :3, 3] = gtR, gtT
    Pt[:, :3, :3], Pt[:, :3, 3] = R, T
    res = verify_poses(verts, Pq, Pt, cloud_t=surface_points, mode="adds", device=device)
    return res.losses[0] if single else res.losses


# --------------------------------------------------------------------------------------
# the callers of the batched ADD-S in choosePose.py, on the device
# --------------------------------------------------------------------------------------
@_on_device
def relative_pose_table(RList, TList, pair0: int = 0, count: Optional[int] = None, device=None) -> torch.Tensor:
    """float64 [count, 4, 4] on the device: relative_poses[i][j] of choosePose.py:98-107 for the
    flat pair indices k = i * n + j in [pair0, pair0 + count) (default: the whole n x n table)."""
    device = _device(device)
    R = _to_dev(np.asarray(RList, dtype=np.float64).reshape(-1, 9) if not isinstance(RList, torch.Tensor)
                else RList.reshape(-1, 9), torch.float64, device)
    t = _to_dev(np.asarray(TList, dtype=np.float64).reshape(-1, 3) if not isinstance(TList, torch.Tensor)
                else TList.reshape(-1, 3), torch.float64, device)
    n = R.shape[0]
    if t.shape[0] != n:
        raise ValueError("RList and TList must have the same length")
    count = n * n - pair0 if count is None else int(count)
    out = torch.empty((count, 4, 4), dtype=torch.float64, device=device)
    _lib.check(_lib.load().isr_rel_pose_table(_ptr(R), _ptr(t), n, int(pair0), count, _ptr(out), _stream()))
    return out


@_on_device
def adds_rigid(verts, poses_gt, poses_pred, surface_points, valid_mask=None, device=None) -> VerifyResult:
    """ADDS(verts, gtR, gtT, R, T) of choosePose.py:20-22 for B pose pairs (poses_gt, poses_pred:
    [B,4,4], host or device) with the SURFACE PREPARED ONCE: candidate k scores
    (poses_pred[k]^-1 poses_gt[k]) . verts against the surface cloud itself, which is the same
    distance whenever poses_pred[k] is a rigid motion.  No host synchronisation."""
    device = _device(device)
    V, S = _points(verts, device), _points(surface_points, device)
    Pq, Pt = _poses(poses_gt, device), _poses(poses_pred, device)
    if Pq.shape[0] != Pt.shape[0]:
        raise ValueError("poses_gt and poses_pred must have the same batch size")
    b = Pq.shape[0]
    if b == 0 or V.shape[0] == 0 or S.shape[0] == 0:
        raise ValueError("adds_rigid: empty input")
    valid = None
    if valid_mask is not None:
        valid = _to_dev(np.asarray(valid_mask).astype(np.uint8) if not isinstance(
            valid_mask, torch.Tensor) else valid_mask.to(torch.uint8), torch.uint8, device)
    lib = _lib.load()
    M = torch.empty_like(Pq)
    _lib.check(lib.isr_rigid_relative(_ptr(Pq), _ptr(Pt), b, _ptr(M), _stream()))
    ws = _workspace(lib.isr_adds_fixed_target_workspace_bytes(V.shape[0], S.shape[0], b), device)
    losses = torch.empty((b,), dtype=torch.float64, device=device)
    best = torch.empty((2,), dtype=torch.int64, device=device)
    _lib.check(lib.isr_adds_fixed_target(_ptr(V), V.shape[0], _ptr(S), S.shape[0], _ptr(M), _ptr(valid), b,
                                         _ptr(losses), _ptr(best), _ptr(ws), ws.numel(), _stream()))
    return VerifyResult(losses, best)


@_on_device
def rigid_relative(poses_q, poses_t, device=None) -> torch.Tensor:
    """M[k] = poses_t[k]^-1 . poses_q[k] for rigid poses_t (float64 [B,4,4] on the device)."""
    device = _device(device)
    Pq, Pt = _poses(poses_q, device), _poses(poses_t, device)
    if Pq.shape[0] != Pt.shape[0]:
        raise ValueError("poses_q and poses_t must have the same batch size")
    M = torch.empty_like(Pq)
    _lib.check(_lib.load().isr_rigid_relative(_ptr(Pq), _ptr(Pt), Pq.shape[0], _ptr(M), _stream()))
    return M


@_on_device
def adds_fixed(verts, poses, surface_points, device=None) -> VerifyResult:
    """Mean 1-NN distance from poses[k] . verts into the surface cloud ITSELF (prepared once)."""
    device = _device(device)
    V, S, M = _points(verts, device), _points(surface_points, device), _poses(poses, device)
    b = M.shape[0]
    if b == 0 or V.shape[0] == 0 or S.shape[0] == 0:
        raise ValueError("adds_fixed: empty input")
    lib = _lib.load()
    ws = _workspace(lib.isr_adds_fixed_target_workspace_bytes(V.shape[0], S.shape[0], b), device)
    losses = torch.empty((b,), dtype=torch.float64, device=device)
    best = torch.empty((2,), dtype=torch.int64, device=device)
    _lib.check(lib.isr_adds_fixed_target(_ptr(V), V.shape[0], _ptr(S), S.shape[0], _ptr(M), None, b,
                                         _ptr(losses), _ptr(best), _ptr(ws), ws.numel(), _stream()))
    return VerifyResult(losses, best)


@_on_device
def adds_bounds(verts, poses, target: SoaCloud, presorted: bool = False, device=None):
    """(lower, upper) float64 [B] with lower <= ADD-S(poses[k] . verts -> target) <= upper, from
    the tile spheres of a target prepared with ``prepare_cloud(..., stage_centroids=True)``
    alone (isr_adds_bounds).  The kernel tests the stage spheres once per 32 consecutive
    vertices, so the vertices are put in curve order first unless `presorted`.  Returns None
    when the target's spheres do not fit shared memory."""
    device = _device(device)
    V, M = _points(verts, device), _poses(poses, device)
    if not presorted and V.shape[0] > 32:
        V = V[spatial_order(V, device).to(torch.int64)].contiguous()
    stages = target.npad // _lib.ISR_SOA_TILE
    if target.batch != 1 or target.stage_c is None or target.sub_c is None or target.centroid is None:
        raise ValueError("adds_bounds: the target must be one prepared cloud with its tile spheres")
    if stages * 17 * 16 > 200 * 1024:
        return None
    b = M.shape[0]
    lo = torch.empty((b,), dtype=torch.float64, device=device)
    hi = torch.empty((b,), dtype=torch.float64, device=device)
    _lib.check(_lib.load().isr_adds_bounds(_ptr(V), V.shape[0], _ptr(M), b, _ptr(target.centroid),
                                           _ptr(target.stage_c), stages, _ptr(target.sub_c), _ptr(lo), _ptr(hi),
                                           _stream()))
    return lo, hi


@_on_device
def vote(losses, threshold: float, device=None):
    """choosePose.py:135-151 on a loss table [rows, cols] (device float64): returns
    (error uint8 [rows, cols], votes int32 [rows], best int64 [2] = {first argmax, its votes})."""
    device = _device(device)
    L = _to_dev(losses, torch.float64, device)
    if L.dim() != 2:
        raise ValueError("vote expects a [rows, cols] loss table")
    rows, cols = L.shape
    error = torch.empty((rows, cols), dtype=torch.uint8, device=device)
    votes = torch.empty((rows,), dtype=torch.int32, device=device)
    best = torch.empty((2,), dtype=torch.int64, device=device)
    _lib.check(_lib.load().isr_vote(_ptr(L), rows, cols, float(threshold), _ptr(error), _ptr(votes), _ptr(best),
                                    _stream()))
    return error, votes, best


# --------------------------------------------------------------------------------------
# K3 -- ICP
# --------------------------------------------------------------------------------------
@dataclasses.dataclass
class IcpResult:
    """Mirror of Open3D's RegistrationResult (icp.py:99,104-106)."""
    transformation: np.ndarray
    fitness: float
    inlier_rmse: float
    n_corr: int
    iterations: int
    converged: bool
    _corr_idx: Optional[torch.Tensor] = None
    _inlier: Optional[torch.Tensor] = None

    @property
    def correspondence_set(self) -> np.ndarray:
        """k x 2 int32 (source index, target index), like Open3D's correspondence_set_."""
        if self._corr_idx is None:
            return np.zeros((0, 2), dtype=np.int32)
        src = torch.nonzero(self._inlier, as_tuple=False)[:, 0]
        tgt = self._corr_idx[src].to(torch.int64)
        return torch.stack([src, tgt], dim=1).to(torch.int32).cpu().numpy()

    def __repr__(self) -> str:
        return (f"RegistrationResult with fitness={self.fitness:e}, inlier_rmse={self.inlier_rmse:e},"
                f" and correspondence_set size of {self.n_corr}\n"
                "Access transformation to get result.")


def _on_self_device(fn):
    """Method form of _on_device: run on the device the object's buffers live on."""
    @functools.wraps(fn)
    def wrapper(self, *args, **kwargs):
        with torch.cuda.device(self.device):
            return fn(self, *args, **kwargs)

    return wrapper


class IcpProblem:
    """Device-side buffers of one (source, target) pair for `starts` simultaneous ICP starts.

    Exposed so that the multi-GPU driver (dist.py) can interleave its all-reduce between
    `accumulate` and `solve`; single-GPU callers use `run`.
    """

    def __init__(self, source, target, inits, device=None):
        self.device = _device(device)
        with torch.cuda.device(self.device):
            self._init(source, target, inits)

    def _init(self, source, target, inits):
        self.src, self.src_lo = _points_hilo(source, self.device)
        self.tgt = _points(target, self.device)
        if self.src.dim() != 2 or self.tgt.dim() != 2:
            raise ValueError("icp: source and target must be [N,3]")
        self.ns, self.nt = self.src.shape[0], self.tgt.shape[0]
        if self.nt < 1:
            raise ValueError("icp: empty target cloud")
        inits = np.asarray(inits, dtype=np.float64).reshape(-1, 4, 4)
        self.starts = len(inits)
        st = np.zeros(self.starts, dtype=_lib.ICP_STATE_DTYPE)
        st["T"] = inits.reshape(self.starts, 16)
        self.states = torch.from_numpy(st.view(np.uint8).reshape(self.starts, -1).copy()).to(self.device)
        self.centroid = centroid_of(self.tgt, self.device)
        self.tgt_soa = prepare_cloud(self.tgt, centroid=self.centroid,
                                     perm=spatial_order(self.tgt, self.device), stage_centroids=True,
                                     device=self.device)
        self.tgt_desc = self.tgt_soa.descriptor(batched=False)
        self.src_perm = spatial_order(self.src, self.device) if self.src.shape[0] > 0 else None
        lib = _lib.load()
        ns1 = max(self.ns, 1)
        self.ws = _workspace(lib.isr_icp_workspace_bytes(ns1, self.nt, self.starts), self.device)
        self.sums = torch.zeros((self.starts, _lib.ISR_ICP_NSUMS), dtype=torch.float64,
                                device=self.device)
        self.corr_idx = torch.zeros((self.starts, ns1), dtype=torch.int32, device=self.device)
        self.inlier = torch.zeros((self.starts, ns1), dtype=torch.uint8, device=self.device)

    @_on_self_device
    def accumulate(self, max_dist: float) -> torch.Tensor:
        if self.ns == 0:
            self.sums.zero_()
            return self.sums
        _lib.check(_lib.load().isr_icp_accumulate(
            _ptr(self.states), self.starts, _ptr(self.src), _ptr(self.src_lo), _ptr(self.src_perm),
            self.ns, _ptr(self.tgt), ctypes.byref(self.tgt_desc), _ptr(self.centroid), float(max_dist),
            _ptr(self.sums), _ptr(self.corr_idx), _ptr(self.inlier), _ptr(self.ws), self.ws.numel(),
            _stream()))
        return self.sums

    # -- the two halves of `accumulate`, for target-sharded ICP (dist.py) ---------------------
    @_on_self_device
    def search(self) -> torch.Tensor:
        """Transform the source by every state's T and find each point's nearest neighbour in
        THIS problem's target -> int32 [starts, ns] (also kept in self.corr_idx)."""
        _lib.check(_lib.load().isr_icp_search(
            _ptr(self.states), self.starts, _ptr(self.src), _ptr(self.src_lo), _ptr(self.src_perm),
            self.ns, ctypes.byref(self.tgt_desc), _ptr(self.centroid), _ptr(self.corr_idx),
            _ptr(self.ws), self.ws.numel(), _stream()))
        return self.corr_idx

    @_on_self_device
    def corr_dist(self, corr_idx: torch.Tensor) -> torch.Tensor:
        """float64 [starts, ns]: exact squared distance of every given correspondence (+inf
        where the index is negative), in the accumulate kernel's arithmetic."""
        if not hasattr(self, "_D"):
            self._D = torch.full((self.starts, max(self.ns, 1)), float("inf"), dtype=torch.float64,
                                 device=self.device)
        _lib.check(_lib.load().isr_icp_corr_dist(
            _ptr(self.states), self.starts, _ptr(self.src), _ptr(self.src_lo), self.ns, _ptr(self.tgt),
            _ptr(corr_idx), _ptr(self._D), _stream()))
        return self._D

    @_on_self_device
    def accumulate_corr(self, corr_idx: torch.Tensor, max_dist: float) -> torch.Tensor:
        """The 17 sums over the given correspondences (index < 0: none on this rank)."""
        _lib.check(_lib.load().isr_icp_accumulate_corr(
            _ptr(self.states), self.starts, _ptr(self.src), _ptr(self.src_lo), _ptr(self.src_perm), self.ns,
            _ptr(self.tgt), self.nt, _ptr(corr_idx), float(max_dist), _ptr(self.sums), _ptr(self.inlier),
            _ptr(self.ws), self.ws.numel(), _stream()))
        return self.sums

    @_on_self_device
    def solve(self, ns_total: int, rel_fitness: float, rel_rmse: float, final_eval: bool,
              sums: Optional[torch.Tensor] = None) -> None:
        sums = self.sums if sums is None else sums
        _lib.check(_lib.load().isr_icp_solve(_ptr(self.states), self.starts, _ptr(sums), int(ns_total),
                                             float(rel_fitness), float(rel_rmse),
                                             1 if final_eval else 0, _stream()))

    @_on_self_device
    def run(self, max_dist: float, max_iteration: int, rel_fitness: float, rel_rmse: float) -> None:
        if self.ns == 0:
            for k in range(max_iteration + 1):
                self.accumulate(max_dist)
                self.solve(0, rel_fitness, rel_rmse, k == max_iteration)
            return
        _lib.check(_lib.load().isr_icp_run(
            _ptr(self.states), self.starts, _ptr(self.src), _ptr(self.src_lo), _ptr(self.src_perm),
            self.ns, _ptr(self.tgt), ctypes.byref(self.tgt_desc), _ptr(self.centroid), float(max_dist),
            int(max_iteration),
            float(rel_fitness), float(rel_rmse), _ptr(self.sums), _ptr(self.corr_idx),
            _ptr(self.inlier), _ptr(self.ws), self.ws.numel(), _stream()))

    def reopen(self) -> None:
        """Clear `done` of every start (offset 176 of IsrIcpState) so that another run continues
        from the current poses and correspondences."""
        self.states.view(torch.int32)[:, 44] = 0

    @_on_self_device
    def run_sharded(self, peer, ns_total: int, max_dist: float, max_iteration: int, rel_fitness: float,
                    rel_rmse: float) -> None:
        """`run` for one source shard: the 17 sums are exchanged between the ranks of `peer`
        (dist.PeerExchange) inside the accumulate / solve kernels (isr_icp_run_sharded)."""
        _lib.check(_lib.load().isr_icp_run_sharded(
            _ptr(self.states), self.starts, _ptr(self.src), _ptr(self.src_lo), _ptr(self.src_perm),
            self.ns, int(ns_total), _ptr(self.tgt), ctypes.byref(self.tgt_desc), _ptr(self.centroid),
            float(max_dist), int(max_iteration), float(rel_fitness), float(rel_rmse), _ptr(self.sums),
            _ptr(self.corr_idx), _ptr(self.inlier), _ptr(self.ws), self.ws.numel(), peer.handle,
            _stream()))

    @_on_self_device
    def results(self, with_correspondences: bool = True) -> list:
        st = self.states.cpu().numpy().view(_lib.ICP_STATE_DTYPE).reshape(self.starts)
        out = []
        for k in range(self.starts):
            s = st[k]
            if int(s["reserved"]) == 1:
                raise RuntimeError("sharded ICP: a peer rank never delivered its sums (exchange timed "
                                   "out inside icp_solve_kernel); results are invalid")
            out.append(IcpResult(
                transformation=s["T"].reshape(4, 4).copy(), fitness=float(s["fitness"]),
                inlier_rmse=float(s["inlier_rmse"]), n_corr=int(s["n_corr"]),
                iterations=int(s["iters"]), converged=bool(s["done"]),
                _corr_idx=self.corr_idx[k] if with_correspondences and self.ns > 0 else None,
                _inlier=self.inlier[k] if with_correspondences and self.ns > 0 else None))
        return out


@_on_device
def evaluate_registration(source, target, max_correspondence_distance: float,
                          transformation=None, device=None) -> IcpResult:
    """o3d.pipelines.registration.evaluate_registration (icp.py:97-98)."""
    T = np.eye(4) if transformation is None else np.asarray(transformation, dtype=np.float64)
    prob = IcpProblem(source, target, T[None], device)
    prob.run(max_correspondence_distance, 0, 0.0, 0.0)
    return prob.results()[0]


@_on_device
def icp(source, target, init=None, max_correspondence_distance: float = 20.0,
        max_iteration: int = 30, relative_fitness: float = 1e-6, relative_rmse: float = 1e-6,
        device=None) -> IcpResult:
    """Point-to-point ICP with Open3D's loop and default criteria (icp.py:101-103)."""
    T = np.eye(4) if init is None else np.asarray(init, dtype=np.float64)
    prob = IcpProblem(source, target, T[None], device)
    prob.run(max_correspondence_distance, max_iteration, relative_fitness, relative_rmse)
    return prob.results()[0]


@dataclasses.dataclass
class MultiStartResult:
    results: list               # IcpResult per start
    chamfer: np.ndarray         # float64 [S] Chamfer(transformed source, target) per start
    order: np.ndarray           # starts sorted by ascending Chamfer (stable: first minimum first)

    @property
    def best(self) -> IcpResult:
        return self.results[int(self.order[0])]


@_on_device
def multistart_icp(source, target, inits, max_correspondence_distance: float = 20.0,
                   max_iteration: int = 30, relative_fitness: float = 1e-6,
                   relative_rmse: float = 1e-6, device=None) -> MultiStartResult:
    """All starts advance together (batch dimension in every kernel); each is then scored
    by the Chamfer distance of its registered source against the target (icp.py:113-117)
    and ranked ascending, first minimum first."""
    inits = np.asarray(inits, dtype=np.float64).reshape(-1, 4, 4)
    prob = IcpProblem(source, target, inits, device)
    prob.run(max_correspondence_distance, max_iteration, relative_fitness, relative_rmse)
    res = prob.results()
    Ts = np.stack([r.transformation for r in res])
    s = len(res)
    vr = verify_poses(prob.src, Ts, np.tile(np.eye(4), (s, 1, 1)), cloud_t=prob.tgt, mode="chamfer",
                      device=prob.device)
    ch = vr.losses.cpu().numpy()
    return MultiStartResult(res, ch, np.argsort(ch, kind="stable"))


@_on_device
def icp_refine_pose(R, t, source, target, max_correspondence_distance: float = 20.0,
                    max_iteration: int = 30, relative_fitness: float = 1e-6,
                    relative_rmse: float = 1e-6, device=None):
    """Point-to-point ICP refinement of the pose (R, t) that maps `source` into the frame of
    `target`, returned in the convention of pose_refine.py:21-22,101-104: ``(R 3x3 float64,
    t (3,) float64, loss)`` with loss = the inlier RMSE of the final evaluation.

    This is NOT the reference's ``refine_pose`` objective: that one keeps R frozen and runs BFGS
    on the translation against a NeRF-key / GL-renderer likelihood (pose_refine.py:74-98), is
    reachable only behind a hard-coded ``useSurfEval=False`` (inference.py:27) and needs modules
    the reference does not ship; it is out of scope.  Only the
    ``(R, t, loss)`` return shape is carried over, so that a caller that unpacks the
    reference's triple keeps working when it refines with ICP instead."""
    src = np.asarray(source) if not isinstance(source, torch.Tensor) else source
    tg = np.asarray(target) if not isinstance(target, torch.Tensor) else target
    if src.ndim != 2 or src.shape[-1] != 3 or tg.ndim != 2 or tg.shape[-1] != 3:
        raise TypeError("icp_refine_pose(R, t, source[N,3], target[M,3], ...): this is the ICP "
                        "refinement, not pose_refine.refine_pose(R, t, query_img, renderer, ...)")
    res = icp(source, target, pose_from_Rt(R, t), max_correspondence_distance, max_iteration,
              relative_fitness, relative_rmse, device=device)
    T = res.transformation
    return T[:3, :3].copy(), T[:3, 3].copy(), float(res.inlier_rmse)


#: name used by SURVEY.md section 8(b) / round 1; same function
refine_pose = icp_refine_pose


# --------------------------------------------------------------------------------------
# PnP hypothesis scoring
# --------------------------------------------------------------------------------------
@_on_device
def score_pnp_hypotheses(points3d, points2d, camera_matrix, poses, reprojection_error: float = 2.0,
                         return_inliers: bool = False, device=None):
    """Consensus of every hypothesis in `poses` [B,4,4] (object -> camera) over the 2-D/3-D
    correspondences (points3d [n,3], points2d [n,2] pixels): the test inside
    cv2.solvePnPRansac as choosePose.py:23-33,280-300 uses it.  Returns int32 counts [B] on the
    device (and the uint8 flags [B, n] when asked)."""
    device = _device(device)
    p3 = _points(points3d, device)
    p2 = _to_dev(points2d, torch.float32, device)
    if p3.dim() != 2 or p2.dim() != 2 or p2.shape != (p3.shape[0], 2):
        raise ValueError("score_pnp_hypotheses: points3d [n,3] and points2d [n,2] expected")
    K = _to_dev(np.asarray(camera_matrix, dtype=np.float64).reshape(3, 3), torch.float64, device)
    P = _poses(poses, device)
    n, b = p3.shape[0], P.shape[0]
    counts = torch.empty((b,), dtype=torch.int32, device=device)
    flags = torch.empty((b, n), dtype=torch.uint8, device=device) if return_inliers else None
    lib = _lib.load()
    step = 65535 * 16
    for b0 in range(0, max(b, 1), step):
        bc = min(step, b - b0)
        _lib.check(lib.isr_pnp_score(_ptr(p3), _ptr(p2), n, _ptr(K), _ptr(P[b0:]), bc,
                                     float(reprojection_error), _ptr(counts[b0:]),
                                     _ptr(flags[b0:]) if flags is not None else None, _stream()))
    return (counts, flags) if return_inliers else counts


@_on_device
def p3p_solve(points3, pixels3, camera_matrix, device=None):
    """All P3P solutions (Grunert, FP64) of B explicit samples: points3 [B,3,3] object points,
    pixels3 [B,3,2].  Returns (poses float64 [B,4,4,4] with NaN beyond the real solutions,
    counts int32 [B]) on the device -- the minimal solver behind SOLVEPNP_P3P."""
    device = _device(device)
    P = _to_dev(np.asarray(points3, dtype=np.float64).reshape(-1, 9), torch.float64, device)
    uv = _to_dev(np.asarray(pixels3, dtype=np.float64).reshape(-1, 6), torch.float64, device)
    K = _to_dev(np.asarray(camera_matrix, dtype=np.float64).reshape(9), torch.float64, device)
    b = P.shape[0]
    poses = torch.empty((b, 4, 4, 4), dtype=torch.float64, device=device)
    cnt = torch.empty((b,), dtype=torch.int32, device=device)
    _lib.check(_lib.load().isr_p3p_solve(_ptr(P), _ptr(uv), _ptr(K), b, _ptr(poses), _ptr(cnt), _stream()))
    return poses, cnt


@dataclasses.dataclass
class PnpResult:
    ok: bool                    # a pose was found (the winning hypothesis has at least one inlier)
    R: np.ndarray               # 3x3 float64
    t: np.ndarray               # (3,) float64
    inliers: np.ndarray         # int indices: consensus set of the winning hypothesis (cv2's `inliers`)
    consensus: int              # inliers of the returned (refitted) pose


@_on_device
def pnp_ransac(points3d, points2d, camera_matrix, iterations: int = 100, reprojection_error: float = 2.0,
               seed: int = 0, refine_rounds: int = 3, device=None) -> PnpResult:
    """cv2.solvePnPRansac(points3d, points2d, cam, None, iterationsCount=iterations,
    reprojectionError=reprojection_error, flags=cv2.SOLVEPNP_P3P) on the device
    (choosePose.py:23-33): P3P hypotheses, consensus, first best hypothesis, Gauss-Newton refit on
    its inliers (isr_pnp_ransac).  One host synchronisation, to read the result."""
    device = _device(device)
    p3 = _points(points3d, device)
    p2 = _to_dev(points2d, torch.float32, device)
    if p3.dim() != 2 or p2.dim() != 2 or p2.shape != (p3.shape[0], 2):
        raise ValueError("pnp_ransac: points3d [n,3] and points2d [n,2] expected")
    n = p3.shape[0]
    if n < 4:
        return PnpResult(False, np.eye(3), np.zeros(3), np.zeros((0,), dtype=np.int64), 0)
    K = _to_dev(np.asarray(camera_matrix, dtype=np.float64).reshape(9), torch.float64, device)
    lib = _lib.load()
    ws = _workspace(lib.isr_pnp_ransac_workspace_bytes(n, int(iterations)), device)
    pose = torch.empty((16,), dtype=torch.float64, device=device)
    counts = torch.empty((2,), dtype=torch.int32, device=device)
    flags = torch.empty((n,), dtype=torch.uint8, device=device)
    _lib.check(lib.isr_pnp_ransac(_ptr(p3), _ptr(p2), n, _ptr(K), int(iterations), int(seed) & (2 ** 64 - 1),
                                  float(reprojection_error), int(refine_rounds), _ptr(pose), _ptr(counts),
                                  _ptr(flags), _ptr(ws), ws.numel(), _stream()))
    c = counts.cpu().numpy()
    T = pose.cpu().numpy().reshape(4, 4)
    ok = bool(c[0] > 0 and np.isfinite(T).all())
    return PnpResult(ok, T[:3, :3].copy(), T[:3, 3].copy(), np.nonzero(flags.cpu().numpy())[0], int(c[1]))


# --------------------------------------------------------------------------------------
# measurement helper
# --------------------------------------------------------------------------------------
def measure_fp32_peak(packed: bool = False, iters: int = 4096, reps: int = 5) -> float:
    """FFMA-chain microbenchmark -> measured FP32 TFLOP/s of the current device."""
    device = _device()
    lib = _lib.load()
    import ctypes
    sm = ctypes.c_int(0)
    _lib.check(lib.isr_device_info(ctypes.byref(sm), None, None))
    sink = torch.zeros(4, dtype=torch.float32, device=device)
    flops = ctypes.c_double(0)
    blocks = sm.value * 8
    best = 0.0
    for _ in range(reps + 1):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(lib.isr_bench_ffma(blocks, iters, 1 if packed else 0, _ptr(sink),
                                      ctypes.byref(flops), _stream()))
        e1.record()
        e1.synchronize()
        ms = e0.elapsed_time(e1)
        best = max(best, flops.value / (ms * 1e-3) / 1e12)
    return best
