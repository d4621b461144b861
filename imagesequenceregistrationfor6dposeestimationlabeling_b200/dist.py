"""Multi-GPU sharding of the hot path over the GPUs of one box (one process per GPU,
``torch.distributed``; NCCL over NVLink/NVSwitch on GPUs, gloo in the CPU tests).

Candidate verification shards by candidate: every rank scores a contiguous block against
its own replica of the cloud (1.2 MB) -- no data-path collective.  The only exchange is
the selection: an exact first-minimum argmin built from two 8-byte MIN all-reduces
(float64 loss, then the lowest global index among the ranks holding that loss), i.e.
``list.index(min(list))`` of verfication.py:105-106 across ranks.

Single-pair ICP shards the SOURCE points by default; the target is replicated (12 MB at
1 M points, L2-resident).  Per iteration the 17 partial Kabsch sums per start are exchanged
between the accumulate and solve kernels; every rank then solves the same 3x3 problem, so no
broadcast is needed and all ranks hold bit-identical poses.  The exchange is fused into the
two kernels (`PeerExchange`: stores into the peers' HBM over NVLink, flag wait in the solve
kernel, fixed rank-order sum) -- no collective call, no extra launch, the loop is one C call;
an NCCL all-reduce per iteration is the alternative (`exchange="nccl"`).  Nothing is read
back to the host inside the loop.  

`shard="target"` is the north star's variant: every rank holds a contiguous slice of the
TARGET, searches it for all source points, and the ranks agree on each point's nearest
neighbour with two MIN all-reduces (exact float64 distance, then the lowest global target
index among the holders -- the single-GPU tie rule); each rank then accumulates the
correspondences that landed in its slice and the same 17-double SUM all-reduce follows.
It moves 12 bytes per source point per iteration where the source-sharded form moves 136
bytes in total, and the pruned search gains little from a thinner target, so it is the
slower of the two; it exists for targets that do not fit one GPU.

The compute back end is injected so the rank logic can be tested on CPU with gloo:
the product passes nothing and gets the CUDA path; tests pass oracle-backed scorers.
"""
from __future__ import annotations

import os
from typing import Callable, Optional

import numpy as np
import torch
import torch.distributed as td

_I64_MAX = np.iinfo(np.int64).max


def init_from_env(backend: Optional[str] = None) -> tuple:
    """torchrun-style rendezvous (RANK / LOCAL_RANK / WORLD_SIZE / MASTER_*) -> (rank, world)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", str(rank)))
    use_cuda = torch.cuda.is_available()
    if use_cuda:
        torch.cuda.set_device(local % torch.cuda.device_count())
    if world > 1 and not td.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        kwargs = {}
        if use_cuda and (backend or "nccl") == "nccl":
            kwargs["device_id"] = torch.device("cuda", torch.cuda.current_device())
        td.init_process_group(backend or ("nccl" if use_cuda else "gloo"), rank=rank,
                              world_size=world, **kwargs)
    return rank, world


def shard_bounds(n: int, rank: int, world: int) -> tuple:
    """Contiguous block [lo, hi) of rank `rank`: ceil(n / world) each, the tail may be short."""
    per = (n + world - 1) // world
    lo = min(rank * per, n)
    return lo, min(lo + per, n)


def _world(group) -> int:
    return td.get_world_size(group) if td.is_initialized() else 1


def _is_nccl(group) -> bool:
    return td.is_initialized() and td.get_backend(group) == "nccl"


def _all_reduce(t: torch.Tensor, op, group=None) -> None:
    """In-place all-reduce that also works for a CUDA tensor on a gloo group (two ranks sharing
    one GPU in the tests, or a box without NCCL): staged through the host."""
    if t.is_cuda and not _is_nccl(group):
        c = t.cpu()
        td.all_reduce(c, op=op, group=group)
        t.copy_(c)
    else:
        td.all_reduce(t, op=op, group=group)


def global_first_argmin(local_loss: torch.Tensor, local_index: torch.Tensor, group=None) -> tuple:
    """Exact cross-rank first-minimum.  local_loss: float64 [1] (use +inf when the rank has no
    candidate), local_index: int64 [1] GLOBAL index of the local first minimum.
    Returns (loss [1], index [1]) tensors, identical on every rank."""
    if _world(group) == 1:
        return local_loss, local_index
    m = local_loss.clone()
    _all_reduce(m, td.ReduceOp.MIN, group)
    cand = torch.where(local_loss == m, local_index, torch.full_like(local_index, _I64_MAX))
    _all_reduce(cand, td.ReduceOp.MIN, group)
    return m, cand


def _cuda_scorer(cloud_q, poses_q, poses_t, cloud_t, mode, valid_mask):
    from . import api

    res = api.verify_poses(cloud_q, poses_q, poses_t, cloud_t=cloud_t, mode=mode,
                           valid_mask=valid_mask)
    return res.losses, res.best[0:1], res.best[1:2].view(torch.float64)


def verify_poses_sharded(cloud_q, poses_q, poses_t, cloud_t=None, mode: str = "chamfer",
                         valid_mask=None, group=None, gather_losses: bool = False,
                         scorer: Optional[Callable] = None):
    """Every rank passes the FULL candidate arrays; it scores its own block and the ranks
    agree on the selection.  Returns (best_index [1] int64, best_loss [1] float64,
    losses): `losses` is this rank's block, or all B losses when gather_losses is set."""
    scorer = scorer or _cuda_scorer
    if td.is_initialized():
        rank, world = td.get_rank(group), td.get_world_size(group)
    else:
        rank, world = 0, 1
    b = len(poses_q)
    lo, hi = shard_bounds(b, rank, world)
    dev = None
    if hi > lo:
        vm = None if valid_mask is None else valid_mask[lo:hi]
        losses, lidx, lbest = scorer(cloud_q, poses_q[lo:hi], poses_t[lo:hi], cloud_t, mode, vm)
        dev = losses.device
        lidx = lidx.to(torch.int64) + lo
        lbest = lbest.to(torch.float64)
    else:  # more ranks than candidates
        dev = torch.device("cuda", torch.cuda.current_device()) if _is_nccl(group) else torch.device("cpu")
        losses = torch.zeros((0,), dtype=torch.float64, device=dev)
        lidx = torch.full((1,), _I64_MAX, dtype=torch.int64, device=dev)
        lbest = torch.full((1,), float("inf"), dtype=torch.float64, device=dev)
    best_loss, best_idx = global_first_argmin(lbest, lidx, group)
    if gather_losses and world > 1:
        per = (b + world - 1) // world
        pad = torch.full((per,), float("inf"), dtype=torch.float64, device=dev)
        pad[: hi - lo] = losses
        out = torch.empty((world * per,), dtype=torch.float64, device=dev)
        td.all_gather_into_tensor(out, pad, group=group)
        losses = out[:b]
    return best_idx, best_loss, losses


def _cuda_multistart(source, target, inits, max_dist, max_iteration, rel_fitness, rel_rmse):
    from . import api

    ms = api.multistart_icp(source, target, inits, max_dist, max_iteration, rel_fitness, rel_rmse)
    return ([r.transformation for r in ms.results], [r.fitness for r in ms.results],
            [r.inlier_rmse for r in ms.results], [r.iterations for r in ms.results], ms.chamfer)


def multistart_icp_sharded(source, target, inits, max_correspondence_distance: float = 20.0,
                           max_iteration: int = 30, relative_fitness: float = 1e-6,
                           relative_rmse: float = 1e-6, group=None, runner: Optional[Callable] = None):
    """BASELINE configs[4] over the GPUs of one box: the
    symmetry-seeded starts are independent, so rank r runs the contiguous block
    shard_bounds(S, r, world) of `inits` (batched ICP + Chamfer score of each registered
    source, api.multistart_icp) -- no data-path collective.  The ranks then exchange one
    fixed-size record per start (4x4 pose, fitness, rmse, iterations, Chamfer: 20 doubles) with
    one all-gather, and rank by ascending Chamfer, first minimum first (icp.py:113-117 applied
    per start).  Every rank passes the full arrays and gets the same dict:
    transformations [S,4,4], fitness [S], inlier_rmse [S], iterations [S], chamfer [S], order [S],
    best (index of the first minimum; also agreed through global_first_argmin)."""
    runner = runner or _cuda_multistart
    if td.is_initialized():
        rank, world = td.get_rank(group), td.get_world_size(group)
    else:
        rank, world = 0, 1
    inits = np.asarray(inits, dtype=np.float64).reshape(-1, 4, 4)
    s = len(inits)
    lo, hi = shard_bounds(s, rank, world)
    per = (s + world - 1) // world
    rec = np.full((per, 20), np.inf)
    if hi > lo:
        Ts, fit, rmse, iters, ch = runner(source, target, inits[lo:hi], max_correspondence_distance,
                                          max_iteration, relative_fitness, relative_rmse)
        rec[: hi - lo, :16] = np.asarray(Ts, dtype=np.float64).reshape(hi - lo, 16)
        rec[: hi - lo, 16], rec[: hi - lo, 17] = fit, rmse
        rec[: hi - lo, 18], rec[: hi - lo, 19] = iters, ch
    if world > 1:
        dev = (torch.device("cuda", torch.cuda.current_device())
               if td.get_backend(group) == "nccl" else torch.device("cpu"))
        mine = torch.from_numpy(rec).to(dev)
        out = torch.empty((world * per, 20), dtype=torch.float64, device=dev)
        td.all_gather_into_tensor(out, mine, group=group)
        rec = out.cpu().numpy()
    rec = rec[:s]
    ch = rec[:, 19].copy()
    # C1 on the per-rank minima: the same first-minimum rule as the candidate selection
    lbest = int(np.argmin(ch[lo:hi])) + lo if hi > lo else 0
    lloss = torch.tensor([ch[lbest] if hi > lo else float("inf")], dtype=torch.float64)
    lidx = torch.tensor([lbest if hi > lo else _I64_MAX], dtype=torch.int64)
    if world > 1 and td.get_backend(group) == "nccl":
        lloss, lidx = lloss.cuda(), lidx.cuda()
    _, best = global_first_argmin(lloss, lidx, group)
    order = np.argsort(ch, kind="stable")
    assert int(best.item()) == int(order[0])
    return {"transformations": rec[:, :16].reshape(s, 4, 4).copy(), "fitness": rec[:, 16].copy(),
            "inlier_rmse": rec[:, 17].copy(), "iterations": rec[:, 18].astype(np.int64), "chamfer": ch,
            "order": order, "best": int(order[0])}


class PeerExchange:
    """This rank's end of the kernel-fused exchange of isr_icp_run_sharded (include/isr.h):
    a small buffer in this GPU's HBM that every peer maps through CUDA IPC and writes over
    NVLink.  The IPC handles travel once through torch.distributed (all_gather, any backend);
    the loop itself contains no collective call.  Collective: every rank of `group` must
    construct it, and `close()` it, at the same point.  Construction never raises half-way
    through its collectives: every rank takes part in all of them and the ranks agree on the
    outcome -- `ok` is the same everywhere, `error` says what went wrong on this rank (or that
    a peer failed)."""

    def __init__(self, group=None, timeout_s: Optional[float] = None):
        import ctypes

        from . import _lib

        self.group = group
        if td.is_initialized():
            self.rank, self.world = td.get_rank(group), td.get_world_size(group)
        else:
            self.rank, self.world = 0, 1
        self.handle, self.ok, self.error = None, False, None
        nbytes = _lib.ISR_PEER_HANDLE_BYTES
        mine = (ctypes.c_ubyte * nbytes)()
        try:
            if self.world > _lib.ISR_PEER_MAX_RANKS:
                raise ValueError(f"{self.world} ranks > {_lib.ISR_PEER_MAX_RANKS}")
            h = ctypes.c_void_p()
            _lib.check(_lib.load().isr_peer_create(self.rank, self.world, ctypes.byref(h), mine))
            self.handle = h
            timeout_s = timeout_s or float(os.environ.get("ISR_PEER_TIMEOUT_S", "0") or 0)
            if timeout_s:  # how long the kernel waits for a peer's sums (default ~10 s)
                _lib.check(_lib.load().isr_peer_set_timeout(h, float(timeout_s)))
        except Exception as e:  # noqa: BLE001 -- agreed with the peers below
            self.error = e
        handles = bytes(mine)
        if self.world > 1:
            # rides on whatever backend the group has; a byte tensor on the group's device
            dev = (torch.device("cuda", torch.cuda.current_device())
                   if td.get_backend(group) == "nccl" else torch.device("cpu"))
            t = torch.frombuffer(bytearray(handles), dtype=torch.uint8).to(dev)
            out = torch.empty((self.world * nbytes,), dtype=torch.uint8, device=dev)
            td.all_gather_into_tensor(out, t, group=group)
            handles = bytes(out.cpu().numpy().tobytes())
        if self._agree(self.handle is not None):
            try:
                buf = (ctypes.c_ubyte * len(handles)).from_buffer_copy(handles)
                _lib.check(_lib.load().isr_peer_connect(self.handle, buf))
                connected = True
            except Exception as e:  # noqa: BLE001
                self.error, connected = e, False
            self.ok = self._agree(connected)
        if not self.ok:
            if self.error is None:
                self.error = RuntimeError("a peer rank could not create or map the exchange buffers")
            self._release()

    def _agree(self, flag: bool) -> bool:
        """MIN over the ranks of a local success flag."""
        if self.world == 1:
            return bool(flag)
        dev = (torch.device("cuda", torch.cuda.current_device())
               if td.get_backend(self.group) == "nccl" else torch.device("cpu"))
        ok = torch.tensor([1 if flag else 0], dtype=torch.int32, device=dev)
        td.all_reduce(ok, op=td.ReduceOp.MIN, group=self.group)
        return bool(int(ok.item()))

    def _release(self) -> None:
        from . import _lib

        if self.handle is not None and self.handle.value:
            _lib.check(_lib.load().isr_peer_destroy(self.handle))
        self.handle = None

    def close(self) -> None:
        if self.handle is not None and self.handle.value:
            if self.world > 1:
                torch.cuda.synchronize()
                td.barrier(group=self.group)  # no peer may still be writing into this buffer
            self._release()
        self.ok = False


_peer_cache: dict = {}


def peer_exchange(group=None, required: bool = True) -> Optional[PeerExchange]:
    """The process-wide PeerExchange of `group` (created on first use; collective).  The ranks
    agree on the outcome: if creating or mapping the buffers fails on ANY rank (GPUs without
    peer access, IPC disabled in the container), every rank drops its end and gets None --
    or the error, when `required`."""
    key = id(group) if group is not None else None
    if key not in _peer_cache:
        _peer_cache[key] = PeerExchange(group)
    px = _peer_cache[key]
    if not px.ok:
        if required:
            raise RuntimeError(f"kernel-fused peer exchange unavailable: {px.error}")
        return None
    return px


def close_peer_exchanges() -> None:
    """Collective: release every cached PeerExchange (before destroy_process_group)."""
    for key in sorted(_peer_cache, key=lambda k: (k is not None, k)):
        _peer_cache.pop(key).close()


class CudaIcpBackend:
    """accumulate / solve of one source shard on this rank's GPU (api.IcpProblem)."""

    def __init__(self, source_shard, target, inits):
        from . import api

        self.prob = api.IcpProblem(source_shard, target, inits)

    def accumulate(self, max_dist: float) -> torch.Tensor:
        return self.prob.accumulate(max_dist)

    def solve(self, sums, ns_total, rel_fitness, rel_rmse, final_eval) -> None:
        self.prob.solve(ns_total, rel_fitness, rel_rmse, final_eval, sums)

    def results(self):
        return self.prob.results(with_correspondences=False)


class CudaIcpTargetShardBackend:
    """search / accumulate of ALL source points against one target slice (api.IcpProblem)."""

    def __init__(self, source, target_slice, inits):
        from . import api

        self.prob = api.IcpProblem(source, target_slice, inits)

    def search(self):
        """-> (local neighbour index int32 [S, ns], exact squared distance float64 [S, ns])."""
        idx = self.prob.search()
        return idx, self.prob.corr_dist(idx)

    def accumulate(self, local_idx, max_dist: float) -> torch.Tensor:
        return self.prob.accumulate_corr(local_idx, max_dist)

    def solve(self, sums, ns_total, rel_fitness, rel_rmse, final_eval) -> None:
        self.prob.solve(ns_total, rel_fitness, rel_rmse, final_eval, sums)

    def results(self):
        return self.prob.results(with_correspondences=False)


def _icp_target_sharded(source, target, inits, max_dist, max_iteration, rel_fitness, rel_rmse, group,
                        backend_factory, rank, world):
    nt, ns = len(target), len(source)
    if nt < world:
        raise ValueError(f"target-sharded ICP needs at least one target point per rank ({nt} < {world})")
    lo, hi = shard_bounds(nt, rank, world)
    be = (backend_factory or CudaIcpTargetShardBackend)(source, target[lo:hi], inits)
    for k in range(max_iteration + 1):
        idx, D = be.search()
        gidx = idx.to(torch.int64) + lo
        if world > 1:
            Dmin = D.clone()
            _all_reduce(Dmin, td.ReduceOp.MIN, group)
            # lowest global target index among the ranks that hold the minimum
            gidx = torch.where(D == Dmin, gidx, torch.full_like(gidx, _I64_MAX))
            _all_reduce(gidx, td.ReduceOp.MIN, group)
        mine = (gidx >= lo) & (gidx < hi)
        local = torch.where(mine, gidx - lo, torch.full_like(gidx, -1)).to(torch.int32).contiguous()
        sums = be.accumulate(local, max_dist)
        if world > 1:
            _all_reduce(sums, td.ReduceOp.SUM, group)
        be.solve(sums, ns, rel_fitness, rel_rmse, k == max_iteration)
    return be.results()


def icp_sharded(source, target, init=None, max_correspondence_distance: float = 20.0,
                max_iteration: int = 30, relative_fitness: float = 1e-6,
                relative_rmse: float = 1e-6, group=None, backend_factory: Optional[Callable] = None,
                shard: str = "source", exchange: str = "auto", spatial: bool = True):
    """Every rank passes the FULL source and target.  shard="source": rank r registers block
    shard_bounds(ns, r, world) of the source -- of its Hilbert-curve order when `spatial` (the
    default on the CUDA back end: a compact patch per rank), else of its rows -- against the
    whole target; per iteration the 17 sums per
    start are exchanged between accumulate and solve -- exchange="peer": inside those two
    kernels through peer memory (PeerExchange / isr_icp_run_sharded; the whole loop is one C
    call), exchange="nccl": one all-reduce per iteration, "auto": peer whenever it applies
    (CUDA back end, every shard non-empty, at most 64 starts, at most 8 ranks).
    shard="target": rank r holds target rows shard_bounds(nt, r, world) (module docstring).
    No form synchronises with the host inside the loop.  Returns this rank's list of results
    (identical on all ranks); `init` may be [4,4] or [S,4,4]."""
    if shard not in ("source", "target"):
        raise ValueError("shard must be 'source' or 'target'")
    if exchange not in ("auto", "peer", "nccl"):
        raise ValueError("exchange must be 'auto', 'peer' or 'nccl'")
    if td.is_initialized():
        rank, world = td.get_rank(group), td.get_world_size(group)
    else:
        rank, world = 0, 1
    inits = np.eye(4)[None] if init is None else np.asarray(init, dtype=np.float64).reshape(-1, 4, 4)
    if shard == "target":
        return _icp_target_sharded(source, target, inits, max_correspondence_distance, max_iteration,
                                   relative_fitness, relative_rmse, group, backend_factory, rank, world)
    ns = len(source)
    lo, hi = shard_bounds(ns, rank, world)
    from . import _lib

    if backend_factory is None and spatial and world > 1:
        # Shard along the Hilbert curve, not by row number: rank r gets the r-th contiguous run
        # of the curve-ordered source, i.e. one compact patch of the surface at full density.
        # Rows [lo, hi) of a cloud in file order (farthest-point-sampling order, genFeat.py:199-202)
        # are a `world`-times sparser sample of the WHOLE surface: every rank would then walk all
        # of the target's tiles and the search would not get faster with more ranks.  The order
        # is an integer sort of the same keys on every rank (deterministic), so the ranks agree.
        from . import api

        # (one upload of the whole source; order and gather on the device -- a float64 source keeps
        # its precision, the order only needs float32 coordinates)
        dev = torch.device("cuda", torch.cuda.current_device())
        src_dev = source.to(dev) if isinstance(source, torch.Tensor) else torch.from_numpy(
            np.ascontiguousarray(source)).to(dev)
        perm = api.spatial_order(src_dev.to(torch.float32))[lo:hi].to(torch.int64)
        source_shard = src_dev.index_select(0, perm)
    else:
        source_shard = source[lo:hi]

    # decided from (ns, world, starts) alone, hence identically on every rank
    peer_ok = (backend_factory is None and world <= _lib.ISR_PEER_MAX_RANKS
               and len(inits) <= _lib.ISR_PEER_MAX_STARTS
               and all(shard_bounds(ns, r, world)[0] < shard_bounds(ns, r, world)[1] for r in range(world)))
    if exchange == "peer" and not peer_ok:
        raise ValueError("exchange='peer' needs the CUDA back end, a non-empty shard on every rank, "
                         f"<= {_lib.ISR_PEER_MAX_STARTS} starts and <= {_lib.ISR_PEER_MAX_RANKS} ranks")
    if peer_ok and (exchange == "peer" or (exchange == "auto" and world > 1)):
        from . import api

        px = peer_exchange(group, required=(exchange == "peer"))
        if px is not None:
            prob = api.IcpProblem(source_shard, target, inits)
            # the solve side waits for its peers inside a kernel (bounded spin): start together
            if world > 1:
                torch.cuda.synchronize()
                td.barrier(group=group)
            prob.run_sharded(px, ns, max_correspondence_distance, max_iteration, relative_fitness,
                             relative_rmse)
            return prob.results(with_correspondences=False)
    factory = backend_factory or CudaIcpBackend
    be = factory(source_shard, target, inits)
    for k in range(max_iteration + 1):
        sums = be.accumulate(max_correspondence_distance)
        if world > 1:
            _all_reduce(sums, td.ReduceOp.SUM, group)
        be.solve(sums, ns, relative_fitness, relative_rmse, k == max_iteration)
    return be.results()
