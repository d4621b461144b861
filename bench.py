#!/usr/bin/env python
"""Bench of the registration-and-verification hot path (BASELINE.json metric:
"candidate poses verified/sec @100k pts; ICP iters/sec @1M pts; % FP32 peak").

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
                  [--metric verify|icp] [--global-candidates G] [--mix default|aligned|haar]

A step is one pass of the hot path over one batch of synthetic input: the Chamfer
verification of `--candidates` (default 1000 = BASELINE configs[1], T-LESS-shaped)
PnP-RANSAC candidate poses against a 100k-point cloud, including the best-pose selection.
With N > 1 (torchrun, one rank per GPU) every rank scores its own 1000 candidates (weak
scaling, BASELINE configs[2] shards candidates) and the ranks agree on the argmin through
NCCL inside the timed region; `--global-candidates 10000` fixes the TOTAL instead (BASELINE
configs[2] as written: a 10k sweep split over the ranks, strong scaling).  Rank 0 prints ONE
JSON line.

  value      candidates/s with all inputs resident in HBM (device-timed, max over ranks),
             library profiling OFF
  e2e        the same through the public API with pinned HOST inputs and the result read
             back to the host every step (copies inside the timed region)
  roofline   the nearest-neighbour kernel of the step (K2, nn2_pruned_kernel), from a SEPARATE
             profiled pass over the same batch (CUDA events recorded inside libisr around every
             launch + a device counter of evaluated pairs): FP32 flops of the point pairs it
             EVALUATED (8 per pair) / its duration, against the FP32 CUDA-core
             peak.  The kernel skips tiles that provably cannot hold a nearest neighbour (exact
             results, like the KD-tree of the reference), so `brute_force_equivalent` -- 8 flop
             x every pair it ANSWERED for -- is reported next to it and may exceed the peak.
             The scan executes 3 FMA (6 flop) per pair; `executed_frac` is the honest pipe load
  roofline_exhaustive  the same figures for the exhaustive kernel (nn2_kernel, every pair
             evaluated, pruning switched off) on a bounded batch of the same candidates
  cpu_baseline  the float64 CPU oracle (scipy cKDTree stand-in for Open3D) on a bounded
             sample of the same candidates, on this box's host cores
  secondary  ICP iterations/s on a 1M x 1M pair (BASELINE configs[3]; N > 1: source sharded
             along the Hilbert curve, sums exchanged inside the kernel over NVLink; result
             asserted equal to the single-GPU pose), config 5 (N > 1: starts sharded), ADD-S and
             vote throughput, candidate-mix sensitivity, K1 / K1' HBM rooflines on > 2 GB outputs

`--metric icp` prints the dense-ICP line as the headline instead (same JSON contract; strong
scaling of ONE 1M x 1M problem over the ranks).
`--impl reference` times the reference's own CPU path (its arithmetic lives in Open3D, which
cannot be installed here, so the oracle port stands in): same metric, unit and config.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "candidate poses verified/sec @100k pts"
UNIT = "candidates/s"
N_POINTS = 100_000
CLOUD_SEED, POSE_SEED, CAND_SEED = 1, 3, 10
FLOP_PER_PAIR = 8.0


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--candidates", type=int, default=1000, help="candidates per GPU per step")
    ap.add_argument("--points", type=int, default=N_POINTS)
    ap.add_argument("--icp-points", type=int, default=1_000_000)
    ap.add_argument("--icp-iters", type=int, default=50)   # BASELINE configs[3]: 50 iterations
    ap.add_argument("--skip-icp", action="store_true")
    ap.add_argument("--skip-extra", action="store_true", help="skip the config-5 / ADD-S timings")
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--metric", default="verify", choices=["verify", "icp"])
    ap.add_argument("--global-candidates", type=int, default=0,
                    help="total candidates per step, split over the ranks (strong scaling); 0 = --candidates per GPU")
    ap.add_argument("--mix", default="default", choices=["default", "aligned", "haar"])
    return ap.parse_args()


def total_candidates(args, world):
    return args.global_candidates if args.global_candidates > 0 else args.candidates * world


def make_workload(n_points, n_total, mix="default"):
    from imagesequenceregistrationfor6dposeestimationlabeling_b200 import synth

    cloud = synth.make_cloud(n_points, CLOUD_SEED)
    R_true, _ = synth.true_pose(POSE_SEED)
    Rs, _, k0 = synth.make_candidates(n_total, CAND_SEED, R_true=R_true, t_true=np.zeros(3), mix=mix)
    Mq, Mt = synth.verification_matrices(Rs, R_true)
    return cloud, Mq, Mt, k0


def config_dict(args, world):
    tot = total_candidates(args, world)
    per = (tot + world - 1) // world
    if args.global_candidates > 0:
        wl = (f"large candidate sweep (BASELINE configs[2]): {tot} PnP-RANSAC candidates x {args.points}-pt cloud "
              f"split over {world} GPU(s), bidirectional Chamfer + first-min selection")
    else:
        wl = (f"T-LESS-shaped pose verification (BASELINE configs[1]): {args.candidates} "
              f"PnP-RANSAC candidates x {args.points}-pt cloud per GPU, bidirectional Chamfer "
              "+ first-min selection")
    return {
        "workload": wl,
        "candidates_per_gpu": per,
        "points": args.points,
        "global_candidates": tot,
        "candidate_mix": {"default": "90 % within 30 deg of the true pose, 10 % Haar-random",
                          "aligned": "all within 30 deg of the true pose",
                          "haar": "all Haar-random"}[args.mix],
        "parallelism": f"candidate-sharded x{world}" if world > 1 else "single GPU",
        "l2_policy": "inputs larger than L2: each step streams 2 x candidates x 1.2 MB of "
                     "transformed clouds (2.4 GB at 1000 candidates) through a 126 MB L2",
    }


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thr = threading.Thread(target=self._pump, daemon=True)
            self.thr.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(max(mx)) if mx else None,
                "power_w_max": float(max(pw)) if pw else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def cpu_verify_sample(cloud, Mq, Mt, seconds_target=15.0, max_n=256):
    """Oracle (float64, scipy cKDTree with every host thread) on the first n candidates."""
    from oracle import oracle

    t0 = time.perf_counter()
    oracle.verify_matrices(cloud, cloud, Mq[:2], Mt[:2], bidirectional=True)
    per = (time.perf_counter() - t0) / 2
    n = int(min(max_n, max(4, seconds_target / max(per, 1e-3)), len(Mq)))
    t0 = time.perf_counter()
    losses, best = oracle.verify_matrices(cloud, cloud, Mq[:n], Mt[:n], bidirectional=True)
    dt = time.perf_counter() - t0
    return n, dt, losses


def run_reference(args):
    """The reference's CPU implementation of the path (oracle port: Open3D not installable)."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    from oracle import oracle

    cloud, Mq, Mt, k0 = make_workload(args.points, args.candidates, args.mix)
    cores = len(os.sched_getaffinity(0))
    t0 = time.perf_counter()
    oracle.verify_matrices(cloud, cloud, Mq[:2], Mt[:2], bidirectional=True)
    per = (time.perf_counter() - t0) / 2
    total = args.steps + args.warmup
    s = int(min(64, max(2, 90.0 / max(per * total, 1e-3))))
    times = []
    for it in range(total):
        lo = (it * s) % max(1, len(Mq) - s)
        t0 = time.perf_counter()
        oracle.verify_matrices(cloud, cloud, Mq[lo:lo + s], Mt[lo:lo + s], bidirectional=True)
        dt = time.perf_counter() - t0
        if it >= args.warmup:
            times.append(dt)
    tot = float(np.sum(times))
    value = s * len(times) / tot
    sample = (f"{s} of the {args.candidates} candidates per step (float64 KD-tree Chamfer, both "
              f"directions, {args.points} pts), scipy cKDTree workers=-1 as the Open3D stand-in")
    emit({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot / len(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": config_dict(args, world),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    })


def emit(obj):
    """Print the ONE JSON line on the real stdout (see main: fd 1 is parked on stderr while
    the run is in flight so that library chatter -- NCCL's version banner -- cannot mix in)."""
    sys.stdout.flush()
    if _REAL_STDOUT is not None:
        os.dup2(_REAL_STDOUT, 1)
    print(json.dumps(obj), flush=True)


_REAL_STDOUT = None


def main():
    global _REAL_STDOUT
    args = parse_args()
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import torch.distributed as td

    import imagesequenceregistrationfor6dposeestimationlabeling_b200 as isr
    from imagesequenceregistrationfor6dposeestimationlabeling_b200 import _lib, api, dist, helpers, synth

    rank, world = dist.init_from_env()
    if world != args.gpus and rank == 0:
        print(f"[bench] note: --gpus {args.gpus} but WORLD_SIZE={world}; using {world}", file=sys.stderr)
    dev = torch.device("cuda", torch.cuda.current_device())
    lib = _lib.load()

    n_total = total_candidates(args, world)
    cloud, Mq, Mt, k0 = make_workload(args.points, n_total, args.mix)
    lo, hi = dist.shard_bounds(len(Mq), rank, world)
    b_local = hi - lo

    def barrier():
        if world > 1:
            td.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        td.all_reduce(t, op=td.ReduceOp.MAX)
        return float(t.item())

    def nn_pairs():
        ev, an = ctypes.c_uint64(0), ctypes.c_uint64(0)
        _lib.check(lib.isr_profile_nn_pairs(ctypes.byref(ev), ctypes.byref(an)))
        return float(ev.value), float(an.value)

    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:
        pass
    hbm = float(peaks.get("hbm_gbs", 6650.0))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    # ---------------- device-resident arm: `value` (library profiling OFF) --------------------
    cloud_d = api._points(cloud, dev)
    Mq_d = api._poses(Mq[lo:hi], dev)
    Mt_d = api._poses(Mt[lo:hi], dev)
    off = torch.tensor([lo], dtype=torch.int64, device=dev)

    def step_resident():
        res = api.verify_poses(cloud_d, Mq_d, Mt_d, mode="chamfer")
        return dist.global_first_argmin(res.best[1:2].view(torch.float64), res.best[0:1] + off)

    lib.isr_profile_enable(0)
    for _ in range(args.warmup):
        step_resident()
    barrier()
    sampler = ClockSampler(torch.cuda.current_device())
    if rank == 0:
        sampler.start()
    _lib.reset_launch_count()
    barrier()
    e0.record()
    for _ in range(args.steps):
        best_loss, best_idx = step_resident()
    e1.record()
    barrier()
    t_resident = max_over_ranks(e0.elapsed_time(e1) * 1e-3)
    launches = _lib.launch_count()
    clocks = sampler.stop() if rank == 0 else None
    sel_idx, sel_loss = int(best_idx.item()), float(best_loss.item())
    value = n_total * args.steps / t_resident

    # ---------------- the same step again, profiled: kernel time + evaluated pairs -----------------
    prof_steps = max(1, min(args.steps, 5))
    lib.isr_profile_enable(1)
    lib.isr_profile_collect(None, None)
    nn_pairs()
    e0.record()
    for _ in range(prof_steps):
        step_resident()
    e1.record()
    torch.cuda.synchronize()
    t_prof = e0.elapsed_time(e1) * 1e-3
    ms_kind = (ctypes.c_double * 5)()
    n_kind = (ctypes.c_uint64 * 5)()
    _lib.check(lib.isr_profile_collect(ms_kind, n_kind))
    pairs_evaluated, pairs_answered = nn_pairs()
    lib.isr_profile_enable(0)
    nn_ms, nn_launches = float(ms_kind[1]), int(n_kind[1])
    prep_ms = float(ms_kind[0])
    pairs_total = 2.0 * args.points * args.points * b_local * prof_steps  # both directions
    assert abs(pairs_answered - pairs_total) <= 1e-9 * pairs_total, (pairs_answered, pairs_total)
    achieved = FLOP_PER_PAIR * pairs_evaluated / (nn_ms * 1e-3) / 1e12 if nn_ms > 0 else None
    bf_equiv = FLOP_PER_PAIR * pairs_total / (nn_ms * 1e-3) / 1e12 if nn_ms > 0 else None

    # ---------------- the exhaustive kernel on a bounded batch (pruning off) ---------------
    n_ex = min(b_local, 250)
    api.set_nn_pruning(False)
    try:
        api.verify_poses(cloud_d, Mq_d[:min(n_ex, 16)], Mt_d[:min(n_ex, 16)], mode="chamfer")
        torch.cuda.synchronize()
        lib.isr_profile_enable(1)
        lib.isr_profile_collect(None, None)
        nn_pairs()
        ex_res = api.verify_poses(cloud_d, Mq_d[:n_ex], Mt_d[:n_ex], mode="chamfer")
        torch.cuda.synchronize()
        ex_ms, ex_n = (ctypes.c_double * 5)(), (ctypes.c_uint64 * 5)()
        _lib.check(lib.isr_profile_collect(ex_ms, ex_n))
        ex_eval, ex_answered = nn_pairs()
        lib.isr_profile_enable(0)
    finally:
        api.set_nn_pruning(True)
    pr_res = api.verify_poses(cloud_d, Mq_d[:n_ex], Mt_d[:n_ex], mode="chamfer")
    assert torch.equal(ex_res.losses, pr_res.losses), "pruned and exhaustive losses differ"
    ex_tflops = FLOP_PER_PAIR * ex_answered / (float(ex_ms[1]) * 1e-3) / 1e12

    # ---------------- end-to-end arm: host buffers in, host result out -------------------
    cloud_h = torch.from_numpy(cloud).pin_memory()
    Mq_h = torch.from_numpy(np.ascontiguousarray(Mq[lo:hi])).pin_memory()
    Mt_h = torch.from_numpy(np.ascontiguousarray(Mt[lo:hi])).pin_memory()
    losses_h = torch.empty((b_local,), dtype=torch.float64).pin_memory()
    h2d = cloud_h.numel() * 4 + Mq_h.numel() * 8 + Mt_h.numel() * 8
    d2h = losses_h.numel() * 8 + 16

    def step_e2e():
        res = api.verify_poses(cloud_h, Mq_h, Mt_h, mode="chamfer")  # H2D inside
        bl, bi = dist.global_first_argmin(res.best[1:2].view(torch.float64), res.best[0:1] + off)
        losses_h.copy_(res.losses, non_blocking=True)
        return int(bi.item()), float(bl.item())  # D2H + sync

    step_e2e()
    barrier()
    t0 = time.perf_counter()
    e0.record()
    for _ in range(args.steps):
        e2e_idx, e2e_loss = step_e2e()
    e1.record()
    barrier()
    t_e2e = max_over_ranks(max(e0.elapsed_time(e1) * 1e-3, 0.0))
    t_e2e_wall = max_over_ranks(time.perf_counter() - t0)
    t_e2e = max(t_e2e, t_e2e_wall)
    e2e_value = n_total * args.steps / t_e2e
    assert e2e_idx == sel_idx, (e2e_idx, sel_idx)

    # ---------------- FP32 peak (nominal + live FFMA chain) ------------------------------
    sm_count = ctypes.c_int(0)
    clock_khz = ctypes.c_int(0)
    _lib.check(lib.isr_device_info(ctypes.byref(sm_count), ctypes.byref(clock_khz), None))
    sm_max_mhz = float(peaks.get("sm_max_mhz", clock_khz.value / 1e3))
    peak_nominal = sm_count.value * 128 * 2 * sm_max_mhz * 1e6 / 1e12
    ffma_measured = api.measure_fp32_peak(packed=False)

    def timed(fn, reps=3):
        """Best-of device time of fn (CUDA events on the current stream), after one warm-up."""
        fn()
        torch.cuda.synchronize()
        best = 1e30
        for _ in range(reps):
            e0.record()
            fn()
            e1.record()
            e1.synchronize()
            best = min(best, e0.elapsed_time(e1) * 1e-3)
        return best

    # ---------------- secondary: dense ICP on a 1M x 1M pair (BASELINE configs[3]) ----------------
    secondary = {}
    icp_line = None
    if not args.skip_icp:
        src, tgt, _ = synth.icp_pair(args.icp_points, args.icp_points, 4, 5)
        ns_all = len(src)
        iters = args.icp_iters
        # The product's loop: ONE kernel launch per iteration (search + correspondence sums + exchange
        # + Kabsch fused, nn2.cu / icp_device.cuh).  N = 1: isr_icp_run on the whole source.  N > 1:
        # isr_icp_run_sharded on this rank's run of the Hilbert-ordered source (a compact patch), the
        # 17 sums exchanged inside the kernel through peer memory over NVLink -- what
        # dist.icp_sharded does; no NCCL call, no host work per iteration.
        peer = dist.peer_exchange(required=False) if world > 1 else None
        if world > 1:
            sperm = api.spatial_order(src).cpu().numpy()
            slo, shi = dist.shard_bounds(ns_all, rank, world)
            shard = src[sperm[slo:shi]]
        else:
            shard = src
        pprob = api.IcpProblem(shard, tgt, np.eye(4)[None])

        def icp_run(n_iter):
            if peer is not None:
                pprob.run_sharded(peer, ns_all, 20.0, n_iter - 1, 0.0, 0.0)
            elif world == 1:
                pprob.run(20.0, n_iter - 1, 0.0, 0.0)
            else:  # peer memory unavailable on this box: one NCCL all-reduce per iteration
                for k in range(n_iter):
                    sums_k = pprob.accumulate(20.0)
                    td.all_reduce(sums_k, op=td.ReduceOp.SUM)
                    pprob.solve(ns_all, 0.0, 0.0, k == n_iter - 1, sums_k)

        icp_run(2)       # warm-up: two evaluations (hints warm)
        pprob.reopen()
        barrier()
        e0.record()      # timed with the library's profiling OFF (no event pair around every launch,
        icp_run(iters)   # dependent launches between the iterations), like the headline metric
        e1.record()
        barrier()
        t_icp = max_over_ranks(e0.elapsed_time(e1) * 1e-3)
        rp = pprob.results(with_correspondences=False)[0]
        # separate profiled pass (kernel time, evaluated pairs): the same number of evaluations,
        # continuing from the state the timed run ended in
        pprob.reopen()
        barrier()
        lib.isr_profile_enable(1)
        lib.isr_profile_collect(None, None)
        nn_pairs()
        icp_run(iters)
        barrier()
        ms_icp, n_icp = (ctypes.c_double * 5)(), (ctypes.c_uint64 * 5)()
        _lib.check(lib.isr_profile_collect(ms_icp, n_icp))
        icp_eval, icp_answered = nn_pairs()
        lib.isr_profile_enable(0)
        # parity of the sharded loop: rank 0 repeats the same 2 + iters evaluations on ONE GPU and
        # every rank compares its (identical) pose with it
        parity = None
        if world > 1:
            ref = torch.zeros((18,), dtype=torch.float64, device=dev)
            if rank == 0:
                sprob = api.IcpProblem(src, tgt, np.eye(4)[None])
                sprob.run(20.0, 1, 0.0, 0.0)
                sprob.reopen()
                sprob.run(20.0, iters - 1, 0.0, 0.0)
                rs = sprob.results(with_correspondences=False)[0]
                ref[:16] = torch.from_numpy(rs.transformation.reshape(16)).to(dev)
                ref[16], ref[17] = rs.fitness, rs.inlier_rmse
                del sprob
            td.broadcast(ref, 0)
            refh = ref.cpu().numpy()
            parity = float(np.max(np.abs(rp.transformation.reshape(16) - refh[:16])))
            assert parity < 1e-9 and abs(rp.fitness - refh[16]) < 1e-12 and \
                abs(rp.inlier_rmse - refh[17]) < 1e-9 * max(refh[17], 1.0), \
                ("sharded ICP differs from the single-GPU result", parity, rp.fitness, refh[16], rp.inlier_rmse, refh[17])
        # end to end through the public API from host arrays (H2D, curve sort, preparation, loop,
        # result read back): what a caller of api.icp / dist.icp_sharded sees
        def icp_e2e():
            if world > 1:
                return dist.icp_sharded(src, tgt, np.eye(4), 20.0, max_iteration=iters - 1, relative_fitness=0.0,
                                        relative_rmse=0.0)[0]
            return api.icp(src, tgt, np.eye(4), 20.0, max_iteration=iters - 1, relative_fitness=0.0, relative_rmse=0.0)

        # (one untimed call first: the timed one then finds its buffers in torch's caching allocator,
        # as the second and every later registration of a pipeline does; a cold call pays ~30 ms of cudaMalloc)
        icp_e2e()
        torch.cuda.synchronize()
        barrier()
        t0 = time.perf_counter()
        r_e2e = icp_e2e()
        torch.cuda.synchronize()
        t_icp_e2e = max_over_ranks(time.perf_counter() - t0)
        del pprob
        # the same loop with the exhaustive search (N = 1, one iteration timed)
        t_icp_ex = None
        if world == 1:
            eprob = api.IcpProblem(src, tgt, np.eye(4)[None])
            api.set_nn_pruning(False)
            try:
                eprob.run(20.0, 0, 0.0, 0.0)
                eprob.reopen()
                torch.cuda.synchronize()
                e0.record()
                eprob.run(20.0, 0, 0.0, 0.0)
                e1.record()
                e1.synchronize()
                t_icp_ex = e0.elapsed_time(e1) * 1e-3
            finally:
                api.set_nn_pruning(True)
            del eprob
        # the north star's target-sharded form of the same loop (N > 1 only): every rank
        # searches its slice of the target for ALL source points, two per-point MIN
        # all-reduces pick the neighbour, then the same 17-double SUM all-reduce
        t_icp_tgt = None
        if world > 1:
            tlo, thi = dist.shard_bounds(len(tgt), rank, world)
            tbe = dist.CudaIcpTargetShardBackend(src, tgt[tlo:thi], np.eye(4)[None])

            def icp_iter_target(final=False):
                idx, D = tbe.search()
                gidx = idx.to(torch.int64) + tlo
                Dmin = D.clone()
                td.all_reduce(Dmin, op=td.ReduceOp.MIN)
                gidx = torch.where(D == Dmin, gidx, torch.full_like(gidx, np.iinfo(np.int64).max))
                td.all_reduce(gidx, op=td.ReduceOp.MIN)
                mine = (gidx >= tlo) & (gidx < thi)
                local = torch.where(mine, gidx - tlo, torch.full_like(gidx, -1)).to(torch.int32).contiguous()
                sums = tbe.accumulate(local, 20.0)
                td.all_reduce(sums, op=td.ReduceOp.SUM)
                tbe.solve(sums, ns_all, 0.0, 0.0, final)

            icp_iter_target()
            icp_iter_target(final=True)     # (the same 2 + iters evaluations as the loop above)
            tbe.prob.reopen()
            barrier()
            e0.record()
            for k in range(iters):
                icp_iter_target(final=(k == iters - 1))
            e1.record()
            barrier()
            t_icp_tgt = max_over_ranks(e0.elapsed_time(e1) * 1e-3)
            rt = tbe.results()[0]
            assert abs(rt.fitness - rp.fitness) < 1e-12 and float(np.max(np.abs(rt.transformation - rp.transformation))) < 1e-9, \
                ("target-sharded ICP differs from the source-sharded result", rt.transformation, rp.transformation)
            del tbe
        secondary.update({
            "icp_iters_per_s": iters / t_icp,
            "icp_e2e_iters_per_s": iters / t_icp_e2e,
            "icp_config": f"dense ICP refine (BASELINE configs[3]): {args.icp_points} x {args.icp_points} "
                          f"points, {iters} forced iterations, one fused kernel launch per iteration; "
                          + ("isr_icp_run" if world == 1 else
                             f"source sharded x{world} along the Hilbert curve, sums exchanged inside the kernel through "
                             "peer memory over NVLink (isr_icp_run_sharded), no NCCL call" if peer is not None else
                             f"source sharded x{world}, peer memory unavailable: NCCL all-reduce per iteration"),
            "icp_e2e_config": "api.icp / dist.icp_sharded from host arrays: H2D, curve sort, preparation, loop, result "
                              "read back (wall clock of the second of two identical calls: buffers come from "
                              "torch's caching allocator)",
            "icp_ms_per_iter": 1e3 * t_icp / iters,
            "icp_kernel_ms_per_iter": float(ms_icp[1]) / max(int(n_icp[1]), 1),
            "icp_sharded_max_abs_pose_diff_vs_single_gpu": parity,
            "icp_nn_pairs_evaluated_frac": icp_eval / max(icp_answered, 1.0),
            "icp_nn_tflops_brute_force_equivalent": FLOP_PER_PAIR * icp_answered / (float(ms_icp[1]) * 1e-3) / 1e12,
            "icp_target_sharded_iters_per_s": (iters / t_icp_tgt) if t_icp_tgt else None,
            "icp_exhaustive_iters_per_s": (1.0 / t_icp_ex) if t_icp_ex else None,
            "icp_exhaustive_nn_tflops": (FLOP_PER_PAIR * ns_all * args.icp_points / t_icp_ex / 1e12) if t_icp_ex else None,
            "icp_fitness": rp.fitness, "icp_inlier_rmse": rp.inlier_rmse,
        })
        icp_line = {"value": iters / t_icp, "e2e": iters / t_icp_e2e, "ms": 1e3 * t_icp / iters}

    if not args.skip_extra:
        # ---- K1 / K1' on their own, on outputs of more than 2 GB (16 x the L2: the part of the
        # output that is still in L2 when the kernel ends is < 5 %) -- HBM roofline ------------------
        nb1 = max(64, int(2.2e9 / (args.points * 12)))
        P1 = api._poses(np.tile(np.eye(4), (nb1, 1, 1)), dev)
        P1[:, :3, :3] = api._poses(Mq[lo:lo + 1], dev)[0, :3, :3]

        def kernel_seconds(fn, kind, reps=3):
            """Device time of the kernel(s) of one profile kind inside fn (CUDA events recorded
            by libisr around the launch), best of `reps`."""
            fn()
            torch.cuda.synchronize()
            best = 1e30
            mk, nk = (ctypes.c_double * 5)(), (ctypes.c_uint64 * 5)()
            for _ in range(reps):
                lib.isr_profile_enable(1)
                lib.isr_profile_collect(None, None)
                fn()
                _lib.check(lib.isr_profile_collect(mk, nk))
                lib.isr_profile_enable(0)
                best = min(best, float(mk[kind]) * 1e-3)
            return best

        out1 = torch.empty((nb1, args.points, 3), dtype=torch.float32, device=dev)

        def k1():
            for b0 in range(0, nb1, 65535):
                bc = min(65535, nb1 - b0)
                _lib.check(lib.isr_transform_points(cloud_d.data_ptr(), args.points, P1[b0:].data_ptr(), bc,
                                                    out1[b0:].data_ptr(), torch.cuda.current_stream().cuda_stream))

        t_k1 = kernel_seconds(k1, 0)
        k1_bytes = args.points * 12 + nb1 * args.points * 12
        del out1
        npad = _lib.soa_padded_len(args.points)
        nb7 = max(32, int(2.2e9 / (7 * npad * 4)))
        cen = api.centroid_of(cloud_d)
        out7 = torch.empty((nb7, 7, npad), dtype=torch.float32, device=dev)

        def k1p():
            _lib.check(lib.isr_prepare_cloud(cloud_d.data_ptr(), None, None, args.points, P1.data_ptr(), 16,
                                             P1.data_ptr(), 16, cen.data_ptr(), nb7, out7.data_ptr(), npad, None, 0,
                                             torch.cuda.current_stream().cuda_stream))

        t_k1p = kernel_seconds(k1p, 0)
        k1p_bytes = args.points * 12 + nb7 * 7 * npad * 4
        del out7
        secondary.update({
            "k1_transform_gbs": k1_bytes / t_k1 / 1e9,
            "k1_transform_frac_of_hbm": k1_bytes / t_k1 / 1e9 / hbm,
            "k1_config": f"isr_transform_points: {nb1} poses x {args.points} pts, {k1_bytes / 1e9:.2f} GB written "
                         "(17 x the L2, so the bytes counted are DRAM bytes to within 5 %)",
            "k1_prepare_gbs": k1p_bytes / t_k1p / 1e9,
            "k1_prepare_frac_of_hbm": k1p_bytes / t_k1p / 1e9 / hbm,
            "k1_prepare_config": f"isr_prepare_cloud (7 planes): {nb7} poses x {args.points} pts, {k1p_bytes / 1e9:.2f} GB written",
            "hbm_peak_gbs": hbm,
        })
        # ---- sensitivity of the pruned search to the candidate mix (alignment decides how much it prunes)
        mixes = {}
        for mix in ("aligned", "haar"):
            _, Mq_m, Mt_m, _ = make_workload(args.points, min(args.candidates, 500), mix)
            Mq_md, Mt_md = api._poses(Mq_m, dev), api._poses(Mt_m, dev)
            tm = timed(lambda: api.verify_poses(cloud_d, Mq_md, Mt_md, mode="chamfer"), reps=2)
            mixes[mix + "_candidates_per_s"] = len(Mq_m) / tm
        mixes["note"] = ("device-resident, one GPU, this rank: every candidate within 30 deg of the true pose / every "
                         "candidate Haar-random; the headline mix is 90 % / 10 %")
        secondary["candidate_mix"] = mixes
        # ---- BASELINE configs[4]: 64 symmetry-seeded starts x 250k points, 30-iteration ICP each
        # (default criteria), Chamfer-ranked -- end to end from host arrays; N > 1: starts sharded
        s5, t5, _ = synth.icp_pair(250000, 250000, 6, 7)
        inits5 = np.stack([synth.pose_matrix(synth.rotvec_to_matrix([0, 0, 2 * np.pi * k / 64]), [0, 0, 0])
                           for k in range(64)])
        api.multistart_icp(s5, t5, inits5[:2], 20.0, max_iteration=2)   # warm-up
        barrier()
        t0 = time.perf_counter()
        if world > 1:
            ms5 = dist.multistart_icp_sharded(s5, t5, inits5, 20.0, max_iteration=30)
            it5 = int(np.sum(ms5["iterations"] + 1))
            best5 = int(ms5["best"])
        else:
            m5 = api.multistart_icp(s5, t5, inits5, 20.0, max_iteration=30)
            it5 = int(sum(r_.iterations + 1 for r_ in m5.results))
            best5 = int(m5.order[0])
        torch.cuda.synchronize()
        dt5 = max_over_ranks(time.perf_counter() - t0)
        secondary.update({
            "config5_multistart_seconds": dt5,
            "config5_multistart_evaluations_per_s": it5 / dt5,
            "config5_config": "64 starts x 250000 x 250000 points, <= 30 iterations each (default criteria), then "
                              f"Chamfer ranking; {it5} ICP evaluations in total; wall clock incl. host prep and H2D"
                              + (f"; starts sharded x{world} (dist.multistart_icp_sharded), one all-gather of 20 doubles per start"
                                 if world > 1 else ""),
            "config5_best_start": best5,
        })
        if world == 1:
            # ADD-S scoring as choosePose.py:20-22,124-134 calls it: 20k CAD vertices against the
            # 100k-point surface cloud, one pose pair per call, batched
            verts = synth.make_cloud(20000, seed=3)
            R_t, t_t = synth.true_pose(3)
            nb = 2000
            Rs_a, ts_a, _ = synth.make_candidates(nb, seed=10, R_true=R_t, t_true=t_t)
            gR, gT = np.tile(R_t, (nb, 1, 1)), np.tile(t_t, (nb, 1))
            api.adds(verts, gR[:64], gT[:64], Rs_a[:64], ts_a[:64], cloud)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            la = api.adds(verts, gR, gT, Rs_a, ts_a, cloud).cpu().numpy()
            dta = time.perf_counter() - t0
            Pg = np.stack([synth.pose_matrix(gR[k], gT[k]) for k in range(nb)])
            Pp = np.stack([synth.pose_matrix(Rs_a[k], ts_a[k]) for k in range(nb)])
            api.adds_rigid(verts, Pg[:64], Pp[:64], cloud)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            lr = api.adds_rigid(verts, Pg, Pp, cloud).losses.cpu().numpy()
            dtr = time.perf_counter() - t0
            assert float(np.max(np.abs(lr - la) / la)) < 1e-6, "rigid ADD-S differs from the two-cloud form"
            # the whole vote of choosePose.py:98-151 for n images (n^2 pose pairs), pose lists in, chosen image out
            nv = 192
            rngv = np.random.default_rng(4)
            gRv = np.stack([synth.random_rotation(rngv) for _ in range(nv)])
            gtv = rngv.normal(scale=20.0, size=(nv, 3)) + [0, 0, 700.0]
            # predictions: 80 % within a few degrees of the truth, 20 % failed (Haar-random)
            pRv = np.stack([gRv[k] @ synth.rotvec_to_matrix(rngv.normal(scale=0.05, size=3)) if rngv.random() < 0.8
                            else synth.random_rotation(rngv) for k in range(nv)])
            ptv = gtv + rngv.normal(scale=1.5, size=(nv, 3))
            helpers.choose_image_from_poses(pRv[:16], ptv[:16], gRv[:16], gtv[:16], verts, 120.0, surface_points=cloud)
            torch.cuda.synchronize()
            vst = {}
            t0 = time.perf_counter()
            _, img, _ = helpers.choose_image_from_poses(pRv, ptv, gRv, gtv, verts, 120.0, surface_points=cloud, stats=vst)
            dtv = time.perf_counter() - t0
            t0 = time.perf_counter()
            _, img2, _ = helpers.choose_image_from_poses(pRv[:96], ptv[:96], gRv[:96], gtv[:96], verts, 120.0,
                                                         surface_points=cloud, use_bounds=False)
            dtx = time.perf_counter() - t0
            secondary.update({
                "adds_pose_pairs_per_s": nb / dtr,
                "adds_config": f"ADD-S (one-directional): {nb} pose pairs x 20000 vertices vs 100000 surface points, host arrays "
                               "in, host losses out; surface prepared once (api.adds_rigid = the path of choose_image)",
                "adds_two_cloud_pose_pairs_per_s": nb / dta,
                "adds_two_cloud_config": "the same pairs through api.adds (both clouds transformed per pair; any 4x4)",
                "vote_pose_pairs_per_s": nv * nv / dtv,
                "vote_config": f"choosePose.py:98-151 for {nv} images = {nv * nv} pose pairs: relative-pose tables, "
                               "sphere bounds on every pair's ADD-S (isr_adds_bounds), exact ADD-S for the "
                               f"{vst.get('exact')} pairs the bounds leave undecided, 0.1 x diameter test, row sums and "
                               "argmax on the device; pose lists in, chosen image out "
                               f"(wall clock; 1280 images = 1.64 M pairs would take {1280 * 1280 / (nv * nv / dtv):.1f} s)",
                "vote_exact_only_pose_pairs_per_s": 96 * 96 / dtx,
                "vote_exact_only_config": "the same vote for 96 images with every pair scored exactly (use_bounds=False)",
            })

    # ---------------- CPU baseline (rank 0, N == 1 only) ---------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.skip_cpu:
        n_cpu, dt_cpu, cpu_losses = cpu_verify_sample(cloud, Mq, Mt)
        res = api.verify_poses(cloud_d, api._poses(Mq[:n_cpu], dev), api._poses(Mt[:n_cpu], dev))
        gl = res.losses.cpu().numpy()
        rel = float(np.max(np.abs(gl - cpu_losses) / np.abs(cpu_losses)))
        assert rel < 1e-5, f"GPU losses differ from the oracle: {rel}"
        assert res.best_index == int(np.argmin(cpu_losses))
        cpu = {"value": n_cpu / dt_cpu, "unit": UNIT, "cores": len(os.sched_getaffinity(0)),
               "kind": "port",
               "sample": f"first {n_cpu} of the {n_total} candidates, float64 scipy cKDTree "
                         f"(workers=-1) Chamfer; GPU losses on the same sample agree to {rel:.1e} rel"}

    if world > 1:
        barrier()
        dist.close_peer_exchanges()
        td.destroy_process_group()
    if rank != 0:
        return

    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "nn_kernel_traffic.json")) as f:
            per_pair = json.load(f).get("pruned_dram_bytes_per_cloud_pair")
        # ncu capture held 64 cloud pairs per launch; a bench launch holds this many
        traffic = per_pair * (2.0 * b_local * prof_steps / max(nn_launches, 1))
    except Exception:
        pass
    per_launch_flop = FLOP_PER_PAIR * pairs_evaluated / max(nn_launches, 1)
    strong = args.global_candidates > 0
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t_resident / args.steps,
        "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": config_dict(args, world),
        "selected_candidate": sel_idx, "planted_candidate": k0, "selected_loss": sel_loss,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {
            "kernel": "nn2_pruned_kernel (K2 tiled brute-force nearest neighbour: FP32 3-FMA filter scan "
                      "of every 256-query x 64-target tile pair that can hold a neighbour + FP64 resolve "
                      "inside the error window; tiles ruled out by their bounding spheres are skipped)",
            "bound": "fp32", "achieved": achieved, "peak": peak_nominal, "unit": "TFLOP/s",
            "frac": achieved / peak_nominal if achieved else None,
            "achieved_definition": "8 flop x point pairs evaluated (device counter) / kernel time, from a separate "
                                   f"profiled pass of {prof_steps} step(s) over the same batch (the timed region runs "
                                   "with profiling off)",
            "pairs_evaluated_frac": pairs_evaluated / pairs_total,
            "brute_force_equivalent": bf_equiv,
            "brute_force_equivalent_frac": bf_equiv / peak_nominal if bf_equiv else None,
            "peak_source": f"nominal {sm_count.value} SM x 128 lanes x 2 flop x {sm_max_mhz:.0f} MHz "
                           "(MEASURED_PEAKS.json holds no FP32 CUDA-core figure)",
            "peak_measured_ffma": ffma_measured,
            "frac_of_measured_ffma": achieved / ffma_measured if achieved else None,
            "executed_flop_per_pair": 6.0,
            "executed_frac": (achieved * 6.0 / 8.0) / peak_nominal if achieved else None,
            "algorithmic_ceiling_frac": 8.0 / 6.0,
            "flop_per_launch": per_launch_flop, "launches": nn_launches,
            "avg_launch_ms": nn_ms / max(nn_launches, 1),
            "kernel_share_of_step": nn_ms * 1e-3 / t_prof,
            "prepare_share_of_step": prep_ms * 1e-3 / t_prof,
            "profiled_step_ms": 1e3 * t_prof / prof_steps,
            "traffic": traffic,
        },
        "roofline_exhaustive": {
            "kernel": "nn2_kernel (pruning off: FP32 3-FMA filter scan over EVERY pair + FP64 resolve)",
            "bound": "fp32", "achieved": ex_tflops, "peak": peak_nominal, "unit": "TFLOP/s",
            "frac": ex_tflops / peak_nominal,
            "frac_of_measured_ffma": ex_tflops / ffma_measured,
            "executed_frac": ex_tflops * 6.0 / 8.0 / peak_nominal,
            "candidates": n_ex, "candidates_per_s": n_ex / (float(ex_ms[1]) * 1e-3),
            "launches": int(ex_n[1]), "avg_launch_ms": float(ex_ms[1]) / max(int(ex_n[1]), 1),
            "pairs_evaluated_frac": ex_eval / max(ex_answered, 1.0),
            "losses_identical_to_pruned": True,
        },
        "cpu_baseline": cpu,
        "secondary": secondary,
    }
    if args.metric == "icp" and icp_line is not None:
        # the dense-ICP line as the headline: ONE 1M x 1M problem over the ranks (strong scaling)
        verify_summary = {k: out[k] for k in ("value", "unit", "ms_per_step", "e2e", "roofline")}
        out.update({
            "metric": "ICP iters/sec @1M pts", "value": icp_line["value"], "unit": "iterations/s",
            "ms_per_step": icp_line["ms"], "scaling": "strong", "dtype": "f64",
            "config": {"workload": secondary["icp_config"], "points": args.icp_points, "iterations": args.icp_iters,
                       "l2_policy": "inputs larger than L2 at N = 1 (source + target planes and rows: 84 MB of a "
                                    "126 MB L2 are touched per iteration; the pose changes every iteration)"},
            "e2e": {"value": icp_line["e2e"], "unit": "iterations/s", "h2d_bytes_per_step": 2 * args.icp_points * 12 // max(args.icp_iters, 1),
                    "d2h_bytes_per_step": 184 // max(args.icp_iters, 1)},
            "verify": verify_summary,
        })
        out["roofline"] = dict(out["roofline"], note="roofline of the verification kernel; the ICP iteration is latency-"
                               "bound (secondary.icp_kernel_ms_per_iter, DESIGN.md section 3)")
    emit(out)


if __name__ == "__main__":
    main()
