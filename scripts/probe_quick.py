"""Developer probe (not the bench): the four timings a kernel change is judged by, in ~20 s.
    python scripts/probe_quick.py [verify] [adds] [icp] [ms]
  verify  Chamfer verification, 512 candidates x 100k points, device-resident -> candidates/s
  adds    ADD-S, 512 pose pairs x 20k vertices vs 100k surface points      -> pose pairs/s
  icp     fused ICP iteration at 1M x 1M: whole source, and the 1/8 curve shard of an 8-GPU run
  ms      config 5 (64 starts x 250k) wall clock
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from imagesequenceregistrationfor6dposeestimationlabeling_b200 import api, synth

which = sys.argv[1:] or ["verify", "adds", "icp"]
torch.cuda.set_device(0)
dev = torch.device("cuda", 0)


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e-3)
    return best


if "verify" in which:
    B = 512
    cloud = synth.make_cloud(100000, 1)
    R_true, _ = synth.true_pose(3)
    Rs, _, k0 = synth.make_candidates(1000, 10, R_true=R_true, t_true=np.zeros(3))
    Mq, Mt = synth.verification_matrices(Rs[:B], R_true)
    cd, Mqd, Mtd = api._points(cloud, dev), api._poses(Mq, dev), api._poses(Mt, dev)
    t = timed(lambda: api.verify_poses(cd, Mqd, Mtd))
    print(f"verify: {B / t:.0f} candidates/s ({t / B * 1e6:.1f} us per candidate)", flush=True)
    if "ctr" in which:  # per-warp event counts of the pruned search (isr_profile_nn_counters)
        import ctypes
        from imagesequenceregistrationfor6dposeestimationlabeling_b200 import _lib
        lib = _lib.load()
        lib.isr_profile_enable(1)
        lib.isr_profile_nn_pairs(None, None)
        api.verify_poses(cd, Mqd, Mtd)
        c = (ctypes.c_uint64 * 8)()
        _lib.check(lib.isr_profile_nn_counters(c))
        lib.isr_profile_enable(0)
        w = max(int(c[4]), 1)
        print(f"  per warp: {c[0] / w / 4:.2f} unit equivalents scanned, {c[3] / w:.1f} exact tests, {c[2] / w:.1f} "
              f"candidate stages, {c[5] / w:.1f} flagged tiles, {c[6] / w:.1f} resolve passes; slowest warp "
              f"{(int(c[7]) >> 44) * 1024} cycles", flush=True)
if "adds" in which:
    B = 512
    surface = synth.make_cloud(100000, 1)
    verts = synth.make_cloud(20000, seed=3)
    R_t, t_t = synth.true_pose(3)
    Rs_a, ts_a, _ = synth.make_candidates(B, seed=10, R_true=R_t, t_true=t_t)
    Pq = np.tile(api.pose_from_Rt(R_t, t_t), (B, 1, 1))
    Pt = np.stack([api.pose_from_Rt(Rs_a[k], ts_a[k]) for k in range(B)])
    vd, sd, Pqd, Ptd = api._points(verts, dev), api._points(surface, dev), api._poses(Pq, dev), api._poses(Pt, dev)
    t = timed(lambda: api.verify_poses(vd, Pqd, Ptd, cloud_t=sd, mode="adds"))
    print(f"adds: {B / t:.0f} pose pairs/s ({t / B * 1e6:.1f} us per pair)", flush=True)
if "icp" in which:
    n, iters = 1_000_000, 50
    src, tgt, _ = synth.icp_pair(n, n, 4, 5)
    perm = api.spatial_order(src).cpu().numpy()
    for world in (1, 8):
        shard = src[perm[:(n + world - 1) // world]]
        prob = api.IcpProblem(shard, tgt, np.eye(4)[None])
        prob.run(20.0, 1, 0.0, 0.0)
        prob.reopen()
        t = timed(lambda: (prob.reopen(), prob.run(20.0, iters - 1, 0.0, 0.0)), reps=2) / iters
        r = prob.results(False)[0]
        print(f"icp 1/{world} shard ({len(shard)} pts): {t * 1e3:.4f} ms / iteration ({1 / t:.0f} it/s) rmse {r.inlier_rmse:.5f}",
              flush=True)
        del prob
if "ms" in which:
    s5, t5, _ = synth.icp_pair(250000, 250000, 6, 7)
    inits5 = np.stack([synth.pose_matrix(synth.rotvec_to_matrix([0, 0, 2 * np.pi * k / 64]), [0, 0, 0])
                       for k in range(64)])
    api.multistart_icp(s5, t5, inits5[:2], 20.0, max_iteration=2)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ms5 = api.multistart_icp(s5, t5, inits5, 20.0, max_iteration=30)
    torch.cuda.synchronize()
    print(f"config 5: {time.perf_counter() - t0:.3f} s, best start {int(ms5.order[0])}", flush=True)
