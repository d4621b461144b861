"""Developer probe: where does the time of batched ADD-S go (20k vertices vs 100k surface)?"""
import ctypes, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import imagesequenceregistrationfor6dposeestimationlabeling_b200 as isr
from imagesequenceregistrationfor6dposeestimationlabeling_b200 import synth, api, _lib

torch.cuda.set_device(0)
lib = _lib.load()
cloud = synth.make_cloud(100000, 1)
verts = synth.make_cloud(20000, seed=3)
R_t, t_t = synth.true_pose(3)
for B in (1000, 8000):
    Rs_a, ts_a, _ = synth.make_candidates(B, seed=10, R_true=R_t, t_true=t_t)
    gR, gT = np.tile(R_t, (B, 1, 1)), np.tile(t_t, (B, 1))
    api.adds(verts, gR[:64], gT[:64], Rs_a[:64], ts_a[:64], cloud); torch.cuda.synchronize()
    t0 = time.perf_counter(); r = api.adds(verts, gR, gT, Rs_a, ts_a, cloud).cpu(); dt = time.perf_counter() - t0
    # device-resident inputs
    Pq = np.tile(np.eye(4), (B, 1, 1)); Pt = np.tile(np.eye(4), (B, 1, 1))
    Pq[:, :3, :3], Pq[:, :3, 3] = gR, gT; Pt[:, :3, :3], Pt[:, :3, 3] = Rs_a, ts_a
    dev = api._device()
    vd, cd, Pqd, Ptd = api._points(verts, dev), api._points(cloud, dev), api._poses(Pq, dev), api._poses(Pt, dev)
    api.verify_poses(vd, Pqd, Ptd, cloud_t=cd, mode="adds"); torch.cuda.synchronize()
    lib.isr_profile_enable(1); lib.isr_profile_collect(None, None)
    t0 = time.perf_counter(); api.verify_poses(vd, Pqd, Ptd, cloud_t=cd, mode="adds"); torch.cuda.synchronize(); dt2 = time.perf_counter() - t0
    ms = (ctypes.c_double * 5)(); ln = (ctypes.c_uint64 * 5)()
    lib.isr_profile_collect(ms, ln); lib.isr_profile_enable(0)
    print(f"B={B}: api.adds from host {B/dt:.0f} pairs/s ({dt*1e3:.1f} ms); verify_poses device-resident {B/dt2:.0f} pairs/s ({dt2*1e3:.1f} ms); "
          f"kernel ms: transform/prepare {ms[0]:.2f} ({ln[0]} launches), nn {ms[1]:.2f} ({ln[1]}), reduce {ms[2]:.2f} ({ln[2]})")
