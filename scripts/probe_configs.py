"""Developer probe: wall/device time of BASELINE configs 1 and 5 shaped runs through the public API."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import imagesequenceregistrationfor6dposeestimationlabeling_b200 as isr
from imagesequenceregistrationfor6dposeestimationlabeling_b200 import synth, api

torch.cuda.set_device(0)
def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize(); best = min(best, time.perf_counter() - t0)
    return best, r
# config 1: 100k x 100k, 30 iterations + final Chamfer against a 50k "CAD"
src, tgt, _ = synth.icp_pair(100000, 100000, 1, 2)
srcd, tgtd = api._points(src, api._device()), api._points(tgt, api._device())
t, r = timed(lambda: isr.icp(srcd, tgtd, np.eye(4), 20.0, max_iteration=30, relative_fitness=0.0, relative_rmse=0.0))
print(f"config1 icp 100k x 100k, 30 forced iterations (device inputs): {t*1e3:.2f} ms total, {t/31*1e3:.3f} ms per evaluation; iters {r.iterations} fitness {r.fitness:.4f}")
prob = isr.IcpProblem(srcd, tgtd, np.eye(4)[None])
t, _ = timed(lambda: (prob.run(20.0, 30, 0.0, 0.0)))
print(f"   loop only (problem prepared): {t*1e3:.2f} ms")
cad = api._points(synth.make_cloud(50000, 3), api._device())
t, c = timed(lambda: isr.chamfer_distance(srcd, cad))
print(f"config1 chamfer 100k vs 50k: {t*1e3:.2f} ms -> {float(c):.4f}")
# config 5: 64 starts x 250k
src, tgt, _ = synth.icp_pair(250000, 250000, 6, 7)
inits = np.stack([synth.rotation_about_z(2 * np.pi * k / 64) for k in range(64)]) if hasattr(synth, "rotation_about_z") else None
if inits is None:
    inits = np.tile(np.eye(4), (64, 1, 1))
    for k in range(64):
        a = 2 * np.pi * k / 64
        inits[k, :2, :2] = [[np.cos(a), -np.sin(a)], [np.sin(a), np.cos(a)]]
srcd, tgtd = api._points(src, api._device()), api._points(tgt, api._device())
t, m = timed(lambda: isr.multistart_icp(srcd, tgtd, inits, 20.0, 30), reps=2)
print(f"config5 multistart 64 x 250k, 30 it + Chamfer ranking: {t*1e3:.1f} ms; best start {int(m.order[0])} chamfer {m.chamfer[m.order[0]]:.4f}")
