"""Developer probe: pruned verification (256 candidates x 100k) and one 1M x 1M ICP iteration,
with the per-warp counters of the pruned search.  Faster than `perf_probe.py prune` (no
exhaustive legs)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import imagesequenceregistrationfor6dposeestimationlabeling_b200 as isr
from imagesequenceregistrationfor6dposeestimationlabeling_b200 import synth, api, _lib

torch.cuda.set_device(0)
lib = _lib.load()

def timeit(fn, reps=3, warm=1):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e-3)
    return min(ts)

def counters():
    c8 = (ctypes.c_uint64 * 8)()
    _lib.check(lib.isr_profile_nn_counters(c8))
    w = max(c8[4], 1)
    s = (f"per warp: scanned {c8[0]/w/4:.1f} (unit equivalents), stage cand {c8[2]/w:.1f}, exact tests {c8[3]/w:.1f}, "
         f"flagged {c8[5]/w:.1f}, resolve passes {c8[6]/w:.1f}; slowest warp {(c8[7] >> 44) * 1024 / 1e6:.2f} Mcyc "
         f"scanned {(c8[7] >> 24) & 0xFFFFF}")
    ev, an = ctypes.c_uint64(0), ctypes.c_uint64(0)
    _lib.check(lib.isr_profile_nn_pairs(ctypes.byref(ev), ctypes.byref(an)))
    return s, ev.value, an.value

N, B = 100000, 256
cloud = synth.make_cloud(N, 1)
R_true, _ = synth.true_pose(3)
Rs, _, k0 = synth.make_candidates(B, 10, R_true=R_true, t_true=np.zeros(3))
Mq, Mt = synth.verification_matrices(Rs, R_true)
cd = api._points(cloud, api._device()); Mqd = api._poses(Mq, api._device()); Mtd = api._poses(Mt, api._device())
best = timeit(lambda: isr.verify_poses(cd, Mqd, Mtd), reps=3)
lib.isr_profile_enable(1); counters()
r = isr.verify_poses(cd, Mqd, Mtd); torch.cuda.synchronize()
s, ev, an = counters(); lib.isr_profile_enable(0)
print(f"verify B={B}: {B/best:.1f} cand/s; {ev/max(an,1)*100:.2f}% pairs, {8*ev/best/1e12:.2f} TF/s evaluated; {s}; loss sum {float(r.losses.sum().item()):.12f}")
if "noicp" not in sys.argv:
    src, tgt, _ = synth.icp_pair(1000000, 1000000, 4, 5)
    prob = isr.IcpProblem(src, tgt, np.eye(4)[None])
    def one():
        prob.accumulate(20.0); prob.solve(prob.ns, 0.0, 0.0, False)
    best = timeit(one, reps=5, warm=2)
    lib.isr_profile_enable(1); counters(); one(); torch.cuda.synchronize(); s, ev, an = counters(); lib.isr_profile_enable(0)
    print(f"icp 1Mx1M: {best*1e3:.3f} ms/it -> {1/best:.1f} it/s; {ev/max(an,1)*100:.3f}% pairs; {s}")
