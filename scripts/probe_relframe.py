"""Developer probe: how much tighter are the sub-tile boxes when the TARGET stays in its native
(Hilbert-curve) frame?  Chamfer verification with (Mq, Mt) vs (inv(Mt) Mq, I)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import imagesequenceregistrationfor6dposeestimationlabeling_b200 as isr
from imagesequenceregistrationfor6dposeestimationlabeling_b200 import synth, api, _lib

torch.cuda.set_device(0)
lib = _lib.load()
N, B = 100000, 256
cloud = synth.make_cloud(N, 1)
R_true, _ = synth.true_pose(3)
Rs, _, k0 = synth.make_candidates(B, 10, R_true=R_true, t_true=np.zeros(3))
Mq, Mt = synth.verification_matrices(Rs, R_true)
Mrel = np.linalg.inv(Mt) @ Mq
I = np.tile(np.eye(4), (B, 1, 1))
cd = api._points(cloud, api._device())

def counters():
    c8 = (ctypes.c_uint64 * 8)()
    _lib.check(lib.isr_profile_nn_counters(c8))
    w = max(c8[4], 1)
    s = f"scanned {c8[0]/w/4:.1f} unit-eq, exact tests {c8[3]/w:.1f}, flagged {c8[5]/w:.1f}, passes {c8[6]/w:.1f}"
    ev, an = ctypes.c_uint64(0), ctypes.c_uint64(0)
    _lib.check(lib.isr_profile_nn_pairs(ctypes.byref(ev), ctypes.byref(an)))
    return s

for name, A, Bm, bidir in (("general", Mq, Mt, "chamfer"), ("relative", Mrel, I, "chamfer"),
                           ("general 1-dir", Mq, Mt, "adds"), ("relative 1-dir", Mrel, I, "adds")):
    Ad, Bd = api._poses(A, api._device()), api._poses(Bm, api._device())
    f = lambda: isr.verify_poses(cd, Ad, Bd, mode=bidir)
    f(); torch.cuda.synchronize()
    ts = []
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); r = f(); e1.record(); e1.synchronize(); ts.append(e0.elapsed_time(e1) * 1e-3)
    lib.isr_profile_enable(1); counters(); r = f(); torch.cuda.synchronize(); s = counters(); lib.isr_profile_enable(0)
    print(f"{name:16s}: {B/min(ts):8.1f} cand/s; {s}; best {r.best_index} (k0 {k0}); loss[k0] {float(r.losses[k0]):.12f}")
