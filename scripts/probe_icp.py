"""Developer probe: ICP iteration time at several sizes (pruned search, hints warm)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import imagesequenceregistrationfor6dposeestimationlabeling_b200 as isr
from imagesequenceregistrationfor6dposeestimationlabeling_b200 import synth, _lib

torch.cuda.set_device(0)
lib = _lib.load()
for n, starts in ((1000000, 1), (500000, 1), (250000, 1), (100000, 1), (250000, 8)):
    src, tgt, _ = synth.icp_pair(n, n, 4, 5)
    inits = np.stack([synth.pose_matrix(synth.rotvec_to_matrix([0, 0, 2 * np.pi * k / max(starts, 1)]), [0, 0, 0])
                      for k in range(starts)])
    prob = isr.IcpProblem(src, tgt, inits)
    def one():
        prob.accumulate(20.0); prob.solve(prob.ns, 0.0, 0.0, False)
    for _ in range(3): one()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    lib.isr_profile_enable(1); lib.isr_profile_collect(None, None)
    e0.record()
    for _ in range(10): one()
    e1.record(); e1.synchronize()
    ms = (ctypes.c_double * 5)(); ln = (ctypes.c_uint64 * 5)()
    lib.isr_profile_collect(ms, ln); lib.isr_profile_enable(0)
    t = e0.elapsed_time(e1) / 10
    print(f"icp {n} x {n}, {starts} start(s): {t:.3f} ms/it -> {1e3 / t:.1f} it/s; nn {ms[1] / 10:.3f} ms, prep {ms[0] / 10:.3f}, acc {ms[3] / 10:.3f}, solve {ms[4] / 10:.3f}; "
          f"T00 {prob.results(False)[0].transformation[0, 0]:.12f}")
