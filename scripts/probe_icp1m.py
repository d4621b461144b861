"""Developer probe: the 1M x 1M ICP search alone (isr_icp_run, 20 iterations)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import imagesequenceregistrationfor6dposeestimationlabeling_b200 as isr
from imagesequenceregistrationfor6dposeestimationlabeling_b200 import synth, _lib

torch.cuda.set_device(0)
lib = _lib.load()
n = int(os.environ.get("PROBE_N", "1000000"))
src, tgt, _ = synth.icp_pair(n, n, 4, 5)
prob = isr.IcpProblem(src, tgt, np.eye(4)[None])
prob.run(20.0, 2, 0.0, 0.0); prob.reopen(); torch.cuda.synchronize()
best = 1e9
for rep in range(3):
    lib.isr_profile_enable(1); lib.isr_profile_collect(None, None)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); prob.run(20.0, 19, 0.0, 0.0); e1.record(); e1.synchronize(); prob.reopen()
    ms = (ctypes.c_double * 5)(); ln = (ctypes.c_uint64 * 5)()
    lib.isr_profile_collect(ms, ln); lib.isr_profile_enable(0)
    best = min(best, e0.elapsed_time(e1) / 20)
c8 = (ctypes.c_uint64 * 8)(); lib.isr_profile_nn_counters(c8)
print(f"icp {n}: {best:.3f} ms/it -> {1e3/best:.1f} it/s; nn {ms[1]/20:.3f} ms; warps {c8[4]//20}; slowest warp {(c8[7] >> 44) * 1024 / 1e6:.2f} Mcyc scanned {(c8[7] >> 24) & 0xFFFFF} tests {c8[7] & 0xFFFFFF}")
