import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from imagesequenceregistrationfor6dposeestimationlabeling_b200 import _lib
lib = ctypes.CDLL(_lib.LIB_PATH)
torch.cuda.set_device(0)
sink = torch.zeros(4, device="cuda"); src = torch.full((32,), 1.0000001, device="cuda")
fl = ctypes.c_double(0)
for mode in (2, 9, 8):
    best = 0
    for _ in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        lib.isr_bench_ffma2_pattern(148 * 8, 4096, mode, ctypes.c_void_p(sink.data_ptr()), ctypes.c_void_p(src.data_ptr()), ctypes.byref(fl), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
        e1.record(); e1.synchronize()
        best = max(best, fl.value / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    print("mode", mode, "TF/s", round(best, 2))
