"""Developer probe (ISR_PHASE_LOG build of libisr.so): where one fused ICP iteration spends its
time -- per-CTA phase stamps (SM cycles) and the launch's wall-clock marks (globaltimer, ns).
    ISR_LIBISR_PATH=.../variants/phase.so python scripts/probe_phases.py [points] [world] [rank]"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from imagesequenceregistrationfor6dposeestimationlabeling_b200 import _lib, api, dist, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
world = int(sys.argv[2]) if len(sys.argv) > 2 else 8
rank = int(sys.argv[3]) if len(sys.argv) > 3 else 0
torch.cuda.set_device(0)
lib = _lib.load()
src, tgt, _ = synth.icp_pair(n, n, 4, 5)
perm = api.spatial_order(src).cpu().numpy()
lo, hi = dist.shard_bounds(n, rank, world)
prob = api.IcpProblem(src[perm[lo:hi]], tgt, np.eye(4)[None])
prob.run(20.0, 3, 0.0, 0.0)
prob.reopen()
torch.cuda.synchronize()
cap = 40000
# (run with ISR_ICP_PDL=0: the stamps start at CTA start.  The log keeps the LAST launch of a
# 7-launch run, i.e. the re-cut launch list)
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 6
log = torch.zeros((cap, 4), dtype=torch.int64, device="cuda")
lib.isr_debug_cta_log(ctypes.c_void_p(log.data_ptr()), cap)
prob.run(20.0, iters, 0.0, 0.0)
torch.cuda.synchronize()
lib.isr_debug_cta_log(None, 0)
L = log.cpu().numpy().astype(np.uint64)
marks = L[cap - 1].astype(np.int64)
L2 = L[cap // 2:cap - 1]
L = L[:cap // 2 - 1]
keep = L[:, 0] > 0
L, L2 = L[keep], L2[:len(keep)][keep]
lo32 = np.uint64(0xFFFFFFFF)
search_end = L[:, 0].astype(np.float64)
q_ready = (L[:, 1] & lo32).astype(np.float64)
hints = (L[:, 1] >> np.uint64(32)).astype(np.float64)
rows = (L[:, 2] & lo32).astype(np.float64)
epi = (L[:, 2] >> np.uint64(32)).astype(np.float64)
total = ((L[:, 3] & np.uint64(0xFFFFF)).astype(np.float64)) * 256
print(f"{len(L)} CTAs logged (1/{world} shard, rank {rank}, {hi - lo} points)")
def st(name, v):
    print(f"  {name}: median {np.median(v):.0f}  mean {v.mean():.0f}  p90 {np.percentile(v, 90):.0f}  max {v.max():.0f} cycles")
st("query copy + pose + split", q_ready)
st("hints", hints - q_ready)
st("row spheres (+ cuts)", rows - hints)
st("walk + tests + scans", search_end - rows)
st("epilogue (gather, sums)", epi - search_end)
st("tail (tickets .. solve)", total - epi)
st("whole CTA", total)
lo = np.uint64(0xFFFFFFFF)
sh = np.uint64(32)
print("inside walk + tests + scans:")
st("  consume: exact tests + copy issue", (L2[:, 0] & lo).astype(np.float64))
st("  wait for the sub-tile", (L2[:, 0] >> sh).astype(np.float64))
st("  filter scan", (L2[:, 1] & lo).astype(np.float64))
st("  resolve", (L2[:, 1] >> sh).astype(np.float64))
st("  produce (sphere walk -> FIFO)", (L2[:, 2] & lo).astype(np.float64))
st("  nearest-first sort", (L2[:, 2] >> sh).astype(np.float64))
nsc = (L2[:, 3] >> np.uint64(48)).astype(np.float64)
nte = ((L2[:, 3] >> sh) & np.uint64(0xFFFF)).astype(np.float64)
npa = ((L2[:, 3] >> np.uint64(16)) & np.uint64(0xFFFF)).astype(np.float64)
nqu = (L2[:, 3] & np.uint64(0xFFFF)).astype(np.float64)
print(f"  per CTA: {nsc.mean():.1f} scanned tiles, {nqu.mean():.1f} quarter scans, {nte.mean():.1f} exact tests, {npa.mean():.1f} resolve passes")
print(f"  per event (mean cycles): exact test {(L2[:, 0] & lo).astype(np.float64).sum() / max(nte.sum(), 1):.0f}, "
      f"quarter scan {(L2[:, 1] & lo).astype(np.float64).sum() / max(nqu.sum(), 1):.0f}, "
      f"resolve pass {(L2[:, 1] >> sh).astype(np.float64).sum() / max(npa.sum(), 1):.0f}")
slow = np.argsort(-total)[:8]
for o in slow:
    print(f"   slow CTA: total {total[o]:.0f}: copy {q_ready[o]:.0f} hints {hints[o] - q_ready[o]:.0f} rows {rows[o] - hints[o]:.0f} "
          f"search {search_end[o] - rows[o]:.0f} epilogue {epi[o] - search_end[o]:.0f} tail {total[o] - epi[o]:.0f}")
