"""Developer probe: per-CTA cost of the pruned search inside one fused ICP iteration
(isr_debug_cta_log).  python scripts/probe_cta_log.py [points] [world] [rank]"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from imagesequenceregistrationfor6dposeestimationlabeling_b200 import _lib, api, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
world = int(sys.argv[2]) if len(sys.argv) > 2 else 4
rank = int(sys.argv[3]) if len(sys.argv) > 3 else 0
torch.cuda.set_device(0)
lib = _lib.load()
src, tgt, _ = synth.icp_pair(n, n, 4, 5)
perm = api.spatial_order(src).cpu().numpy()
m = (n + world - 1) // world
shard = src[perm[rank * m:(rank + 1) * m]]
prob = api.IcpProblem(shard, tgt, np.eye(4)[None])
prob.run(20.0, 2, 0.0, 0.0)
prob.reopen()
torch.cuda.synchronize()
cap = 40000
log = torch.zeros((cap, 4), dtype=torch.int64, device="cuda")
lib.isr_debug_cta_log(ctypes.c_void_p(log.data_ptr()), cap)
# (iterations > 0: the log keeps the LAST launch of a longer run, i.e. the re-cut launch list)
prob.run(20.0, int(sys.argv[4]) if len(sys.argv) > 4 else 0, 0.0, 0.0)
torch.cuda.synchronize()
lib.isr_debug_cta_log(None, 0)
L = log.cpu().numpy().astype(np.uint64)
L = L[L[:, 0] > 0]
if len(sys.argv) > 4:
    L = L[:int(api._lib.load().isr_version() * 0 + len(L))]
cyc = L[:, 0].astype(np.float64)
nscan = (L[:, 1] >> np.uint64(32)).astype(np.int64)
ntest = (L[:, 1] & np.uint64(0xFFFFFFFF)).astype(np.int64)
nquart = (L[:, 2] >> np.uint64(32)).astype(np.int64)
ncand = (L[:, 2] & np.uint64(0xFFFFFFFF)).astype(np.int64)
blk = (L[:, 3] >> np.uint64(32)).astype(np.int64)
code = ((L[:, 3] >> np.uint64(20)) & np.uint64(0xFFF)).astype(np.int64)   # rows mask << 4 | target part
npass = (L[:, 3] & np.uint64(0xFFFFF)).astype(np.int64)
print(f"{len(L)} CTAs; cycles: mean {cyc.mean():.0f} median {np.median(cyc):.0f} p90 {np.percentile(cyc, 90):.0f} "
      f"p99 {np.percentile(cyc, 99):.0f} max {cyc.max():.0f}")
print(f"sum of cycles / 148 SMs / 16 warps = {cyc.sum() / 148 / 16:.0f} (ideal balanced makespan at full occupancy)")
for name, v in (("scanned sub-tiles", nscan), ("exact tests", ntest), ("quarter units", nquart),
                ("candidate stages", ncand), ("resolve passes", npass)):
    print(f"  {name}: mean {v.mean():.1f} median {np.median(v):.0f} p99 {np.percentile(v, 99):.0f} max {v.max()}")
# linear model of the cycles
A = np.stack([np.ones_like(cyc), nscan, ntest, nquart, ncand, npass], 1).astype(np.float64)
coef, *_ = np.linalg.lstsq(A, cyc, rcond=None)
print("cycles ~ %.0f + %.0f*scanned + %.0f*tests + %.0f*quarters + %.0f*stages + %.0f*passes" % tuple(coef))
nrows = np.array([bin(int(c) >> 4).count("1") for c in code])
print("CTAs by rows owned: " + ", ".join(f"{k} rows: {int((nrows == k).sum())}" for k in range(1, 9)) +
      f"; target-part CTAs: {int(((code & 7) != 0).sum())}")
order = np.argsort(-cyc)[:12]
# geometry of the slow blocks: radius of the block and of its rows (stored order of the shard)
sp = api.spatial_order(shard).cpu().numpy()
stored = shard[sp]
for o in order:
    b = int(blk[o])
    pts = stored[b * 256:(b + 1) * 256]
    c = pts.mean(0)
    rad = np.sqrt(((pts - c) ** 2).sum(1).max())
    rows = [np.sqrt(((pts[r * 32:(r + 1) * 32] - pts[r * 32:(r + 1) * 32].mean(0)) ** 2).sum(1).max())
            for r in range(len(pts) // 32)]
    print(f"  blk {b} rows {code[o] >> 4:08b} part {code[o] & 7}: {cyc[o]:.0f} cyc, scanned {nscan[o]}, tests {ntest[o]}, quarters {nquart[o]}, "
          f"stages {ncand[o]}, passes {npass[o]}; block radius {rad:.2f} mm, row radii " +
          " ".join(f"{x:.1f}" for x in rows))
med = np.argsort(cyc)[len(cyc) // 2]
print("median CTA:", cyc[med], nscan[med], ntest[med], nquart[med], ncand[med], npass[med])
