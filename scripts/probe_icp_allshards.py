"""Developer probe: time of one fused ICP iteration for EVERY curve shard of an N-GPU run, measured
one after the other on one GPU (the sharded loop runs at the pace of its slowest shard).
    python scripts/probe_icp_allshards.py [world] [points] [iters]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from imagesequenceregistrationfor6dposeestimationlabeling_b200 import api, dist, synth

world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 20
torch.cuda.set_device(0)
src, tgt, _ = synth.icp_pair(n, n, 4, 5)
perm = api.spatial_order(src).cpu().numpy()
ts = []
for r in range(world):
    lo, hi = dist.shard_bounds(n, r, world)
    prob = api.IcpProblem(src[perm[lo:hi]], tgt, np.eye(4)[None])
    prob.run(20.0, 1, 0.0, 0.0)
    prob.reopen()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); prob.run(20.0, iters - 1, 0.0, 0.0); e1.record(); e1.synchronize()
    ts.append(e0.elapsed_time(e1) / iters)
    del prob
print(f"world {world}: per-shard ms/iteration " + " ".join(f"{t:.3f}" for t in ts) +
      f"; slowest {max(ts):.3f} ms -> {1e3 / max(ts):.0f} it/s")
