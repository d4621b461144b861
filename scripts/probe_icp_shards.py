"""Developer probe (not the bench): time of one fused ICP iteration for the source shard that
rank 0 of an N-GPU run would own, measured on ONE GPU.

    python scripts/probe_icp_shards.py [points] [iters]

For world in 1, 2, 4, 8: the first 1 / world of the source -- (a) along the Hilbert curve (what
dist.icp_sharded does), (b) by row number (a `world`-times sparser sample of the whole
surface) -- registered against the full target with isr_icp_run.  Prints ms per iteration.
Environment knobs of the search (read once per process): ISR_NN_PARTS_MAX.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import imagesequenceregistrationfor6dposeestimationlabeling_b200 as isr
from imagesequenceregistrationfor6dposeestimationlabeling_b200 import api, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 20
torch.cuda.set_device(0)
src, tgt, _ = synth.icp_pair(n, n, 4, 5)
perm = api.spatial_order(src).cpu().numpy()


def time_run(shard, label):
    prob = api.IcpProblem(shard, tgt, np.eye(4)[None])
    prob.run(20.0, 1, 0.0, 0.0)   # warm-up: two evaluations (hints warm)
    prob.reopen()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    prob.run(20.0, iters - 1, 0.0, 0.0)
    e1.record()
    e1.synchronize()
    ms = e0.elapsed_time(e1) / iters
    r = prob.results(False)[0]
    print(f"{label}: {len(shard)} source points, {ms:.4f} ms / iteration ({1e3 / ms:.0f} it/s), "
          f"fitness {r.fitness:.4f} rmse {r.inlier_rmse:.5f}", flush=True)
    del prob


print("ISR_NN_PARTS_MAX", os.environ.get("ISR_NN_PARTS_MAX", "(default 8)"))
for world in (1, 2, 4, 8):
    m = (n + world - 1) // world
    time_run(src[perm[:m]], f"world {world} curve shard")
    if world > 1:
        time_run(src[:m], f"world {world} row shard  ")
