"""Developer probe: target parts on / off on one curve shard -- identical results, time per iteration.
    python scripts/probe_tp.py [world] [rank] [factor_x10]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from imagesequenceregistrationfor6dposeestimationlabeling_b200 import api, dist, synth

world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
rank = int(sys.argv[2]) if len(sys.argv) > 2 else 7
torch.cuda.set_device(0)
n = 1_000_000
src, tgt, _ = synth.icp_pair(n, n, 4, 5)
perm = api.spatial_order(src).cpu().numpy()
lo, hi = dist.shard_bounds(n, rank, world)
shard = src[perm[lo:hi]]
res = {}
for label, env in (("off", {"ISR_NN_TP_SLOTS": "0"}), ("on", {"ISR_NN_TP_SLOTS": "96"}),
                   ("on, factor 1.0", {"ISR_NN_TP_SLOTS": "96", "ISR_NN_TP_FACTOR_X10": "10"}),
                   ("on, factor 2.0", {"ISR_NN_TP_SLOTS": "96", "ISR_NN_TP_FACTOR_X10": "20"}),
                   ("on, factor 8.0", {"ISR_NN_TP_SLOTS": "96", "ISR_NN_TP_FACTOR_X10": "80"})):
    os.environ.pop("ISR_NN_TP_FACTOR_X10", None)
    os.environ.update(env)
    prob = api.IcpProblem(shard, tgt, np.eye(4)[None])
    prob.run(20.0, 1, 0.0, 0.0)
    prob.reopen()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); prob.run(20.0, 19, 0.0, 0.0); e1.record(); e1.synchronize()
    r = prob.results(True)[0]
    res[label] = r
    print(f"{label}: {e0.elapsed_time(e1) / 20:.4f} ms / iteration, fitness {r.fitness:.6f} rmse {r.inlier_rmse:.9f}", flush=True)
    del prob
base = res["off"]
for k, r in res.items():
    same = np.array_equal(r.transformation, base.transformation) and r.inlier_rmse == base.inlier_rmse
    corr = np.array_equal(np.asarray(r.correspondence_set), np.asarray(base.correspondence_set))
    print(f"{k}: pose bits equal {same}, correspondences equal {corr}")
