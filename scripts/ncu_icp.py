"""ncu target: fused ICP evaluations at 1M x 1M on the 1/world curve shard of the source (the
last launch runs with warm hints).  python scripts/ncu_icp.py [world] [evaluations]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import imagesequenceregistrationfor6dposeestimationlabeling_b200 as isr
from imagesequenceregistrationfor6dposeestimationlabeling_b200 import api, synth

world = int(sys.argv[1]) if len(sys.argv) > 1 else 1
evals = int(sys.argv[2]) if len(sys.argv) > 2 else 4
torch.cuda.set_device(0)
n = 1000000
src, tgt, _ = synth.icp_pair(n, n, 4, 5)
perm = api.spatial_order(src).cpu().numpy()
shard = src[perm[:(n + world - 1) // world]]
prob = isr.IcpProblem(shard, tgt, np.eye(4)[None])
prob.run(20.0, evals - 1, 0.0, 0.0)
torch.cuda.synchronize()
print(prob.results(False)[0].fitness)
