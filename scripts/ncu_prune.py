"""ncu target: one pruned verification chunk (64 candidates x 100k) and one pruned ICP pass."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import imagesequenceregistrationfor6dposeestimationlabeling_b200 as isr
from imagesequenceregistrationfor6dposeestimationlabeling_b200 import synth

torch.cuda.set_device(0)
N, B = 100000, 64
cloud = synth.make_cloud(N, 1)
R_true, _ = synth.true_pose(3)
Rs, _, k0 = synth.make_candidates(B, 10, R_true=R_true, t_true=np.zeros(3))
Mq, Mt = synth.verification_matrices(Rs, R_true)
r = isr.verify_poses(cloud, Mq, Mt)
torch.cuda.synchronize()
print(r.best_index, k0)
if len(sys.argv) > 1 and sys.argv[1] == "icp":
    src, tgt, _ = synth.icp_pair(1000000, 1000000, 4, 5)
    prob = isr.IcpProblem(src, tgt, np.eye(4)[None])
    prob.accumulate(20.0); prob.solve(prob.ns, 0.0, 0.0, False)
    torch.cuda.synchronize()
