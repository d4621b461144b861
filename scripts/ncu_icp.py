"""ncu target: four 1M x 1M ICP evaluations (the 4th pruned search runs with warm hints)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import imagesequenceregistrationfor6dposeestimationlabeling_b200 as isr
from imagesequenceregistrationfor6dposeestimationlabeling_b200 import synth

torch.cuda.set_device(0)
src, tgt, _ = synth.icp_pair(1000000, 1000000, 4, 5)
prob = isr.IcpProblem(src, tgt, np.eye(4)[None])
prob.run(20.0, 3, 0.0, 0.0)
torch.cuda.synchronize()
print(prob.results(False)[0].fitness)
