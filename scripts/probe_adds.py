"""Developer probe: exact ADD-S throughput (api.adds_rigid: 20k vertices against the 100k-point
surface, prepared once).  Environment: ISR_NN_PARTS_FORCE.   python scripts/probe_adds.py [pairs]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from imagesequenceregistrationfor6dposeestimationlabeling_b200 import api, synth

nb = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
torch.cuda.set_device(0)
cloud = synth.make_cloud(100000, seed=1)
verts = synth.make_cloud(20000, seed=3)
R_t, t_t = synth.true_pose(3)
Rs_a, ts_a, _ = synth.make_candidates(nb, seed=10, R_true=R_t, t_true=t_t)
Pg = np.stack([synth.pose_matrix(R_t, t_t) for k in range(nb)])
Pp = np.stack([synth.pose_matrix(Rs_a[k], ts_a[k]) for k in range(nb)])
api.adds_rigid(verts, Pg[:64], Pp[:64], cloud)
torch.cuda.synchronize()
for _ in range(2):
    t0 = time.perf_counter()
    lr = api.adds_rigid(verts, Pg, Pp, cloud).losses.cpu().numpy()
    dt = time.perf_counter() - t0
print(f"PARTS_FORCE={os.environ.get('ISR_NN_PARTS_FORCE', '-')}: {nb / dt:.0f} pose pairs/s, checksum {lr.sum():.9f}")
