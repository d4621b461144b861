import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from imagesequenceregistrationfor6dposeestimationlabeling_b200 import synth, api
torch.cuda.set_device(0)
N = 100000; B = 8
cloud = synth.make_cloud(N, 1)
R_true, _ = synth.true_pose(3)
Rs, _, k0 = synth.make_candidates(B, 10, R_true=R_true, t_true=np.zeros(3))
Mq, Mt = synth.verification_matrices(Rs, R_true)
cen = api.centroid_of(cloud)
for use_perm in (False, True):
    perm = api.spatial_order(cloud) if use_perm else None
    q = api.prepare_cloud(cloud, Mq, centroid=cen, centre_poses=Mt, perm=perm)
    t = api.prepare_cloud(cloud, Mt, centroid=cen, centre_poses=Mt, perm=perm, stage_centroids=use_perm)
    print("perm", use_perm, "batch", B, "queries", B * N, "warp-subtiles total", B * 98 * 4 * 3136, file=sys.stderr)
    for _ in range(2):
        api.nearest_neighbors_soa(q, t, return_index=False, use_lo=False)
        torch.cuda.synchronize()
