"""Short driver for ncu captures of the HBM / L2-bound kernels: K1, K1', tile spheres, the stepwise
ICP accumulate kernel, the FP64 transform, the ADD-S sphere bounds."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from imagesequenceregistrationfor6dposeestimationlabeling_b200 import synth, api
torch.cuda.set_device(0)
cloud = synth.make_cloud(100000, 1)
rng = np.random.default_rng(0)
P = np.tile(np.eye(4), (256, 1, 1))
for k in range(256): P[k, :3, :3] = synth.random_rotation(rng)
cd = api._points(cloud, api._device()); Pd = api._poses(P, api._device())
cen = api.centroid_of(cd)
for _ in range(2):
    api.transform_points(cd, Pd)
    api.prepare_cloud(cd, Pd, centroid=cen, centre_poses=Pd)
src, tgt, _ = synth.icp_pair(1000000, 1000000, 4, 5)
prob = api.IcpProblem(src, tgt[:100000], np.eye(4)[None])   # 1M sources (real K3), small target to keep the NN short
prob.accumulate(20.0); prob.accumulate(20.0)
d64 = torch.from_numpy(src.astype(np.float64)).cuda()
api.transform_points_f64(d64, P[0], out=d64); api.transform_points_f64(d64, P[1], out=d64)
verts = synth.make_cloud(20000, seed=3)
verts = verts[api.spatial_order(verts).cpu().numpy()]
t7 = api.prepare_cloud(cd, centroid=cen, perm=api.spatial_order(cd), stage_centroids=True)
api.adds_bounds(verts, Pd, t7, presorted=True); api.adds_bounds(verts, Pd, t7, presorted=True)
torch.cuda.synchronize(); print("ok")
