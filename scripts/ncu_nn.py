"""Short driver for ncu captures of K2: B candidates x N x N, one direction, no indices."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from imagesequenceregistrationfor6dposeestimationlabeling_b200 import synth, api
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
N = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
idx = len(sys.argv) > 3 and sys.argv[3] == "idx"
torch.cuda.set_device(0)
cloud = synth.make_cloud(N, 1)
rng = np.random.default_rng(0)
P = np.tile(np.eye(4), (B, 1, 1))
for k in range(B): P[k, :3, :3] = synth.random_rotation(rng)
mode = os.environ.get("NCU_MODE", "exact")
if os.environ.get("PROBE_SORT", "0") == "1":
    lo, hi = cloud.min(0), cloud.max(0)
    g = np.clip(((cloud - lo) / (hi - lo + 1e-9) * 1023).astype(np.uint64), 0, 1023)
    def spread(v):
        v = (v | (v << 16)) & 0x030000FF
        v = (v | (v << 8)) & 0x0300F00F
        v = (v | (v << 4)) & 0x030C30C3
        v = (v | (v << 2)) & 0x09249249
        return v
    code = spread(g[:, 0]) | (spread(g[:, 1]) << 1) | (spread(g[:, 2]) << 2)
    cloud = cloud[np.argsort(code, kind="stable")]
if mode == "direct":
    q = api.pack_soa(cloud, P); t = api.pack_soa(cloud)
else:
    cen = api.centroid_of(cloud)
    perm = api.spatial_order(cloud)
    q = api.prepare_cloud(cloud, P, centroid=cen, perm=perm)
    t = api.prepare_cloud(cloud, centroid=cen, perm=perm, stage_centroids=True)
for _ in range(3):
    r = api.nearest_neighbors_soa(q, t, return_index=idx)
torch.cuda.synchronize()
print("ok", float(r.d2.sum()))
