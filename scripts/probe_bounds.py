"""Developer probe: isr_adds_bounds on its own (time per pose pair, width of the bounds, share of
pairs decided at 0.1 x 120 mm) for aligned / failed predictions."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from imagesequenceregistrationfor6dposeestimationlabeling_b200 import api, synth
torch.cuda.set_device(0)
surface = synth.make_cloud(100000, 1); verts = synth.make_cloud(20000, seed=3)
verts = verts[api.spatial_order(verts).cpu().numpy()]
cen = api.centroid_of(surface)
tgt = api.prepare_cloud(surface, centroid=cen, perm=api.spatial_order(surface), stage_centroids=True)
rng = np.random.default_rng(0)
B = 4096
for name, scale in (("aligned 3deg", 0.05), ("aligned 10deg", 0.17), ("failed", None)):
    M = np.stack([synth.pose_matrix(synth.rotvec_to_matrix(rng.normal(scale=scale, size=3)) if scale else synth.random_rotation(rng),
                                    rng.normal(scale=1.5, size=3)) for _ in range(B)])
    Md = api._poses(M, torch.device("cuda", 0))
    api.adds_bounds(verts, Md, tgt, presorted=True); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); lo, hi = api.adds_bounds(verts, Md, tgt, presorted=True); e1.record(); e1.synchronize()
    t = e0.elapsed_time(e1) * 1e-3
    ex = api.adds_fixed(verts, Md[:256], surface).losses.cpu().numpy()
    lo, hi = lo.cpu().numpy(), hi.cpu().numpy()
    print(f"{name}: {t / B * 1e6:.2f} us per pair; exact mean {ex.mean():.2f}, lower {lo[:256].mean():.2f}, upper {hi[:256].mean():.2f}; "
          f"decided {np.mean((hi < 12) | (lo >= 12)):.3f}; bracket ok {bool(np.all(lo[:256] <= ex) and np.all(ex <= hi[:256]))}")
