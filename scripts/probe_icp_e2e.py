"""Developer probe: where the wall clock of api.icp goes at 1M x 1M (host arrays in, result out)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from imagesequenceregistrationfor6dposeestimationlabeling_b200 import api, synth, _lib

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
iters = 50
torch.cuda.set_device(0)
src, tgt, _ = synth.icp_pair(n, n, 4, 5)
api.icp(src[:5000], tgt[:5000], np.eye(4), 20.0, max_iteration=2)
torch.cuda.synchronize()
def T(label, t0):
    torch.cuda.synchronize(); t1 = time.perf_counter(); print(f"  {label}: {1e3 * (t1 - t0):.2f} ms"); return t1
dev = torch.device("cuda:0")
for rep in range(3):
    print("pass", rep)
    t00 = t0 = time.perf_counter()
    s, slo = api._points_hilo(src, dev); t1 = T("H2D src", t0)
    t = api._points(tgt, dev); t1 = T("H2D tgt", t1)
    c = api.centroid_of(t, dev); t1 = T("centroid", t1)
    po = api.spatial_order(t, dev); t1 = T("spatial_order tgt", t1)
    soa = api.prepare_cloud(t, centroid=c, perm=po, stage_centroids=True, device=dev); t1 = T("prepare tgt", t1)
    ps = api.spatial_order(s, dev); t1 = T("spatial_order src", t1)
    lib = _lib.load()
    ws = api._workspace(lib.isr_icp_workspace_bytes(n, n, 1), dev); t1 = T(f"workspace {ws.numel() / 1e6:.0f} MB", t1)
    z = torch.zeros((1, n), dtype=torch.int32, device=dev); z2 = torch.zeros((1, n), dtype=torch.uint8, device=dev); t1 = T("zeros", t1)
    print(f"  pieces total {1e3 * (t1 - t00):.2f} ms")
    del s, slo, t, c, po, soa, ps, ws, z, z2
    t0 = time.perf_counter()
    prob = api.IcpProblem(src, tgt, np.eye(4)[None]); t1 = T("IcpProblem", t0)
    prob.run(20.0, iters - 1, 0.0, 0.0); t1 = T("run", t1)
    r = prob.results()[0]; t1 = T("results", t1)
    cs = r.correspondence_set; t1 = T("correspondence_set", t1)
    print(f"  total {1e3 * (t1 - t0):.2f} ms")
    del prob, r, cs
