"""Developer probe (not the bench): raw kernel timings on one GPU."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import imagesequenceregistrationfor6dposeestimationlabeling_b200 as isr
from imagesequenceregistrationfor6dposeestimationlabeling_b200 import synth, api, _lib

torch.cuda.set_device(0)
def timeit(fn, reps=3, warm=1):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e-3)
    return min(ts), float(np.median(ts))

print("ffma scalar TF/s", isr.measure_fp32_peak(False), "packed", isr.measure_fp32_peak(True))
def morton_sort(c):
    lo, hi = c.min(0), c.max(0)
    g = np.clip(((c - lo) / (hi - lo + 1e-9) * 1023).astype(np.uint64), 0, 1023)
    def spread(v):
        v = (v | (v << 16)) & 0x030000FF
        v = (v | (v << 8)) & 0x0300F00F
        v = (v | (v << 4)) & 0x030C30C3
        v = (v | (v << 2)) & 0x09249249
        return v
    code = spread(g[:, 0]) | (spread(g[:, 1]) << 1) | (spread(g[:, 2]) << 2)
    return c[np.argsort(code, kind="stable")]
SORT = os.environ.get("PROBE_SORT", "0") == "1"
which = sys.argv[1:] or ["nn", "verify", "icp"]
if "nn" in which:
    N = 100000
    cloud = synth.make_cloud(N, 1)
    if SORT: cloud = morton_sort(cloud)
    for B in (128,):
        P = np.tile(np.eye(4), (B, 1, 1))
        rng = np.random.default_rng(0)
        for k in range(B): P[k, :3, :3] = synth.random_rotation(rng)
        cen = api.centroid_of(cloud)
        perm = api.spatial_order(cloud) if os.environ.get("PROBE_LIBSORT", "1") == "1" else None
        for name, q, t in (("direct", api.pack_soa(cloud, P), api.pack_soa(cloud)),
                           ("exact", api.prepare_cloud(cloud, P, centroid=cen, perm=perm),
                            api.prepare_cloud(cloud, centroid=cen, perm=perm, stage_centroids=True))):
            for idx in (False, True):
                best, med = timeit(lambda: api.nearest_neighbors_soa(q, t, return_index=idx))
                fl = 8.0 * N * N * B
                print(f"nn {name} B={B} idx={idx}: {best*1e3:.1f} ms best, {med*1e3:.1f} med -> {fl/best/1e12:.2f} TF/s algorithmic ({fl/best/1e12/74.5*100:.1f}% of 74.5)")
if "verify" in which:
    N = 100000; B = 256
    cloud = synth.make_cloud(N, 1)
    if SORT: cloud = morton_sort(cloud)
    R_true, _ = synth.true_pose(3)
    Rs, _, k0 = synth.make_candidates(B, 10, R_true=R_true, t_true=np.zeros(3))
    Mq, Mt = synth.verification_matrices(Rs, R_true)
    cd = api._points(cloud, api._device()); Mqd = api._poses(Mq, api._device()); Mtd = api._poses(Mt, api._device())
    best, med = timeit(lambda: isr.verify_poses(cd, Mqd, Mtd), reps=2)
    print(f"verify B={B}: {best:.3f}s -> {B/best:.1f} cand/s; {16.0*N*N*B/best/1e12:.2f} TF/s")
if "icp" in which:
    src, tgt, _ = synth.icp_pair(1000000, 1000000, 4, 5)
    if SORT: src, tgt = morton_sort(src), morton_sort(tgt)
    prob = isr.IcpProblem(src, tgt, np.eye(4)[None])
    def one():
        prob.accumulate(20.0); prob.solve(prob.ns, 0.0, 0.0, False)
    best, med = timeit(one, reps=2)
    print(f"icp 1Mx1M one iteration: {best:.3f}s -> {1/best:.2f} it/s; {8e12/best/1e12:.2f} TF/s")

if "prune" in which:
    import ctypes
    lib = _lib.load()
    def pairs():
        c8 = (ctypes.c_uint64 * 8)()
        _lib.check(lib.isr_profile_nn_counters(c8))
        w = max(c8[4], 1)
        if c8[7]: print(f"   slowest warp: {(c8[7] >> 44) * 1024 / 1e6:.2f} Mcycles, scanned {(c8[7] >> 24) & 0xFFFFF}, exact tests {c8[7] & 0xFFFFFF}")
        print(f"   per warp: scanned sub-tiles {c8[0]/w/4:.1f}, stage spheres {c8[1]/w:.1f}, stage candidates {c8[2]/w:.1f}, exact sub tests {c8[3]/w:.1f}, flagged units {c8[5]/w:.1f}, resolve passes {c8[6]/w:.1f}; warps {c8[4]}")
        ev, an = ctypes.c_uint64(0), ctypes.c_uint64(0)
        _lib.check(lib.isr_profile_nn_pairs(ctypes.byref(ev), ctypes.byref(an)))
        return ev.value, an.value
    N = 100000; B = int(os.environ.get("PROBE_B", "256"))
    cloud = synth.make_cloud(N, 1)
    R_true, _ = synth.true_pose(3)
    Rs, _, k0 = synth.make_candidates(B, 10, R_true=R_true, t_true=np.zeros(3))
    Mq, Mt = synth.verification_matrices(Rs, R_true)
    cd = api._points(cloud, api._device()); Mqd = api._poses(Mq, api._device()); Mtd = api._poses(Mt, api._device())
    res = {}
    for on in (True, False):
        api.set_nn_pruning(on)
        best, med = timeit(lambda: isr.verify_poses(cd, Mqd, Mtd), reps=2)
        lib.isr_profile_enable(1); pairs()
        r = isr.verify_poses(cd, Mqd, Mtd); torch.cuda.synchronize()
        ev, an = pairs(); lib.isr_profile_enable(0)
        ms = (ctypes.c_double * 5)(); ln = (ctypes.c_uint64 * 5)()
        lib.isr_profile_collect(ms, ln)
        print("   ms by kind (transform, nn, reduce, icp_acc, icp_solve):", [round(x, 3) for x in ms], list(ln))
        res[on] = r.losses.cpu().numpy()
        print(f"verify B={B} prune={on}: {best:.4f}s -> {B/best:.1f} cand/s; evaluated {ev:.3e} of {an:.3e} pairs ({ev/max(an,1)*100:.2f}%), {8*ev/best/1e12:.2f} TF/s on evaluated pairs; best {r.best_index} (k0 {k0})")
    print("verify losses identical:", np.array_equal(res[True], res[False]))
    src, tgt, _ = synth.icp_pair(1000000, 1000000, 4, 5)
    for on in (True, False):
        api.set_nn_pruning(on)
        prob = isr.IcpProblem(src, tgt, np.eye(4)[None])
        def one():
            prob.accumulate(20.0); prob.solve(prob.ns, 0.0, 0.0, False)
        best, med = timeit(one, reps=3 if on else 1, warm=1)
        lib.isr_profile_enable(1); pairs(); one(); torch.cuda.synchronize(); ev, an = pairs(); lib.isr_profile_enable(0)
        ms = (ctypes.c_double * 5)(); ln = (ctypes.c_uint64 * 5)()
        lib.isr_profile_collect(ms, ln)
        print("   ms by kind (transform, nn, reduce, icp_acc, icp_solve):", [round(x, 3) for x in ms], list(ln))
        print(f"icp 1Mx1M prune={on}: {best*1e3:.2f} ms/it -> {1/best:.1f} it/s; evaluated {ev/max(an,1)*100:.3f}% of pairs; T[0,:]={prob.results(False)[0].transformation[0]}")
    api.set_nn_pruning(True)
