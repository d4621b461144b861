"""GPU parity tests proper: every call goes through the C ABI (csrc/libisr.so)."""
import os

import numpy as np
import pytest

from oracle import c_oracle, oracle

pytestmark = pytest.mark.gpu


def _rand_poses(b, seed, trans=100.0):
    from imagesequenceregistrationfor6dposeestimationlabeling_b200 import synth
    rng = np.random.default_rng(seed)
    P = np.tile(np.eye(4), (b, 1, 1))
    for k in range(b):
        P[k, :3, :3] = synth.random_rotation(rng)
        P[k, :3, 3] = rng.normal(scale=trans, size=3)
    return P


# ---- K1 ---------------------------------------------------------------------------------
@pytest.mark.parametrize("n", [1, 5, 256, 1000, 4099])
def test_transform_points_matches_float64(gpu, n):
    rng = np.random.default_rng(n)
    pts = rng.normal(scale=60, size=(n, 3)).astype(np.float32)
    P = _rand_poses(7, n + 1, trans=700.0)
    got = gpu.transform_points(pts, P).cpu().numpy()
    ref = np.stack([oracle.transform(pts, P[k]) for k in range(len(P))])
    assert got.shape == (7, n, 3) and got.dtype == np.float32
    # FP64 math, one rounding to float32: at most half an ulp from the float64 result
    # (+1 ulp slack for the different summation order inside BLAS)
    ulp = np.spacing(np.abs(ref).astype(np.float32)).astype(np.float64)
    assert np.all(np.abs(got.astype(np.float64) - ref) <= 1.0 * ulp)


def test_prepare_cloud_hi_lo_planes(gpu):
    """SoA7: hi + lo reproduces the float64 transformed, centred coordinate; norm = |hi|^2."""
    rng = np.random.default_rng(0)
    pts = rng.normal(scale=60, size=(3000, 3)).astype(np.float32)
    P = _rand_poses(3, 2, trans=700.0)
    cen = rng.normal(scale=5, size=3)
    soa = gpu.prepare_cloud(pts, P, centroid=cen, centre_poses=P)
    d = soa.data.cpu().numpy().astype(np.float64)
    assert d.shape == (3, 7, 3072)
    for k in range(3):
        c = P[k, :3, :3] @ cen + P[k, :3, 3]
        ref = oracle.transform(pts, P[k]) - c
        np.testing.assert_allclose(d[k, 0:3, :3000].T + d[k, 4:7, :3000].T, ref, rtol=0, atol=2e-11)
        hi = d[k, 0:3, :3000].astype(np.float32)
        nrm = (hi.astype(np.float64) ** 2).sum(0)
        np.testing.assert_allclose(d[k, 3, :3000], nrm, rtol=3e-7)
        assert np.all(d[k, 0:3, 3000:] == np.float32(1e18)) and np.all(d[k, 4:7, 3000:] == 0)
    cg = gpu.centroid_of(pts).cpu().numpy()
    np.testing.assert_allclose(cg, pts.astype(np.float64).mean(0), rtol=0, atol=1e-10)


def _hilbert_code(q, bits=10):
    """3-D Hilbert index (Skilling's transpose form), numpy restatement of sort.cu."""
    X = [q[:, c].astype(np.uint32).copy() for c in range(3)]
    M = 1 << (bits - 1)
    Q = M
    while Q > 1:
        P = Q - 1
        for c in range(3):
            m = (X[c] & Q) != 0
            X[0] = np.where(m, X[0] ^ P, X[0])
            t = np.where(m, 0, (X[0] ^ X[c]) & P).astype(np.uint32)
            X[0] ^= t
            X[c] ^= t
        Q >>= 1
    X[1] ^= X[0]
    X[2] ^= X[1]
    t = np.zeros_like(X[0])
    Q = M
    while Q > 1:
        t = np.where((X[2] & Q) != 0, t ^ (Q - 1), t).astype(np.uint32)
        Q >>= 1
    code = np.zeros(len(q), dtype=np.uint64)
    for bit in range(bits - 1, -1, -1):
        for c in range(3):
            code = (code << np.uint64(1)) | (((X[c] ^ t) >> bit) & 1).astype(np.uint64)
    return code


def test_hilbert_restatement_is_a_continuous_curve():
    g = np.stack(np.meshgrid(*[np.arange(16)] * 3, indexing="ij"), -1).reshape(-1, 3)
    code = _hilbert_code(g, bits=4)
    assert len(np.unique(code)) == len(code)
    steps = np.abs(np.diff(g[np.argsort(code)], axis=0)).sum(1)
    assert steps.min() == 1 and steps.max() == 1


@pytest.mark.parametrize("n", [1, 5, 2048, 2049, 70001])
def test_spatial_order_is_a_hilbert_sorted_permutation(gpu, n):
    rng = np.random.default_rng(n)
    pts = rng.normal(scale=50, size=(n, 3)).astype(np.float32)
    perm = gpu.spatial_order(pts).cpu().numpy()
    assert sorted(perm.tolist()) == list(range(n))
    lo, hi = pts.min(0), pts.max(0)
    ext = max(float((hi - lo).max()), 1e-30)
    q = np.clip(((pts - lo) * np.float32(1023.0 / ext)), 0, 1023).astype(np.uint64)
    code = _hilbert_code(q)
    key = (code[perm].astype(np.uint64) << np.uint64(32)) | perm.astype(np.uint64)
    # allow for float rounding of the quantisation at cell borders: codes must be sorted in
    # all but a handful of places, and identical inputs always give the identical permutation
    if n > 1:
        assert np.mean(np.diff(key.astype(np.float64)) > 0) > 0.999
    np.testing.assert_array_equal(perm, gpu.spatial_order(pts).cpu().numpy())


def test_pack_soa_layout_and_padding(gpu):
    pts = np.arange(30, dtype=np.float32).reshape(10, 3)
    soa = gpu.pack_soa(pts)
    d = soa.data.cpu().numpy()
    assert d.shape == (1, 3, 1024) and soa.n == 10
    np.testing.assert_array_equal(d[0, :, :10], pts.T)
    assert np.all(d[0, :, 10:] == np.float32(1e18))


# ---- K2 ---------------------------------------------------------------------------------
@pytest.mark.parametrize("nq,nt", [(1, 1), (7, 3), (1000, 1000), (1025, 4097), (3000, 20000),
                                   (300, 70000)])
def test_nn_direct_bit_exact_vs_fma_emulation(gpu, nq, nt):
    """Direct-difference kernel: d2 bits and indices equal the C float32 emulation of the
    kernel's rounding sequence."""
    rng = np.random.default_rng(nq * 7 + nt)
    q = rng.normal(scale=40, size=(nq, 3)).astype(np.float32)
    t = rng.normal(scale=40, size=(nt, 3)).astype(np.float32)
    res = gpu.nearest_neighbors(q, t, mode="direct")
    d2, idx = res.d2.cpu().numpy(), res.idx.cpu().numpy()
    rd2, ridx = c_oracle.nn_f32_fma(q, t)
    np.testing.assert_array_equal(d2.view(np.uint32), rd2.view(np.uint32))
    np.testing.assert_array_equal(idx, ridx)


@pytest.mark.parametrize("nq,nt,scale,offset", [
    (1, 1, 40, 0), (7, 3, 40, 0), (1000, 1000, 40, 0), (1025, 4097, 40, 0), (3000, 20000, 40, 0),
    (300, 70000, 40, 0), (2000, 30000, 40, 700), (2000, 30000, 0.01, 0), (2000, 30000, 5000, -3000)])
def test_nn_exact_equals_float64_bruteforce(gpu, nq, nt, scale, offset):
    """Production kernel (FP32 filter + FP64 resolve): the index IS the float64 brute-force
    argmin (lowest index on ties) and d2 is the float64 distance rounded to float32 --
    for object-frame, camera-frame (700 mm offset), tiny and huge coordinate ranges."""
    rng = np.random.default_rng(nq * 7 + nt)
    q = (rng.normal(scale=scale, size=(nq, 3)) + offset).astype(np.float32)
    t = (rng.normal(scale=scale, size=(nt, 3)) + offset).astype(np.float32)
    res = gpu.nearest_neighbors(q, t)
    d2, idx = res.d2.cpu().numpy(), res.idx.cpu().numpy()
    rd2, ridx = c_oracle.nn_f64(q, t)
    np.testing.assert_array_equal(idx, ridx)
    np.testing.assert_allclose(d2, rd2, rtol=1.2e-7, atol=0)


def test_nn_indices_match_float64_oracle_on_surface_cloud(gpu):
    """FP32 direct-difference argmin == float64 KD-tree argmin."""
    from imagesequenceregistrationfor6dposeestimationlabeling_b200 import synth
    t = synth.make_cloud(100000, seed=1)
    q = synth.make_cloud(20000, seed=2)
    for offset in (np.zeros(3, np.float32), np.array([0, 0, 700], np.float32)):
        dk, ik = oracle.nearest(q + offset, t + offset)
        res = gpu.nearest_neighbors(q + offset, t + offset)
        np.testing.assert_array_equal(res.idx.cpu().numpy(), ik)          # exact, no exceptions
        np.testing.assert_allclose(res.dist.cpu().numpy(), dk, rtol=1e-7, atol=0)
        res = gpu.nearest_neighbors(q + offset, t + offset, mode="direct")
        idx = res.idx.cpu().numpy()
        mism = np.nonzero(idx != ik)[0]
        # FP32 direct form: any mismatch must be a float32-resolution tie
        if len(mism):
            d_alt = np.linalg.norm((q + offset)[mism].astype(np.float64)
                                   - (t + offset)[idx[mism]].astype(np.float64), axis=1)
            assert np.all(np.abs(d_alt - dk[mism]) <= 1e-6 * np.maximum(dk[mism], 1e-3))
        assert len(mism) <= 2
        np.testing.assert_allclose(res.dist.cpu().numpy(), dk, rtol=1e-5, atol=1e-6)


def test_tile_spheres_bound_their_points(gpu):
    """Every stored point lies inside the sphere of its 64-point sub-tile and of its
    1024-point stage (prepare.cu); padding-only tiles are marked r = -1."""
    from imagesequenceregistrationfor6dposeestimationlabeling_b200 import api, synth
    for n, offset in ((100000, 0.0), (5000, 700.0), (64, 0.0), (1, 3.0)):
        cloud = synth.make_cloud(max(n, 8), seed=3)[:n] + np.float32(offset)
        P = _rand_poses(3, n, trans=50.0)
        c = api.prepare_cloud(cloud, P, perm=api.spatial_order(cloud), stage_centroids=True)
        data = c.data.cpu().numpy().astype(np.float64)
        pts = np.moveaxis(data[:, 0:3] + data[:, 4:7], 1, 2)            # [B, npad, 3] hi + lo
        nst = c.npad // 1024
        assert c.stage_c.shape == (3, nst + (nst + 31) // 32, 4)
        chunk = c.stage_c.cpu().numpy().astype(np.float64)[:, nst:]
        for spheres, tile in ((c.sub_c.cpu().numpy(), 64), (c.stage_c.cpu().numpy()[:, :nst], 1024)):
            nt = c.npad // tile
            assert spheres.shape == (3, nt, 4)
            for b in range(3):
                for k in range(nt):
                    lo, hi = k * tile, min((k + 1) * tile, n)
                    if hi <= lo:
                        assert spheres[b, k, 3] == -1.0
                        continue
                    d = np.linalg.norm(pts[b, lo:hi] - spheres[b, k, :3].astype(np.float64), axis=1)
                    assert d.max() <= spheres[b, k, 3]
                    assert spheres[b, k, 3] <= d.max() * 1.001 + 1e-4 * (1 + abs(offset))
        # the chunk spheres (32 stages each) bound their points as well
        for b in range(3):
            for k in range(chunk.shape[1]):
                lo, hi = k * 32768, min((k + 1) * 32768, n)
                if hi <= lo:
                    assert chunk[b, k, 3] == -1.0
                    continue
                assert np.linalg.norm(pts[b, lo:hi] - chunk[b, k, :3], axis=1).max() <= chunk[b, k, 3]
        # the sub-tile boxes: centre = the sphere's, half-extents r * k / 1023 per axis
        sub, box = c.sub_c.cpu().numpy().astype(np.float64), c.sub_box.cpu().numpy().astype(np.int64)
        for b in range(3):
            for k in range(c.npad // 64):
                lo, hi = k * 64, min((k + 1) * 64, n)
                if hi <= lo:
                    assert box[b, k] == 0x3FFFFFFF
                    continue
                h = sub[b, k, 3] * np.array([box[b, k] & 1023, (box[b, k] >> 10) & 1023, (box[b, k] >> 20) & 1023]) / 1023.0
                ext = np.abs(pts[b, lo:hi] - sub[b, k, :3]).max(axis=0)
                assert (ext <= h).all()
                assert (h <= ext * 1.001 + sub[b, k, 3] * 2e-3 + 1e-4 * (1 + abs(offset))).all()


@pytest.mark.parametrize("case", ["aligned", "rotated", "far_apart", "clusters", "lattice_ties",
                                  "tiny_target", "camera_frame"])
def test_nn_pruned_equals_exhaustive(gpu, case):
    """Tile pruning only skips work: indices and d2 are bit-identical to the exhaustive scan
    (and to the float64 brute-force oracle)."""
    from imagesequenceregistrationfor6dposeestimationlabeling_b200 import api, synth
    rng = np.random.default_rng(5)
    if case == "aligned":
        q, t = synth.make_cloud(30000, seed=1), synth.make_cloud(40000, seed=2)
    elif case == "rotated":
        t = synth.make_cloud(40000, seed=2)
        q = (synth.make_cloud(30000, seed=1).astype(np.float64) @ synth.random_rotation(rng).T).astype(np.float32)
    elif case == "far_apart":
        q, t = synth.make_cloud(5000, seed=1) + np.float32(900.0), synth.make_cloud(20000, seed=2)
    elif case == "clusters":
        cen = rng.normal(scale=200, size=(40, 3))
        t = (cen[rng.integers(0, 40, 30000)] + rng.normal(scale=2.0, size=(30000, 3))).astype(np.float32)
        q = (cen[rng.integers(0, 40, 9000)] + rng.normal(scale=30.0, size=(9000, 3))).astype(np.float32)
    elif case == "lattice_ties":
        g = np.stack(np.meshgrid(np.arange(20), np.arange(20), np.arange(20), indexing="ij"), -1)
        t = g.reshape(-1, 3).astype(np.float32)
        t = np.concatenate([t, t[::-1]], axis=0)
        q = (g.reshape(-1, 3) + 0.5).astype(np.float32)
    elif case == "tiny_target":
        q, t = synth.make_cloud(9000, seed=1), synth.make_cloud(100, seed=2)[:37]
    else:
        off = np.array([30, -20, 700], np.float32)
        q, t = synth.make_cloud(20000, seed=1) + off, synth.make_cloud(70000, seed=2) + off
    assert api.get_nn_pruning()
    try:
        pr = gpu.nearest_neighbors(q, t)
        api.set_nn_pruning(False)
        ex = gpu.nearest_neighbors(q, t)
    finally:
        api.set_nn_pruning(True)
    np.testing.assert_array_equal(pr.idx.cpu().numpy(), ex.idx.cpu().numpy())
    np.testing.assert_array_equal(pr.d2.cpu().numpy().view(np.uint32), ex.d2.cpu().numpy().view(np.uint32))
    sub = rng.choice(len(q), size=min(len(q), 3000), replace=False)
    rd2, ridx = c_oracle.nn_f64(q[sub], t)
    np.testing.assert_array_equal(pr.idx.cpu().numpy()[sub], ridx)


def test_verify_and_icp_pruned_equal_exhaustive(gpu):
    from imagesequenceregistrationfor6dposeestimationlabeling_b200 import api, synth
    cloud = synth.make_cloud(20000, seed=1)
    R_true, _ = synth.true_pose(3)
    Rs, _, k0 = synth.make_candidates(24, seed=5, R_true=R_true, t_true=np.zeros(3))
    Mq, Mt = synth.verification_matrices(Rs, R_true)
    src, tgt, _ = synth.icp_pair(15000, 18000, 6, 7)
    out = {}
    try:
        for on in (True, False):
            api.set_nn_pruning(on)
            v = gpu.verify_poses(cloud, Mq, Mt, mode="chamfer")
            r = gpu.icp(src, tgt, np.eye(4), 20.0, max_iteration=8)
            out[on] = (v.losses.cpu().numpy(), v.best_index, r.transformation, r.fitness, r.inlier_rmse,
                       r.correspondence_set)
    finally:
        api.set_nn_pruning(True)
    for a, b in zip(out[True], out[False]):
        np.testing.assert_array_equal(np.asarray(a), np.asarray(b))   # bit-identical
    assert out[True][1] == k0


def test_nn_lattice_ties_pick_lowest_index(gpu):
    g = np.stack(np.meshgrid(np.arange(6), np.arange(6), np.arange(6), indexing="ij"), -1)
    t = g.reshape(-1, 3).astype(np.float32)
    t = np.concatenate([t, t], axis=0)  # every target duplicated: ties everywhere
    q = (g.reshape(-1, 3) + 0.5).astype(np.float32)  # cell centres: 8-way ties (x2)
    rd2, ridx = c_oracle.nn_f32_fma(q, t)
    for mode in ("exact", "direct"):
        res = gpu.nearest_neighbors(q, t, mode=mode)
        np.testing.assert_array_equal(res.idx.cpu().numpy(), ridx)
        assert np.all(res.idx.cpu().numpy() < len(t) // 2)
        np.testing.assert_array_equal(res.d2.cpu().numpy(), rd2)


def test_nn_identical_and_shifted_clouds(gpu):
    from imagesequenceregistrationfor6dposeestimationlabeling_b200 import synth
    a = synth.make_cloud(5000, seed=9)
    res = gpu.nearest_neighbors(a, a)
    assert np.all(res.d2.cpu().numpy() == 0)
    d64, i64 = c_oracle.nn_f64(a, a)
    np.testing.assert_array_equal(res.idx.cpu().numpy(), i64)  # identity unless duplicates
    assert float(gpu.chamfer_distance(a, a)) == 0.0


def test_nn_batched_shared_target(gpu):
    rng = np.random.default_rng(3)
    q = rng.normal(scale=30, size=(5, 700, 3)).astype(np.float32)
    t = rng.normal(scale=30, size=(2500, 3)).astype(np.float32)
    res = gpu.nearest_neighbors(q, t, mode="direct")
    res2 = gpu.nearest_neighbors(q, t)
    for b in range(5):
        rd2, ridx = c_oracle.nn_f32_fma(q[b], t)
        np.testing.assert_array_equal(res.d2[b].cpu().numpy(), rd2)
        np.testing.assert_array_equal(res.idx[b].cpu().numpy(), ridx)
        d64, i64 = c_oracle.nn_f64(q[b], t)
        np.testing.assert_array_equal(res2.idx[b].cpu().numpy(), i64)
        np.testing.assert_allclose(res2.d2[b].cpu().numpy(), d64, rtol=1.2e-7)


def test_chamfer_and_distance_match_oracle(gpu):
    from imagesequenceregistrationfor6dposeestimationlabeling_b200 import synth
    a = synth.make_cloud(30000, seed=11)
    b = synth.make_cloud(25000, seed=12)
    np.testing.assert_allclose(float(gpu.chamfer_distance(a, b)), oracle.chamfer(a, b), rtol=1e-6)
    d = gpu.point_cloud_distance(a, b).cpu().numpy()
    np.testing.assert_allclose(d, oracle.compute_point_cloud_distance(a, b), rtol=1e-5, atol=1e-6)
    assert np.all(gpu.point_cloud_distance(a, np.zeros((0, 3))).cpu().numpy() == 0)


# ---- verification ---------------------------------------------------------------------
def test_verify_chamfer_matches_oracle_and_selects_k0(gpu):
    from imagesequenceregistrationfor6dposeestimationlabeling_b200 import synth
    cloud = synth.make_cloud(20000, seed=1)
    R_true, _ = synth.true_pose(3)
    Rs, _, k0 = synth.make_candidates(40, seed=10, R_true=R_true, t_true=np.zeros(3))
    Mq, Mt = synth.verification_matrices(Rs, R_true)
    res = gpu.verify_poses(cloud, Mq, Mt, mode="chamfer")
    ref, ref_best = oracle.verify_matrices(cloud, cloud, Mq, Mt, bidirectional=True)
    np.testing.assert_allclose(res.losses.cpu().numpy(), ref, rtol=1e-5)
    assert res.best_index == ref_best == k0
    np.testing.assert_allclose(res.best_loss, ref[ref_best], rtol=1e-5)


def test_verify_reference_loop_form(gpu):
    """Same numbers as the literal loop of verfication.py:61-108 (image pairs i, i+1)."""
    from imagesequenceregistrationfor6dposeestimationlabeling_b200 import helpers, synth
    pc1 = synth.make_cloud(8000, seed=4)
    rng = np.random.default_rng(0)
    n = 9
    gt_R = [synth.random_rotation(rng) for _ in range(n)]
    gt_T = [rng.normal(scale=50, size=3) + [0, 0, 700] for _ in range(n)]
    pred_R = [gt_R[i] @ synth.rotvec_to_matrix(rng.normal(scale=0.05, size=3)) for i in range(n)]
    pred_T = [gt_T[i] + rng.normal(scale=2, size=3) for i in range(n)]
    ref_list, ref_idx, ref_min = oracle.verify_chamfer(pc1, gt_R, gt_T, pred_R, pred_T)
    Mq = np.tile(np.eye(4), (n - 1, 1, 1))
    Mt = np.tile(np.eye(4), (n - 1, 1, 1))
    for i in range(n - 1):
        R_rel, _ = helpers.calculate_relative_pose(gt_R[i], gt_T[i], gt_R[i + 1], gt_T[i + 1])
        Mq[i, :3, :3] = pred_R[i + 1].T            # pcpred = pc1 . R2pred
        Mt[i, :3, :3] = R_rel.T @ pred_R[i]        # pcgt = pc1 . R1pred^T . R_rel
    res = gpu.verify_poses(pc1, Mq, Mt, mode="chamfer")
    np.testing.assert_allclose(res.losses.cpu().numpy(), ref_list, rtol=1e-5)
    assert res.best_index == ref_idx
    np.testing.assert_allclose(res.best_loss, ref_min, rtol=1e-5)


def test_verify_duplicates_first_min_and_invalid_mask(gpu):
    from imagesequenceregistrationfor6dposeestimationlabeling_b200 import synth
    cloud = synth.make_cloud(3000, seed=2)
    P = _rand_poses(6, 5, trans=0.0)
    Mq = np.concatenate([P, P])          # candidates 6..11 duplicate 0..5
    Mt = np.tile(np.eye(4), (12, 1, 1))
    res = gpu.verify_poses(cloud, Mq, Mt)
    losses = res.losses.cpu().numpy()
    np.testing.assert_array_equal(losses[:6], losses[6:])   # deterministic, bit-identical
    assert res.best_index == int(np.argmin(losses)) < 6
    valid = np.ones(12, dtype=bool)
    valid[res.best_index] = False
    res2 = gpu.verify_poses(cloud, Mq, Mt, valid_mask=valid)
    assert res2.best_index == res.best_index + 6
    assert np.isinf(res2.losses.cpu().numpy()[res.best_index])


def test_adds_matches_sklearn_restatement(gpu):
    from imagesequenceregistrationfor6dposeestimationlabeling_b200 import synth
    surface = synth.make_cloud(15000, seed=1).astype(np.float64)
    verts = synth.make_cloud(4000, seed=3).astype(np.float64)
    P = _rand_poses(5, 1, trans=5.0)
    G = _rand_poses(5, 2, trans=5.0)
    got = gpu.adds(verts, G[:, :3, :3], G[:, :3, 3], P[:, :3, :3], P[:, :3, 3], surface).cpu().numpy()
    for k in range(5):
        ref = oracle.ADDS(verts, G[k, :3, :3], G[k, :3, 3], P[k, :3, :3], P[k, :3, 3], surface)
        np.testing.assert_allclose(got[k], ref, rtol=1e-5)


# ---- ICP --------------------------------------------------------------------------------
CLOUD_RADIUS = 60.0  # mm; a rotation error dR moves the translation by ~|dR| * radius


def _close_T(T, Tref, rtol=1e-5):
    """north_star tolerance: refined pose within 1e-5 relative.  R: relative Frobenius.
    t = mu_t - R mu_s, so its error is measured against |t| + cloud radius."""
    assert np.linalg.norm(T[:3, :3] - Tref[:3, :3]) <= rtol * np.linalg.norm(Tref[:3, :3])
    assert np.linalg.norm(T[:3, 3] - Tref[:3, 3]) <= rtol * (np.linalg.norm(Tref[:3, 3]) + CLOUD_RADIUS)


def test_evaluate_registration_matches_oracle(gpu):
    from imagesequenceregistrationfor6dposeestimationlabeling_b200 import synth
    src, tgt, _ = synth.icp_pair(20000, 25000, 6, 7, half=True)
    T = synth.pose_matrix(synth.rotvec_to_matrix([0.01, -0.02, 0.015]), [0.5, -0.3, 0.2])
    r = gpu.evaluate_registration(src, tgt, 20.0, T)
    o = oracle.evaluate_registration(src, tgt, 20.0, T)
    assert r.n_corr == len(o.correspondence_set)
    assert r.fitness == o.fitness
    np.testing.assert_allclose(r.inlier_rmse, o.inlier_rmse, rtol=1e-6)
    cs = r.correspondence_set
    assert cs.shape == o.correspondence_set.shape
    np.testing.assert_array_equal(cs, o.correspondence_set)   # exact correspondences


def test_icp_matches_oracle(gpu):
    from imagesequenceregistrationfor6dposeestimationlabeling_b200 import synth
    src, tgt, Tm = synth.icp_pair(30000, 30000, 4, 5)
    r = gpu.icp(src, tgt, np.eye(4), 20.0)
    o = oracle.registration_icp(src, tgt, 20.0, np.eye(4))
    assert r.iterations == o.iterations
    _close_T(r.transformation, o.transformation)
    assert abs(r.fitness - o.fitness) <= 1.0 / len(src)
    np.testing.assert_allclose(r.inlier_rmse, o.inlier_rmse, rtol=1e-5)


def test_icp_refine_pose_returns_the_reference_triple(gpu):
    """``(R, t, loss)`` as pose_refine.py:21-22,101-104 returns it -- R 3x3
    float64, t of shape (3,), a scalar loss -- equal to the ICP result started from (R, t)
    (oracle.registration_icp) and to api.icp; float64 sources keep their precision; the
    reference's own argument list (images, renderer objects) is rejected loudly."""
    from imagesequenceregistrationfor6dposeestimationlabeling_b200 import api, synth
    src, tgt, _ = synth.icp_pair(12000, 14000, 4, 5)
    R0 = synth.rotvec_to_matrix([0.01, -0.008, 0.012])
    t0 = np.array([0.4, -0.3, 0.2])
    R, t, loss = gpu.icp_refine_pose(R0, t0, src, tgt)
    assert isinstance(R, np.ndarray) and R.shape == (3, 3) and R.dtype == np.float64
    assert isinstance(t, np.ndarray) and t.shape == (3,) and t.dtype == np.float64
    assert isinstance(loss, float)
    o = oracle.registration_icp(src, tgt, 20.0, api.pose_from_Rt(R0, t0))
    T = api.pose_from_Rt(R, t)
    _close_T(T, o.transformation)
    np.testing.assert_allclose(loss, o.inlier_rmse, rtol=1e-5)
    r = gpu.icp(src, tgt, api.pose_from_Rt(R0, t0), 20.0)
    np.testing.assert_array_equal(T, r.transformation)
    assert loss == r.inlier_rmse
    np.testing.assert_allclose(R @ R.T, np.eye(3), atol=1e-12)
    assert gpu.refine_pose is gpu.icp_refine_pose
    # a float64 camera-frame source (icp.py:68) and a capped iteration count
    Rg, tg_ = synth.true_pose(3)
    src64 = src.astype(np.float64) @ Rg.T + tg_
    Minv = np.linalg.inv(api.pose_from_Rt(Rg, tg_))
    R2, t2, loss2 = gpu.icp_refine_pose(Minv[:3, :3], Minv[:3, 3], src64, tgt, max_iteration=5)
    o2 = oracle.registration_icp(src64, tgt, 20.0, Minv, max_iteration=5)
    _close_T(api.pose_from_Rt(R2, t2), o2.transformation)
    np.testing.assert_allclose(loss2, o2.inlier_rmse, rtol=1e-5)
    with pytest.raises(TypeError):
        gpu.icp_refine_pose(R0, t0, np.zeros((64, 64, 3, 2)), object())


def test_icp_recovers_known_motion_and_converges(gpu):
    from imagesequenceregistrationfor6dposeestimationlabeling_b200 import synth
    tgt = synth.make_cloud(20000, seed=8)
    Tm = synth.pose_matrix(synth.rotvec_to_matrix([0.004, 0.003, -0.005]), [0.05, -0.04, 0.03])
    src = oracle.transform(tgt, Tm).astype(np.float32)
    r = gpu.icp(src, tgt, np.eye(4), 20.0)
    assert r.fitness == 1.0 and r.iterations < 30 and r.converged
    np.testing.assert_allclose(r.transformation, np.linalg.inv(Tm), atol=2e-5)


def test_icp_script_flow_config1_small(gpu):
    """icp.py:64-117 end to end (camera-frame source, init = inverse predicted pose)."""
    from imagesequenceregistrationfor6dposeestimationlabeling_b200 import synth
    import imagesequenceregistrationfor6dposeestimationlabeling_b200.o3d_compat as o3d
    upper = synth.make_cloud(20000, seed=1, half="upper")
    lower = synth.make_cloud(20000, seed=2, half="lower")
    cad = synth.make_cloud(10000, seed=3).astype(np.float64)
    R_GT, t_GT = synth.true_pose(3)
    R_pred = R_GT @ synth.rotvec_to_matrix(np.deg2rad(2.0) * np.array([0.6, 0.0, 0.8]))
    t_pred = t_GT + np.array([1.2, -1.0, 1.2])
    ev_o, reg_o, ch_o = oracle.icp_script(upper, lower, R_GT, t_GT, R_pred, t_pred, cad)

    actual_upper = upper.dot(R_GT.T) + t_GT
    source = o3d.geometry.PointCloud()
    source.points = o3d.utility.Vector3dVector(actual_upper)
    target = o3d.geometry.PointCloud()
    target.points = o3d.utility.Vector3dVector(lower)
    M = np.eye(4); M[:3, :3] = R_pred; M[:3, 3] = t_pred
    init = np.linalg.inv(M)
    ev = o3d.pipelines.registration.evaluate_registration(source, target, 20, init)
    reg = o3d.pipelines.registration.registration_icp(
        source, target, 20, init, o3d.pipelines.registration.TransformationEstimationPointToPoint())
    assert abs(ev.fitness - ev_o.fitness) <= 1.0 / len(upper)
    np.testing.assert_allclose(ev.inlier_rmse, ev_o.inlier_rmse, rtol=1e-5)
    _close_T(reg.transformation, reg_o.transformation)
    assert abs(reg.fitness - reg_o.fitness) <= 1.0 / len(upper)
    np.testing.assert_allclose(reg.inlier_rmse, reg_o.inlier_rmse, rtol=1e-5)
    assert "RegistrationResult with fitness=" in repr(reg)
    transformed_source = source.transform(reg.transformation)
    assert transformed_source is source
    full = transformed_source + target
    assert len(full) == len(upper) + len(lower)
    cadpc = o3d.geometry.PointCloud(); cadpc.points = o3d.utility.Vector3dVector(cad)
    d1 = np.mean(np.asarray(full.compute_point_cloud_distance(cadpc)))
    d2 = np.mean(np.asarray(cadpc.compute_point_cloud_distance(full)))
    np.testing.assert_allclose((d1 + d2) / 2, ch_o, rtol=1e-5)


def test_icp_threshold_edge_and_empty_correspondences(gpu):
    tgt = np.array([[0, 0, 0], [100, 0, 0]], dtype=np.float32)
    src = np.array([[0, 0, 20.0], [100, 0, 20.0 * (1 - 1e-6)]], dtype=np.float32)
    r = gpu.evaluate_registration(src, tgt, 20.0)
    assert r.n_corr == 1 and r.fitness == 0.5          # d == thr is excluded (strict <)
    np.testing.assert_array_equal(r.correspondence_set, [[1, 1]])
    far = src + np.float32(1000)
    r = gpu.icp(far, tgt, np.eye(4), 20.0)
    assert r.n_corr == 0 and r.fitness == 0 and r.inlier_rmse == 0
    np.testing.assert_array_equal(r.transformation, np.eye(4))
    o = oracle.registration_icp(far, tgt, 20.0, np.eye(4))
    assert r.iterations == o.iterations == 1


def test_icp_search_accumulate_halves_equal_the_fused_call(gpu):
    """isr_icp_search + isr_icp_corr_dist + isr_icp_accumulate_corr (the target-sharded ICP's
    building blocks, here on one rank) reproduce isr_icp_run bit for bit; negative indices
    drop correspondences."""
    import torch
    from imagesequenceregistrationfor6dposeestimationlabeling_b200 import api, dist, synth
    src, tgt, _ = synth.icp_pair(7001, 8000, 6, 7)
    a = gpu.icp(src, tgt, np.eye(4), 20.0, max_iteration=9)
    b = dist.icp_sharded(src, tgt, np.eye(4), 20.0, max_iteration=9, shard="target")[0]
    np.testing.assert_array_equal(a.transformation, b.transformation)
    assert (a.fitness, a.inlier_rmse, a.iterations) == (b.fitness, b.inlier_rmse, b.iterations)
    prob = api.IcpProblem(src, tgt, np.eye(4)[None])
    idx = prob.search()
    D = prob.corr_dist(idx).cpu().numpy()[0]
    dk, ik = oracle.nearest(src, tgt)
    np.testing.assert_array_equal(idx.cpu().numpy()[0], ik)
    np.testing.assert_allclose(D, dk * dk, rtol=1e-12, atol=1e-18)
    full = prob.accumulate_corr(idx, 20.0).cpu().numpy().copy()
    half = idx.clone()
    half[0, ::2] = -1
    part = prob.accumulate_corr(half, 20.0).cpu().numpy().copy()
    other = idx.clone()
    other[0, 1::2] = -1
    rest = prob.accumulate_corr(other, 20.0).cpu().numpy().copy()
    assert part[0, 16] + rest[0, 16] == full[0, 16] == len(src)
    np.testing.assert_allclose(part + rest, full, rtol=1e-12)
    assert np.isinf(prob.corr_dist(half).cpu().numpy()[0, ::2]).all()


def test_icp_kernel_fused_exchange_single_rank_equals_icp_run(gpu):
    """isr_icp_run_sharded with a one-rank PeerExchange (the rank stores its sums into its own
    exchange buffer and the solve kernel waits on the flag and reads them back) reproduces
    isr_icp_run bit for bit -- single start, multi-start with starts finishing at different
    iterations, reuse of the exchange across problems, and max_dist <= 0."""
    from imagesequenceregistrationfor6dposeestimationlabeling_b200 import api, dist, synth
    src, tgt, _ = synth.icp_pair(7001, 8000, 6, 7)
    a = gpu.icp(src, tgt, np.eye(4), 20.0, max_iteration=9)
    b = dist.icp_sharded(src, tgt, np.eye(4), 20.0, max_iteration=9, exchange="peer")[0]
    np.testing.assert_array_equal(a.transformation, b.transformation)
    assert (a.fitness, a.inlier_rmse, a.iterations) == (b.fitness, b.inlier_rmse, b.iterations)
    inits = np.stack([synth.pose_matrix(synth.rotvec_to_matrix([0, 0, 2 * np.pi * k / 5]), [0, 0, 0])
                          for k in range(5)])
    ref = api.IcpProblem(src, tgt, inits)
    ref.run(20.0, 30, 1e-4, 1e-3)
    got = dist.icp_sharded(src, tgt, inits, 20.0, max_iteration=30, relative_fitness=1e-4,
                           relative_rmse=1e-3, exchange="peer")
    want = ref.results(with_correspondences=False)
    assert min(w.iterations for w in want) < 30   # some start stops by the criteria, on the device
    for g, w in zip(got, want):
        np.testing.assert_array_equal(g.transformation, w.transformation)
        assert (g.fitness, g.inlier_rmse, g.iterations) == (w.fitness, w.inlier_rmse, w.iterations)
    z = dist.icp_sharded(src, tgt, np.eye(4), 0.0, max_iteration=3, exchange="peer")[0]
    assert z.fitness == 0.0 and z.inlier_rmse == 0.0
    np.testing.assert_array_equal(z.transformation, np.eye(4))
    with pytest.raises(ValueError):
        dist.icp_sharded(src, tgt, np.tile(np.eye(4), (65, 1, 1)), 20.0, max_iteration=1, exchange="peer")


def test_fused_icp_target_parts_are_bit_identical(gpu, monkeypatch):
    """Shallow grids run the widest single rows of the split blocks as four CTAs that own
    disjoint shares of the target's stages and merge their (FP64 distance, index) answers
    (nn2.cu, kTargetParts).  With the threshold lowered so that every row of every split block
    runs that way, poses, rmse and correspondences equal the run without target parts bit for
    bit, and the per-CTA log shows that part CTAs did run."""
    import ctypes
    import torch
    from imagesequenceregistrationfor6dposeestimationlabeling_b200 import _lib, api, synth
    n = 120_000
    src, tgt, _ = synth.icp_pair(n, n, 4, 5)
    perm = api.spatial_order(src).cpu().numpy()
    shard = src[perm[: n // 4]]          # a compact run of the curve, as dist.icp_sharded cuts it
    lib = _lib.load()
    out = {}
    for label, slots in (("off", "0"), ("on", "96")):
        monkeypatch.setenv("ISR_NN_TP_SLOTS", slots)
        monkeypatch.setenv("ISR_NN_TP_FACTOR_X10", "1")
        prob = api.IcpProblem(shard, tgt, np.eye(4)[None])
        cap = 20000
        log = torch.zeros((cap, 4), dtype=torch.int64, device="cuda")
        lib.isr_debug_cta_log(ctypes.c_void_p(log.data_ptr()), cap)
        try:
            prob.run(20.0, 6, 0.0, 0.0)
            torch.cuda.synchronize()
        finally:
            lib.isr_debug_cta_log(None, 0)
        tparts = (log.cpu().numpy().astype(np.uint64)[:, 3] >> np.uint64(20)) & np.uint64(7)   # target part of the CTA
        out[label] = (prob.results(True)[0], int((tparts != 0).sum()))
    (a, parts_off), (b, parts_on) = out["off"], out["on"]
    assert parts_off == 0 and parts_on >= 4, (parts_off, parts_on)
    np.testing.assert_array_equal(a.transformation, b.transformation)
    assert (a.fitness, a.inlier_rmse, a.iterations) == (b.fitness, b.inlier_rmse, b.iterations)
    np.testing.assert_array_equal(np.asarray(a.correspondence_set), np.asarray(b.correspondence_set))
    o = oracle.registration_icp(shard, tgt, 20.0, np.eye(4), max_iteration=6, relative_fitness=0.0, relative_rmse=0.0)
    np.testing.assert_array_equal(np.asarray(b.correspondence_set), o.correspondence_set)
    np.testing.assert_allclose(b.transformation, o.transformation, rtol=1e-7, atol=1e-7)


def test_fused_icp_launch_list_recut_from_measured_costs_is_bit_identical(gpu, monkeypatch):
    """One-wave grids re-cut the launch list of the fused ICP iteration from the cycles that the
    previous iteration's CTAs measured (nn2.cu, block_rebalance_kernel: light blocks merge, heavy
    ones split down to target parts).  The reduction order of the sums does not depend on how blocks
    are cut into CTAs, so poses, rmse and correspondences equal the run with the static list bit for
    bit (and the oracle), while the per-CTA log shows that the list did change."""
    import ctypes
    import torch
    from imagesequenceregistrationfor6dposeestimationlabeling_b200 import _lib, api, synth
    n = 200_000
    src, tgt, _ = synth.icp_pair(n, n, 4, 5)
    perm = api.spatial_order(src).cpu().numpy()
    shard = src[perm[: n // 2]]
    lib = _lib.load()
    out = {}
    for label in ("0", "1"):
        monkeypatch.setenv("ISR_ICP_REBALANCE", label)
        prob = api.IcpProblem(shard, tgt, np.eye(4)[None])
        cap = 20000
        log = torch.zeros((cap, 4), dtype=torch.int64, device="cuda")
        lib.isr_debug_cta_log(ctypes.c_void_p(log.data_ptr()), cap)   # (every launch overwrites: the last one stays)
        try:
            prob.run(20.0, 6, 0.0, 0.0)
            torch.cuda.synchronize()
        finally:
            lib.isr_debug_cta_log(None, 0)
        L = log.cpu().numpy().astype(np.uint64)
        L = L[L[:, 0] > 0]
        codes = np.sort(((L[:, 3] >> np.uint64(20)) & np.uint64(0xFFF)).astype(np.int64))   # (rows mask, target part)
        out[label] = (prob.results(True)[0], codes)
    (a, codes_a), (b, codes_b) = out["0"], out["1"]
    assert len(codes_a) != len(codes_b) or (codes_a != codes_b).any()   # the list was re-cut
    np.testing.assert_array_equal(a.transformation, b.transformation)
    assert (a.fitness, a.inlier_rmse, a.iterations) == (b.fitness, b.inlier_rmse, b.iterations)
    np.testing.assert_array_equal(np.asarray(a.correspondence_set), np.asarray(b.correspondence_set))
    o = oracle.registration_icp(shard, tgt, 20.0, np.eye(4), max_iteration=6, relative_fitness=0.0, relative_rmse=0.0)
    np.testing.assert_array_equal(np.asarray(b.correspondence_set), o.correspondence_set)
    np.testing.assert_allclose(b.transformation, o.transformation, rtol=1e-7, atol=1e-7)


@pytest.mark.parametrize("n,radius,scale,offset", [
    (1, 1.0, 1.0, 0.0), (300, 3.0, 1.0, 0.0), (6000, 4.0, 1.0, 0.0), (6000, 0.05 * 60 / 1.8, 1.0, 0.0),
    (5000, 0.002, 1.0 / 2000, 0.0), (4000, 4.0, 1.0, 700.0), (3000, 1e-3, 1.0, 900.0)])
def test_radius_count_equals_float64_bruteforce(gpu, n, radius, scale, offset):
    """the counts behind remove_radius_outlier equal the float64 brute
    force (strict <, self included) for object-frame, tiny-scale, camera-frame clouds and for
    a radius far below the coordinate magnitude (all-FP64 fallback)."""
    from imagesequenceregistrationfor6dposeestimationlabeling_b200 import synth
    pts = (synth.make_cloud(max(n, 8), seed=4)[:n].astype(np.float64) * scale + offset).astype(np.float32)
    got = gpu.radius_neighbor_count(pts, radius).cpu().numpy()
    np.testing.assert_array_equal(got, oracle.radius_count(pts, radius))
    assert got.min() >= 1


def test_radius_count_lattice_boundary_and_target(gpu):
    """Neighbours at exactly the radius are excluded (strict <); separate query/target clouds."""
    g = np.stack(np.meshgrid(np.arange(12), np.arange(12), np.arange(12), indexing="ij"), -1)
    pts = g.reshape(-1, 3).astype(np.float32)
    for radius in (1.0, np.sqrt(2.0), 2.0, 2.0000001):
        got = gpu.radius_neighbor_count(pts, radius).cpu().numpy()
        np.testing.assert_array_equal(got, oracle.radius_count(pts, radius))
    q = (pts[::7] + np.float32(0.25))
    got = gpu.radius_neighbor_count(q, 1.3, target=pts).cpu().numpy()
    np.testing.assert_array_equal(got, oracle.radius_count(q, 1.3, target=pts))


def _pnp_scene(n=5000, seed=0):
    from imagesequenceregistrationfor6dposeestimationlabeling_b200 import synth
    rng = np.random.default_rng(seed)
    p3d = synth.make_cloud(n, seed=5)
    R, t = synth.true_pose(3)
    cam = np.array([[1075.0, 0.0, 360.0], [0.0, 1073.0, 270.0], [0.0, 0.0, 1.0]])
    pc = p3d.astype(np.float64) @ R.T + t
    uv = (pc[:, :2] / pc[:, 2:3]) * [cam[0, 0], cam[1, 1]] + [cam[0, 2], cam[1, 2]]
    uv += rng.normal(scale=0.7, size=uv.shape)
    bad = rng.random(n) < 0.3                      # wrong correspondences
    uv[bad] = rng.uniform([0, 0], [720, 540], size=(int(bad.sum()), 2))
    return p3d, uv.astype(np.float32), cam, R, t


def test_pnp_hypothesis_scoring_equals_oracle(gpu):
    """inlier counts and flags of a batch of PnP hypotheses equal the
    restated OpenCV consensus test, including hypotheses at the 2-pixel edge, behind the
    camera and with z == 0; the helper returns the reference's (R, t, inliers) triple of the
    first best hypothesis and the (1, 1, 1) sentinel when nothing reprojects."""
    from imagesequenceregistrationfor6dposeestimationlabeling_b200 import helpers, synth
    p3d, uv, cam, R, t = _pnp_scene()
    rng = np.random.default_rng(1)
    Rs = [R] + [R @ synth.rotvec_to_matrix(rng.normal(scale=s, size=3)) for s in
                (1e-4, 1e-3, 3e-3, 1e-2, 0.1, 1.0)] + [synth.random_rotation(rng) for _ in range(30)] + [R]
    ts = [t] + [t + rng.normal(scale=s, size=3) for s in (0.01, 0.1, 0.5, 2.0, 10.0, 50.0)] + \
         [t * np.array([1, 1, -1.0])] * 15 + [np.zeros(3)] * 15 + [t]
    P = np.stack([synth.pose_matrix(a, b) for a, b in zip(Rs, ts)])
    counts, flags = gpu.score_pnp_hypotheses(p3d, uv, cam, P, 2.0, return_inliers=True)
    counts, flags = counts.cpu().numpy(), flags.cpu().numpy().astype(bool)
    for k in range(len(P)):
        ref = oracle.pnp_inliers(p3d, uv, cam, Rs[k], ts[k], 2.0)
        np.testing.assert_array_equal(flags[k], ref)
        assert counts[k] == ref.sum()
    assert counts[0] == counts[-1] > 0.5 * len(p3d)
    # every hypothesis twice: the winner must be the FIRST of the two equal maxima
    Rb, tb, inl = helpers.select_pnp_hypothesis(p3d, uv, cam, np.stack(Rs + Rs), np.stack(ts + ts))
    ks = int(np.argmax(counts))
    np.testing.assert_array_equal(Rb, Rs[ks])
    np.testing.assert_array_equal(tb, ts[ks])
    np.testing.assert_array_equal(inl, np.nonzero(oracle.pnp_inliers(p3d, uv, cam, Rs[ks], ts[ks]))[0])
    far = helpers.select_pnp_hypothesis(p3d, uv + 1e4, cam, np.stack(Rs[:3]), np.stack(ts[:3]))
    assert far == (1, 1, 1)
    # sizes that are not multiples of the tile, an empty batch, a zero threshold
    c = gpu.score_pnp_hypotheses(p3d[:1001], uv[:1001], cam, P[:17], 0.0).cpu().numpy()
    assert c.shape == (17,) and c.max() == 0
    c = gpu.score_pnp_hypotheses(p3d[:3], uv[:3], cam, P[:1], 1e9).cpu().numpy()
    assert c.tolist() == [3]
    assert gpu.score_pnp_hypotheses(p3d, uv, cam, np.zeros((0, 4, 4))).shape == (0,)


def test_pnp_scoring_equals_cv2_golden(gpu):
    """isr_pnp_score against OpenCV's own numbers (tests/golden/reference_pnp_cv2.npz): flags and
    counts equal the masks derived from cv2.projectPoints for every committed hypothesis."""
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_pnp_cv2.npz"))
    P = np.tile(np.eye(4), (len(g["Rs"]), 1, 1))
    P[:, :3, :3], P[:, :3, 3] = g["Rs"], g["tvecs"]
    counts, flags = gpu.score_pnp_hypotheses(g["p3d"], g["p2d"], g["cam"], P, 2.0, return_inliers=True)
    np.testing.assert_array_equal(flags.cpu().numpy().astype(bool), g["mask"])
    np.testing.assert_array_equal(counts.cpu().numpy(), g["mask"].sum(1))


def test_p3p_and_pnp_ransac_against_cv2_golden(gpu):
    """SURVEY 8(f) row 3, generation side.  (1) isr_p3p_solve reproduces every cv2.solveP3P solution
    of the committed 3-point sets and agrees with oracle.p3p.  (2) api.pnp_ransac on the committed
    scene (3000 correspondences, 25 % gross outliers, 0.7 px noise), called as choosePose.py:300
    calls cv2 (500 iterations, 2 px): the consensus of the returned pose -- counted with the
    cv2-pinned oracle -- is at least that of the pose cv2.solvePnPRansac returned, the pose is
    close to the truth, the inlier list is the winning hypothesis' consensus set, and the run
    is reproducible for a seed.  (3) helpers.pnp keeps the reference's call and return shape."""
    from imagesequenceregistrationfor6dposeestimationlabeling_b200 import api, helpers
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_pnp_cv2.npz"))
    K = g["cam"]
    P3 = np.stack([g["p3d"][ids].astype(np.float64) for ids in g["p3p_sets"]])
    UV = np.stack([g["uv_true"][ids] for ids in g["p3p_sets"]])
    poses, cnt = api.p3p_solve(P3, UV, K)
    poses, cnt = poses.cpu().numpy(), cnt.cpu().numpy()
    for b, (S, k) in enumerate(zip(g["p3p_solutions"], g["p3p_nsol"])):
        ours = [(poses[b, s, :3, :3], poses[b, s, :3, 3]) for s in range(cnt[b])]
        assert cnt[b] == len(oracle.p3p(P3[b], UV[b], K)) >= k
        assert np.isnan(poses[b, cnt[b]:]).all()
        for j in range(k):
            Rc, tc = S[j, :9].reshape(3, 3), S[j, 9:]
            assert min(np.abs(R - Rc).max() + np.abs(t - tc).max() / 100.0 for R, t in ours) < 1e-6
    cv_count = int(oracle.pnp_inliers(g["p3d"], g["p2d"], K, g["ransac_R"], g["ransac_t"], 2.0).sum())
    r = api.pnp_ransac(g["p3d"], g["p2d"], K, iterations=500, reprojection_error=2.0, seed=1)
    assert r.ok
    ours = oracle.pnp_inliers(g["p3d"], g["p2d"], K, r.R, r.t, 2.0)
    assert int(ours.sum()) == r.consensus >= cv_count          # at least cv2's consensus
    np.testing.assert_allclose(r.R @ r.R.T, np.eye(3), atol=1e-9)
    assert np.linalg.norm(r.R - g["R_true"]) < 2e-3 and np.linalg.norm(r.t - g["t_true"]) < 1.0
    assert len(r.inliers) <= r.consensus and len(r.inliers) > 0.6 * len(g["p3d"])
    assert ours[r.inliers].mean() > 0.97                       # the refit keeps the winner's inliers
    r2 = api.pnp_ransac(g["p3d"], g["p2d"], K, iterations=500, reprojection_error=2.0, seed=1)
    np.testing.assert_array_equal(r.R, r2.R)
    np.testing.assert_array_equal(r.inliers, r2.inliers)
    R, t, inl = helpers.pnp(g["p3d"], g["p2d"], K, itr=500, reperr=2)
    assert R.shape == (3, 3) and t.shape == (3,) and inl.ndim == 1
    # degenerate correspondences (every 3-D point the same: no P3P sample has a solution): the
    # reference's failure sentinel
    bad = helpers.pnp(g["p3d"][:50] * 0 + 1.0, g["p2d"][:50], K, itr=64, reperr=2)
    assert isinstance(bad, tuple) and bad == (1, 1, 1)


@pytest.mark.parametrize("n,k", [(1000, 400), (700, 16), (3000, 64), (40, 40)])
def test_knn_normals_equal_the_pytorch3d_restatement(gpu, n, k):
    """SURVEY 8(f) row 4, second half: isr_knn_normals against oracle.estimate_normals (the
    restated pytorch3d estimator) at the reference's size (1000 points, 400 neighbours,
    generateCors.py:200-215) and others: same direction to 1e-4 rad wherever the two smallest
    eigenvalues are separated, same sign, unit length; without disambiguation equal up to sign."""
    from imagesequenceregistrationfor6dposeestimationlabeling_b200 import synth
    pts = synth.make_cloud(n, seed=7)
    got = gpu.estimate_normals(pts, k).cpu().numpy().astype(np.float64)
    ref = oracle.estimate_normals(pts, k)
    np.testing.assert_allclose(np.linalg.norm(got, axis=1), 1.0, atol=1e-5)
    # eigen-gap of the reference neighbourhood covariances: directions are only comparable where it is open
    from scipy.spatial import cKDTree
    P = pts.astype(np.float64)
    _, idx = cKDTree(P).query(P, k=k)
    nb = P[idx.reshape(n, k)]
    cen = nb - nb.mean(1, keepdims=True)
    w = np.linalg.eigvalsh(np.einsum("nki,nkj->nij", cen, cen) / k)
    open_gap = (w[:, 1] - w[:, 0]) > 1e-3 * w[:, 2]
    assert open_gap.mean() > 0.9
    dots = (got * ref).sum(1)
    assert np.all(np.abs(dots[open_gap]) > 1 - 1e-6)
    # the majority-side rule decides the sign; only points whose vote is within one neighbour of k / 2 may differ
    proj = np.einsum("nki,ni->nk", nb - P[:, None, :], ref)
    margin = np.abs((proj > 0).sum(1) - 0.5 * k)
    assert np.all(dots[open_gap & (margin > 1.5)] > 0)
    got2 = gpu.estimate_normals(pts, k, disambiguate_directions=False).cpu().numpy().astype(np.float64)
    assert np.all(np.abs((got2 * ref).sum(1))[open_gap] > 1 - 1e-6)
    if k == n:   # every neighbourhood is the whole cloud: one normal for all
        assert np.all(np.abs(got2 @ got2[0]) > 1 - 1e-6)


def test_pointcloud_shim_transforms_large_clouds_on_the_device(gpu):
    """o3d_compat.PointCloud keeps float64 coordinates; a cloud of >= 50k points is transformed on
    the device (isr_transform_points_f64) and equals the float64 numpy transform to the last
    bit or two; transform returns self and is in place; `+`, distances and ICP then run on the
    device copy and agree with the host path (icp.py:22,110-117)."""
    import imagesequenceregistrationfor6dposeestimationlabeling_b200.o3d_compat as o3d
    from imagesequenceregistrationfor6dposeestimationlabeling_b200 import synth
    src, tgt, _ = synth.icp_pair(60000, 61000, 4, 5)
    R, t = synth.true_pose(3)
    T = synth.pose_matrix(R, t)
    big = o3d.geometry.PointCloud()
    big.points = o3d.utility.Vector3dVector(src.astype(np.float64))
    assert big.transform(T) is big and big._dev is not None and len(big) == len(src)
    want = src.astype(np.float64) @ R.T + t
    np.testing.assert_allclose(np.asarray(big.points), want, rtol=0, atol=2e-13 * 800)
    small = o3d.geometry.PointCloud()
    small.points = o3d.utility.Vector3dVector(src[:1000].astype(np.float64))
    small.transform(T)
    assert small._dev is None
    np.testing.assert_array_equal(np.asarray(small.points), want[:1000])
    # back to the target frame on the device, then the icp.py tail: ICP, transform, merge, distance
    big.transform(np.linalg.inv(T))
    np.testing.assert_allclose(np.asarray(big.points), src.astype(np.float64), atol=1e-9)
    target = o3d.geometry.PointCloud()
    target.points = o3d.utility.Vector3dVector(tgt)
    reg = o3d.pipelines.registration.registration_icp(
        big, target, 20, np.eye(4), o3d.pipelines.registration.TransformationEstimationPointToPoint())
    ref = gpu.icp(np.asarray(big.points), tgt, np.eye(4), 20.0)
    np.testing.assert_allclose(reg.transformation, ref.transformation, rtol=1e-9, atol=1e-9)
    merged = big.transform(reg.transformation) + target
    assert len(merged) == len(src) + len(tgt) and merged._dev is not None
    d = np.asarray(merged.compute_point_cloud_distance(target))
    assert d.shape == (len(src) + len(tgt),) and np.all(d[len(src):] == 0.0)
    host = np.concatenate([src.astype(np.float64) @ reg.transformation[:3, :3].T + reg.transformation[:3, 3],
                           tgt.astype(np.float64)])
    dk, _ = oracle.nearest(host, tgt)
    np.testing.assert_allclose(d, dk, rtol=1e-5, atol=1e-9)


def test_remove_radius_outlier_shim(gpu):
    """o3d.geometry.PointCloud.remove_radius_outlier as generateCors.py:254-258 calls it."""
    import imagesequenceregistrationfor6dposeestimationlabeling_b200.o3d_compat as o3d
    from imagesequenceregistrationfor6dposeestimationlabeling_b200 import synth
    rng = np.random.default_rng(8)
    mverts = np.concatenate([synth.make_cloud(5000, seed=2) / 66.0,            # unit-cube-sized surface
                             rng.uniform(-1.2, 1.2, size=(300, 3))]).astype(np.float32)  # stray points
    pcd = o3d.geometry.PointCloud()
    pcd.points = o3d.utility.Vector3dVector(mverts)
    cl, ind = pcd.remove_radius_outlier(nb_points=20, radius=0.05)
    ref_pts, ref_ind = oracle.remove_radius_outlier(mverts, 20, 0.05)
    assert list(ind) == list(ref_ind) and len(cl) == len(ref_ind)
    np.testing.assert_array_equal(np.asarray(cl.points), ref_pts.astype(np.float64))
    assert 0 < len(ind) < len(mverts)


def test_multistart_icp_matches_individual_runs(gpu):
    from imagesequenceregistrationfor6dposeestimationlabeling_b200 import synth
    src, tgt, _ = synth.icp_pair(8000, 9000, 6, 7)
    inits = np.stack([synth.pose_matrix(synth.rotvec_to_matrix([0, 0, 2 * np.pi * k / 8]), [0, 0, 0])
                      for k in range(8)])
    ms = gpu.multistart_icp(src, tgt, inits, 20.0, max_iteration=15)
    for k in (0, 3, 5):
        single = gpu.icp(src, tgt, inits[k], 20.0, max_iteration=15)
        np.testing.assert_array_equal(ms.results[k].transformation, single.transformation)
        assert ms.results[k].iterations == single.iterations
        o = oracle.registration_icp(src, tgt, 20.0, inits[k], max_iteration=15)
        _close_T(ms.results[k].transformation, o.transformation, rtol=1e-4)
    assert ms.order[0] == int(np.argmin(ms.chamfer))
    ch0 = oracle.chamfer(oracle.transform(src, ms.results[0].transformation), tgt)
    np.testing.assert_allclose(ms.chamfer[0], ch0, rtol=1e-5)


def test_choose_image_on_device_n64(gpu):
    """SURVEY 8(f) row 1 at n = 64 (4096 pose pairs): the relative-pose table built on the device
    equals the reference's loop (choosePose.py:43-51,98-107); ADD-S with the surface prepared
    once equals the general two-cloud form; error matrix, chosen image and top-50 equal
    oracle.choose_image (sklearn KD-tree per pair, choosePose.py:121-151) -- from the saved
    tables and from the pose lists directly."""
    from imagesequenceregistrationfor6dposeestimationlabeling_b200 import api, helpers, synth
    rng = np.random.default_rng(5)
    n = 64
    surface = synth.make_cloud(4000, seed=1).astype(np.float64)
    verts = synth.make_cloud(700, seed=3).astype(np.float64)
    gR = np.stack([synth.random_rotation(rng) for _ in range(n)])
    gt_ = rng.normal(scale=20.0, size=(n, 3)) + [0, 0, 700.0]
    # predictions: most within a few degrees / mm of the truth, a quarter badly off, some near the
    # 0.1 x diameter decision
    ang = np.where(rng.random(n) < 0.25, rng.uniform(0.5, 3.0, n), rng.uniform(0.0, 0.12, n))
    pR = np.stack([gR[k] @ synth.rotvec_to_matrix(ang[k] * synth.random_rotation(rng)[0]) for k in range(n)])
    pt = gt_ + rng.normal(scale=1.5, size=(n, 3))
    tab_p, tab_g = oracle.rel_pose_table(pR, pt), oracle.rel_pose_table(gR, gt_)
    dev_p = api.relative_pose_table(pR, pt).cpu().numpy().reshape(n, n, 4, 4)
    np.testing.assert_allclose(dev_p, tab_p, rtol=0, atol=1e-12)
    np.testing.assert_array_equal(dev_p[:, :, 3], np.tile([0, 0, 0, 1.0], (n, n, 1)))
    part = api.relative_pose_table(gR, gt_, pair0=n * 7 + 3, count=150).cpu().numpy()
    np.testing.assert_allclose(part, tab_g.reshape(-1, 4, 4)[n * 7 + 3:n * 7 + 153], rtol=0, atol=1e-12)
    # rigid form == general form
    P, G = tab_p.reshape(-1, 4, 4)[:300], tab_g.reshape(-1, 4, 4)[:300]
    a = api.adds_rigid(verts, G, P, surface).losses.cpu().numpy()
    b = api.verify_poses(verts, G, P, cloud_t=surface, mode="adds").losses.cpu().numpy()
    np.testing.assert_allclose(a, b, rtol=1e-6)
    # sphere bounds bracket the exact ADD-S (isr_adds_bounds), for aligned and for failed poses
    cen = api.centroid_of(surface.astype(np.float32))
    tgt = api.prepare_cloud(surface.astype(np.float32), centroid=cen, perm=api.spatial_order(surface.astype(np.float32)),
                            stage_centroids=True)
    lo, hi = api.adds_bounds(verts, api.rigid_relative(G, P), tgt)
    lo, hi = lo.cpu().numpy(), hi.cpu().numpy()
    assert np.all(lo <= a) and np.all(a <= hi) and np.all(lo >= 0)
    diameter = 120.0
    err_o, img_o, top_o = oracle.choose_image(tab_p, tab_g, verts, surface, diameter)
    assert 0 < err_o.sum() < n * n                       # both outcomes occur
    st1, st2 = {}, {}
    for err, img, top in (helpers.choose_image(tab_p, tab_g, verts, diameter, surface_points=surface, chunk=1000,
                                               stats=st1),
                          helpers.choose_image(tab_p, tab_g, verts, diameter, surface_points=surface, use_bounds=False),
                          helpers.choose_image_from_poses(pR, pt, gR, gt_, verts, diameter, surface_points=surface,
                                                          rows_per_chunk=5, stats=st2)):
        np.testing.assert_array_equal(err, err_o)
        assert img == img_o
        np.testing.assert_array_equal(top, top_o)
    # (with this 4000-point surface a sub-tile is a 13 mm patch and the bounds decide little; at 100k
    # points they decide most pairs: tests/test_gpu_fullsize.py::test_config2_adds_mode_100k_surface)
    assert st1["pairs"] == st2["pairs"] == n * n and st1["exact"] == st2["exact"] <= n * n
    # the vote kernel on its own: ties -> first maximum, NaN / inf never vote
    L = np.full((5, 7), 1.0)
    L[1, :3] = 0.0; L[3, :3] = 0.0; L[4, 0] = np.nan; L[4, 1] = np.inf; L[2, :] = 0.5
    e, v, best = api.vote(L, 0.5)
    np.testing.assert_array_equal(v.cpu().numpy(), [0, 3, 0, 3, 0])
    assert best.cpu().numpy().tolist() == [1, 3]
    np.testing.assert_array_equal(e.cpu().numpy(), (L < 0.5).astype(np.uint8))


def test_kdtree_shim_and_helpers(gpu):
    from imagesequenceregistrationfor6dposeestimationlabeling_b200 import helpers, synth
    from imagesequenceregistrationfor6dposeestimationlabeling_b200.compat import KDTree
    surface = synth.make_cloud(5000, seed=1).astype(np.float64)
    verts = synth.make_cloud(1200, seed=3).astype(np.float64)
    d, i = KDTree(surface, leaf_size=2).query(verts, k=1)
    dk, ik = oracle.nearest(verts.astype(np.float32), surface.astype(np.float32))
    assert d.shape == (1200, 1) and i.shape == (1200, 1) and i.dtype == np.int64
    np.testing.assert_array_equal(i[:, 0], ik)
    np.testing.assert_allclose(d[:, 0], dk, rtol=1e-5)
    P = _rand_poses(2, 3, trans=4.0)
    a = helpers.ADD(verts, P[0, :3, :3], P[0, :3, 3], P[1, :3, :3], P[1, :3, 3])
    np.testing.assert_allclose(a, oracle.ADD(verts, P[0, :3, :3], P[0, :3, 3], P[1, :3, :3], P[1, :3, 3]),
                               rtol=1e-5)
    helpers.surfacePointsScaled = surface
    s = helpers.ADDS(verts, P[0, :3, :3], P[0, :3, 3], P[1, :3, :3], P[1, :3, 3])
    np.testing.assert_allclose(
        s, oracle.ADDS(verts, P[0, :3, :3], P[0, :3, 3], P[1, :3, :3], P[1, :3, 3], surface), rtol=1e-5)
    helpers.surfacePointsScaled = None


def test_bad_arguments_raise(gpu):
    with pytest.raises(ValueError):
        gpu.nearest_neighbors(np.zeros((4, 2)), np.zeros((4, 3)))
    with pytest.raises(ValueError):
        gpu.nearest_neighbors(np.zeros((4, 3)), np.zeros((0, 3)))
    with pytest.raises(ValueError):
        gpu.verify_poses(np.zeros((4, 3)), np.zeros((2, 4, 4)), np.zeros((3, 4, 4)))
    from imagesequenceregistrationfor6dposeestimationlabeling_b200 import _lib
    lib = _lib.load()
    st = lib.isr_nn2(None, None, 1, 1, None, None, None, 0, None, 0, None)
    assert st < 0 and len(lib.isr_last_error()) > 0
