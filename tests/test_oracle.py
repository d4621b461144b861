"""The CPU oracle against the reference's golden vectors and hand-computable known answers.
(No GPU.)  SURVEY.md section 8(c) lists the KATs; the reference itself has no tests."""
import os

import numpy as np
import pytest

from imagesequenceregistrationfor6dposeestimationlabeling_b200 import synth
from oracle import c_oracle, oracle

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_helpers.npz"))


# ---- pinned by the reference's own functions (tests/golden/make_golden.py) ---------------
def test_golden_add_adds():
    n = len(GOLD["R"])
    for i in range(n):
        for j in range(n):
            a = oracle.ADD(GOLD["verts"], GOLD["R"][i], GOLD["t"][i], GOLD["R"][j], GOLD["t"][j])
            s = oracle.ADDS(GOLD["verts"], GOLD["R"][i], GOLD["t"][i], GOLD["R"][j], GOLD["t"][j],
                            GOLD["surface"])
            assert a == GOLD["ADD"][i, j]
            assert s == GOLD["ADDS"][i, j]
    for k in range(5):
        s = oracle.ADDS(GOLD["verts"], GOLD["R"][0], GOLD["t"][0], GOLD["Rp"][k], GOLD["tp"][k],
                        GOLD["surface"])
        assert s == GOLD["ADDS_p"][k]


def test_golden_relative_poses():
    n = len(GOLD["R"])
    for i in range(n):
        for j in range(n):
            r, t = oracle.compute_rel_poses(GOLD["R"][i], GOLD["t"][i], GOLD["R"][j], GOLD["t"][j])
            np.testing.assert_array_equal(r, GOLD["rel_R"][i, j])
            np.testing.assert_array_equal(t, GOLD["rel_t"][i, j])
            r, t = oracle.calculate_relative_pose(GOLD["R"][i], GOLD["t"][i], GOLD["R"][j], GOLD["t"][j])
            np.testing.assert_allclose(r, GOLD["cal_R"][i, j], rtol=0, atol=1e-12)
            np.testing.assert_allclose(t, GOLD["cal_T"][i, j], rtol=0, atol=1e-9)
    tab = oracle.rel_pose_table(GOLD["R"], GOLD["t"])
    np.testing.assert_array_equal(tab[:, :, :3, :3], GOLD["rel_R"])
    np.testing.assert_array_equal(tab[:, :, :3, 3], GOLD["rel_t"])


def test_adds_kdtree_equals_bruteforce_definition():
    """sklearn KDTree(leaf_size=2) 1-NN == brute-force float64 definition."""
    S = GOLD["surface"].dot(GOLD["Rp"][2].T) + GOLD["tp"][2]
    V = GOLD["verts"].dot(GOLD["R"][0].T) + GOLD["t"][0]
    d2, _ = c_oracle.nn_f64(V, S)
    np.testing.assert_allclose(np.sqrt(d2).mean(), GOLD["ADDS_p"][2], rtol=1e-13)


# ---- KAT 1-3: nearest neighbour ------------------------------------------------------------
def test_identical_clouds():
    a = synth.make_cloud(3000, seed=1)
    d, i = oracle.nearest(a, a)
    assert np.all(d == 0)
    np.testing.assert_array_equal(i, np.arange(len(a)))
    assert oracle.chamfer(a, a) == 0.0


def test_small_shift_keeps_identity_correspondence():
    rng = np.random.default_rng(0)
    a = rng.uniform(-50, 50, size=(500, 3)).round(0)          # >= 1 apart (or duplicate)
    a = np.unique(a, axis=0)
    delta = np.array([0.2, -0.1, 0.2])
    d, i = oracle.nearest(a + delta, a)
    np.testing.assert_array_equal(i, np.arange(len(a)))
    np.testing.assert_allclose(d, np.linalg.norm(delta), rtol=1e-12)
    np.testing.assert_allclose(oracle.chamfer(a + delta, a), np.linalg.norm(delta), rtol=1e-12)


def test_lattice_ties_lowest_index_bruteforce():
    g = np.stack(np.meshgrid(np.arange(4), np.arange(4), np.arange(4), indexing="ij"), -1).reshape(-1, 3)
    t = np.concatenate([g, g]).astype(np.float32)
    q = (g + 0.5).astype(np.float32)
    d2, idx = c_oracle.nn_f64(q, t)
    d2f, idxf = c_oracle.nn_f32_fma(q, t)
    np.testing.assert_array_equal(idx, idxf)
    assert np.all(idx < len(g)) and np.all(d2 == 0.75) and np.all(d2f == 0.75)
    # lowest index among the 8 corners of each cell = the corner at the cell origin
    for k in (0, 5, 21):
        cell = g[k]
        if np.all(cell < 3):
            assert idx[k] == k


def test_ckdtree_equals_bruteforce_on_surface_clouds():
    """Pins the scipy stand-in for Open3D's nanoflann search against the definition."""
    t = synth.make_cloud(30000, seed=1)
    q = synth.make_cloud(4000, seed=2)
    for off in (0.0, 700.0):
        d, i = oracle.nearest(q + np.float32(off), t + np.float32(off))
        d2, ib = c_oracle.nn_f64(q + np.float32(off), t + np.float32(off))
        np.testing.assert_array_equal(i, ib)
        np.testing.assert_allclose(d, np.sqrt(d2), rtol=1e-14)
        # SURVEY section 7 probe: the float32 direct-difference form finds the same neighbours
        _, i32 = c_oracle.nn_f32_fma(q + np.float32(off), t + np.float32(off))
        assert np.mean(i32 == ib) > 0.9995


def test_empty_target_and_source():
    a = synth.make_cloud(10, seed=1)
    assert np.all(oracle.compute_point_cloud_distance(a, np.zeros((0, 3))) == 0)
    assert len(oracle.compute_point_cloud_distance(np.zeros((0, 3)), a)) == 0


# ---- KAT 4: Kabsch ------------------------------------------------------------------------
def test_umeyama_recovers_exact_motion_and_fixes_reflection():
    rng = np.random.default_rng(5)
    src = rng.normal(scale=40, size=(200, 3))
    R = synth.random_rotation(rng)
    t = np.array([3.0, -700.0, 12.5])
    T = oracle.umeyama(src, src @ R.T + t)
    np.testing.assert_allclose(T[:3, :3], R, atol=1e-12)
    np.testing.assert_allclose(T[:3, 3], t, atol=1e-9)
    # a mirrored target must still yield a proper rotation
    M = np.diag([1.0, 1.0, -1.0])
    T = oracle.umeyama(src, src @ M.T)
    assert abs(np.linalg.det(T[:3, :3]) - 1.0) < 1e-12
    np.testing.assert_allclose(T[:3, :3] @ T[:3, :3].T, np.eye(3), atol=1e-12)


def test_umeyama_equals_an_independent_kabsch():
    """The Open3D half of the oracle has no golden vectors (Open3D is not installable here), so
    its rigid estimation is also checked against an implementation that shares no code with
    it: scipy's Rotation.align_vectors (Kabsch on centred vectors) on noisy correspondences."""
    from scipy.spatial.transform import Rotation
    rng = np.random.default_rng(11)
    for n, noise in ((50, 0.0), (400, 0.5), (3000, 3.0)):
        src = rng.normal(scale=40, size=(n, 3))
        R = synth.random_rotation(rng)
        dst = src @ R.T + np.array([5.0, -2.0, 700.0]) + rng.normal(scale=noise, size=(n, 3))
        T = oracle.umeyama(src, dst)
        rot, _ = Rotation.align_vectors(dst - dst.mean(0), src - src.mean(0))
        np.testing.assert_allclose(T[:3, :3], rot.as_matrix(), atol=1e-9)
        np.testing.assert_allclose(T[:3, 3], dst.mean(0) - rot.as_matrix() @ src.mean(0), atol=1e-7)


def test_evaluate_registration_equals_bruteforce_definition():
    """fitness / inlier_rmse / correspondence set straight from their definitions with the C
    brute-force nearest neighbour (no tree): strict d < threshold, one row per inlier."""
    src, tgt, _ = synth.icp_pair(700, 900, 6, 7, angle_deg=4.0, shift_mm=6.0)
    T = synth.pose_matrix(synth.rotvec_to_matrix([0.02, -0.01, 0.03]), [0.5, -0.3, 0.2])
    thr = 2.0
    r = oracle.evaluate_registration(src, tgt, thr, T)
    moved = (src.astype(np.float64) @ T[:3, :3].T + T[:3, 3])
    d2 = ((moved[:, None, :] - tgt.astype(np.float64)[None, :, :]) ** 2).sum(-1)
    j = d2.argmin(1)
    dmin = np.sqrt(d2[np.arange(len(src)), j])
    keep = dmin < thr
    assert 0 < keep.sum() < len(src)
    assert r.fitness == keep.sum() / len(src)
    np.testing.assert_allclose(r.inlier_rmse, np.sqrt((dmin[keep] ** 2).mean()), rtol=1e-12)
    np.testing.assert_array_equal(np.asarray(r.correspondence_set),
                                  np.stack([np.nonzero(keep)[0], j[keep]], axis=1))


# ---- KAT 5-7: ICP -------------------------------------------------------------------------
def test_icp_recovers_small_motion():
    tgt = synth.make_cloud(8000, seed=8)
    Tm = synth.pose_matrix(synth.rotvec_to_matrix([0.004, 0.003, -0.005]), [0.05, -0.04, 0.03])
    src = oracle.transform(tgt, Tm)
    r = oracle.registration_icp(src, tgt, 20.0, np.eye(4))
    assert r.fitness == 1.0 and r.iterations < 30
    np.testing.assert_allclose(r.transformation, np.linalg.inv(Tm), atol=1e-9)
    assert r.inlier_rmse < 1e-9


def test_icp_threshold_is_strict_and_empty_set_is_identity():
    tgt = np.array([[0, 0, 0], [100, 0, 0]], dtype=np.float64)
    src = np.array([[0, 0, 20.0], [100, 0, 20.0 * (1 - 1e-6)]])
    r = oracle.evaluate_registration(src, tgt, 20.0)
    assert r.fitness == 0.5 and len(r.correspondence_set) == 1
    np.testing.assert_array_equal(r.correspondence_set, [[1, 1]])
    r = oracle.registration_icp(src + 1000.0, tgt, 20.0, np.eye(4))
    assert r.fitness == 0 and r.inlier_rmse == 0 and r.iterations == 1
    np.testing.assert_array_equal(r.transformation, np.eye(4))
    r = oracle.evaluate_registration(src, tgt, 0.0)
    assert r.fitness == 0 and len(r.correspondence_set) == 0


def test_icp_max_iteration_zero_is_evaluation():
    src, tgt, _ = synth.icp_pair(2000, 2500, 6, 7)
    a = oracle.registration_icp(src, tgt, 20.0, np.eye(4), max_iteration=0)
    b = oracle.evaluate_registration(src, tgt, 20.0, np.eye(4))
    assert a.fitness == b.fitness and a.inlier_rmse == b.inlier_rmse and a.iterations == 0


# ---- KAT 8-9: pose algebra and selection ----------------------------------------------------
def test_relative_pose_identities():
    rng = np.random.default_rng(1)
    R, t = synth.random_rotation(rng), rng.normal(size=3)
    r, tt = oracle.calculate_relative_pose(R, t, R, t)
    np.testing.assert_allclose(r, np.eye(3), atol=1e-14)
    np.testing.assert_allclose(tt, 0, atol=1e-12)
    R2, t2 = synth.random_rotation(rng), rng.normal(size=3)
    r, tt = oracle.compute_rel_poses(R, t, R2, t2)
    np.testing.assert_array_equal(r, R.T @ R2)
    np.testing.assert_array_equal(tt, t2 - t)


def test_verify_selects_first_occurrence_of_true_pose():
    cloud = synth.make_cloud(2000, seed=1)
    R_true, _ = synth.true_pose(3)
    Rs, _, k0 = synth.make_candidates(10, seed=10, R_true=R_true, t_true=np.zeros(3))
    Rs = np.concatenate([Rs, Rs[k0:k0 + 1]])  # duplicate of the best candidate at the end
    Mq, Mt = synth.verification_matrices(Rs, R_true)
    losses, best = oracle.verify_matrices(cloud, cloud, Mq, Mt)
    assert best == k0 and losses[k0] == losses[-1]


def test_verify_loop_form_equals_matrix_form():
    pc1 = synth.make_cloud(1500, seed=4)
    rng = np.random.default_rng(0)
    n = 5
    gt_R = [synth.random_rotation(rng) for _ in range(n)]
    gt_T = [rng.normal(scale=50, size=3) for _ in range(n)]
    pred_R = [gt_R[i] @ synth.rotvec_to_matrix(rng.normal(scale=0.05, size=3)) for i in range(n)]
    pred_T = gt_T
    ref, idx, mn = oracle.verify_chamfer(pc1, gt_R, gt_T, pred_R, pred_T)
    Mq = np.tile(np.eye(4), (n - 1, 1, 1))
    Mt = np.tile(np.eye(4), (n - 1, 1, 1))
    for i in range(n - 1):
        R_rel, _ = oracle.calculate_relative_pose(gt_R[i], gt_T[i], gt_R[i + 1], gt_T[i + 1])
        Mq[i, :3, :3] = pred_R[i + 1].T
        Mt[i, :3, :3] = R_rel.T @ pred_R[i]
    losses, best = oracle.verify_matrices(pc1, pc1, Mq, Mt)
    np.testing.assert_allclose(losses, ref, rtol=1e-12)
    assert best == idx


def test_choose_image_vote():
    rng = np.random.default_rng(3)
    surface = synth.make_cloud(800, seed=1).astype(np.float64)
    verts = synth.make_cloud(200, seed=3).astype(np.float64)
    n = 4
    R = [synth.random_rotation(rng) for _ in range(n)]
    t = [rng.normal(scale=5, size=3) for _ in range(n)]
    gt = oracle.rel_pose_table(R, t)
    tp = [t[0], t[1], t[2] + np.array([0, 0, 80.0]), t[3]]  # image 2's prediction is wrong
    pred = oracle.rel_pose_table(R, tp)
    err, image_id, top = oracle.choose_image(pred, gt, verts, surface, diameter=120.0)
    assert err.shape == (n, n) and err[0, 1] == 1 and err[0, 2] == 0
    assert image_id in (0, 1, 3) and top[0] == image_id and top[-1] == 2


def test_radius_count_restatement_against_kdtree():
    """The brute-force definition agrees with scipy's ball query away from the boundary, is
    strict at the boundary, and counts the point itself."""
    from scipy.spatial import cKDTree
    rng = np.random.default_rng(3)
    pts = rng.normal(size=(1500, 3)).astype(np.float32)
    cnt = oracle.radius_count(pts, 0.3)
    ref = cKDTree(pts.astype(np.float64)).query_ball_point(pts.astype(np.float64), 0.3, return_length=True)
    np.testing.assert_array_equal(cnt, ref)          # no pair sits exactly on the boundary here
    line = np.array([[0, 0, 0], [1, 0, 0], [2, 0, 0], [4, 0, 0]], dtype=np.float32)
    np.testing.assert_array_equal(oracle.radius_count(line, 1.0), [1, 1, 1, 1])       # strict <
    np.testing.assert_array_equal(oracle.radius_count(line, 1.0000001), [2, 3, 2, 1])
    kept, ind = oracle.remove_radius_outlier(line, 1, 1.5)
    assert list(ind) == [0, 1, 2] and len(kept) == 3


def test_pnp_inliers_known_answers():
    """oracle.pnp_inliers: exact projection -> all inliers; a 3-pixel shift -> none at the
    2-pixel threshold, all at 3.5; the squared-error edge is inclusive; z == 0 projects with z = 1."""
    rng = np.random.default_rng(0)
    P = rng.uniform(-50, 50, size=(200, 3)).astype(np.float32)
    R, t = np.eye(3), np.array([0.0, 0.0, 700.0])
    cam = np.array([[1000.0, 0.0, 320.0], [0.0, 1000.0, 240.0], [0.0, 0.0, 1.0]])
    pc = P.astype(np.float64) + t
    uv = (pc[:, :2] / pc[:, 2:3]) * 1000.0 + [320.0, 240.0]
    assert oracle.pnp_inliers(P, uv, cam, R, t, 2.0).all()
    shifted = uv + [3.0, 0.0]
    assert not oracle.pnp_inliers(P, shifted, cam, R, t, 2.0).any()
    assert oracle.pnp_inliers(P, shifted, cam, R, t, 3.5).all()
    # inclusive edge on an exactly representable case: point on the axis, 2-pixel offset
    P0 = np.array([[0.0, 0.0, 0.0]], dtype=np.float32)
    assert oracle.pnp_inliers(P0, np.array([[322.0, 240.0]]), cam, R, t, 2.0)[0]
    assert not oracle.pnp_inliers(P0, np.array([[322.0, 240.0]]), cam, R, t, 1.999)[0]
    # z == 0: OpenCV divides by 1 instead
    m = oracle.pnp_inliers(np.array([[0.001, 0.0, 0.0]], dtype=np.float32), np.array([[321.0, 240.0]]), cam, R,
                           np.zeros(3), 0.5)
    assert m[0]


def test_pnp_inliers_match_cv2_golden():
    """The PnP half is PINNED by OpenCV itself (tests/golden/make_golden_cv2.py, cv2 4.13): for 8
    hypotheses x 3000 correspondences the consensus masks equal those derived from
    cv2.projectPoints' float32 output -- the array PnPRansacCallback::computeError compares
    against -- whether the squared error is accumulated in float32 (OpenCV's Matx21f norm) or
    rounded once from float64 (the restatement); and the pose cv2.solvePnPRansac returned for
    the reference's own call (choosePose.py:23-33: P3P, 500 iterations, 2 px) is scored
    consistently with the inlier list it returned."""
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_pnp_cv2.npz"))
    d = g["p2d"][None] - g["proj"]
    err32 = (d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1]).astype(np.float32)
    d64 = d.astype(np.float64)
    err64 = (d64[..., 0] ** 2 + d64[..., 1] ** 2).astype(np.float32)
    np.testing.assert_array_equal(err32 <= np.float32(4.0), g["mask"])
    np.testing.assert_array_equal(err64.astype(np.float64) <= 4.0, g["mask"])
    for k in range(len(g["Rs"])):
        m = oracle.pnp_inliers(g["p3d"], g["p2d"], g["cam"], g["Rs"][k], g["tvecs"][k], 2.0)
        np.testing.assert_array_equal(m, g["mask"][k])
    assert g["mask"][0].sum() > 2000 and g["mask"][5].sum() == 0
    # cv2's RANSAC result: its inlier list belongs to the best MINIMAL model, the returned pose is
    # the EPnP refit on those inliers -- so the refit's consensus contains nearly all of them
    m = oracle.pnp_inliers(g["p3d"], g["p2d"], g["cam"], g["ransac_R"], g["ransac_t"], 2.0)
    inl = g["ransac_inliers"]
    assert m[inl].mean() > 0.99 and m.sum() >= len(inl)


def test_p3p_restatement_matches_cv2_solveP3P():
    """oracle.p3p (Grunert) returns every solution cv2.solveP3P returned for the 24 committed
    3-point sets (tests/golden/make_golden_cv2.py), and each of its solutions reprojects the
    three points exactly."""
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_pnp_cv2.npz"))
    K = g["cam"]
    for ids, S, k in zip(g["p3p_sets"], g["p3p_solutions"], g["p3p_nsol"]):
        P = g["p3d"][ids].astype(np.float64)
        uv = g["uv_true"][ids]
        sols = oracle.p3p(P, uv, K)
        assert len(sols) >= k >= 1
        for j in range(k):
            Rc, tc = S[j, :9].reshape(3, 3), S[j, 9:]
            d = min(np.abs(R - Rc).max() + np.abs(t - tc).max() / 100.0 for R, t in sols)
            assert d < 1e-6
        for R, t in sols:
            pc = P @ R.T + t
            pr = (pc / pc[:, 2:3]) @ K.T
            np.testing.assert_allclose(pr[:, :2], uv, atol=1e-6)
            np.testing.assert_allclose(R @ R.T, np.eye(3), atol=1e-9)


def test_estimate_normals_restatement_on_a_sphere():
    """oracle.estimate_normals: on a noiseless sphere the k-NN PCA normal is radial; the
    majority-side rule points it INWARDS (the neighbours of a convex patch lie below the tangent
    plane), which is why generateCors.py:208 negates it."""
    rng = np.random.default_rng(0)
    u = rng.normal(size=(1500, 3))
    u /= np.linalg.norm(u, axis=1, keepdims=True)
    n = oracle.estimate_normals(50.0 * u, 40)
    np.testing.assert_allclose(np.linalg.norm(n, axis=1), 1.0, atol=1e-12)
    assert np.all((n * u).sum(1) < -0.99)
    n2 = oracle.estimate_normals(50.0 * u, 40, disambiguate_directions=False)
    assert np.all(np.abs((n2 * u).sum(1)) > 0.99)


# ---- size-independent properties of the restated path (seeded, small) ----------------------
@pytest.mark.parametrize("seed", [0, 1, 2])
def test_chamfer_is_symmetric_and_rigid_invariant(seed):
    """Chamfer(a, b) = Chamfer(b, a); moving BOTH clouds by the same rigid motion changes
    nothing (to rounding) -- what lets verification work in any frame."""
    rng = np.random.default_rng(seed)
    a = synth.make_cloud(900, seed=20 + seed).astype(np.float64)
    b = synth.make_cloud(1100, seed=30 + seed).astype(np.float64) + rng.normal(scale=0.5, size=3)
    c0 = oracle.chamfer(a, b)
    assert c0 == oracle.chamfer(b, a)
    R, t = synth.random_rotation(rng), rng.normal(scale=300, size=3)
    np.testing.assert_allclose(oracle.chamfer(a @ R.T + t, b @ R.T + t), c0, rtol=1e-10)
    assert oracle.chamfer(a, a) == 0.0


def test_icp_result_is_a_fixed_point_of_evaluation():
    """The registration result is the evaluation AFTER the last update (upstream's loop):
    evaluating its transformation again reproduces fitness, rmse and the correspondence set,
    and restarting ICP from it converges at once without moving."""
    src, tgt, _ = synth.icp_pair(1500, 1800, 6, 7)
    r = oracle.registration_icp(src, tgt, 20.0, np.eye(4), max_iteration=40)
    e = oracle.evaluate_registration(src, tgt, 20.0, r.transformation)
    # (upstream moves the cloud by each update in turn; one transform by the composed matrix
    # differs in the last bits)
    assert e.fitness == r.fitness
    np.testing.assert_allclose(e.inlier_rmse, r.inlier_rmse, rtol=1e-12)
    np.testing.assert_array_equal(np.asarray(e.correspondence_set), np.asarray(r.correspondence_set))
    if r.iterations < 40:  # stopped by the criteria: one more run stops after its first update
        r2 = oracle.registration_icp(src, tgt, 20.0, r.transformation, max_iteration=40)
        assert r2.iterations <= 1
        np.testing.assert_allclose(r2.transformation, r.transformation, atol=1e-5)


def test_verify_is_permutation_equivariant_and_first_min():
    """Shuffling the candidates permutes the losses; with the best candidate duplicated the
    FIRST occurrence is selected (list.index(min(list)), verfication.py:105-106)."""
    cloud = synth.make_cloud(1200, seed=1)
    R_true, _ = synth.true_pose(3)
    Rs, _, k0 = synth.make_candidates(7, seed=10, R_true=R_true, t_true=np.zeros(3))
    Mq, Mt = synth.verification_matrices(Rs, R_true)
    losses, best = oracle.verify_matrices(cloud, cloud, Mq, Mt)
    perm = np.random.default_rng(0).permutation(7)
    lp, bp = oracle.verify_matrices(cloud, cloud, Mq[perm], Mt[perm])
    np.testing.assert_array_equal(lp, losses[perm])
    assert perm[bp] == best == k0
    Mq2, Mt2 = np.concatenate([Mq, Mq[k0:k0 + 1]]), np.concatenate([Mt, Mt[k0:k0 + 1]])
    _, b2 = oracle.verify_matrices(cloud, cloud, Mq2, Mt2)
    assert b2 == k0
