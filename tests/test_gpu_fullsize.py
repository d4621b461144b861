"""BASELINE.json configurations at (or near) full size.  The oracle checks a seeded subset
where the full comparison would take too long; size-independent properties cover the rest
(planted best candidate, permutation invariance, symmetry of duplicated candidates,
idempotence of a converged ICP, agreement between batched and single runs)."""
import numpy as np
import pytest

from oracle import c_oracle, oracle

pytestmark = pytest.mark.gpu


def test_config1_ruapc_pair_100k_icp30_chamfer(gpu):
    """Config 1 (the reference's own CPU-runnable case): 100k-pt upper/lower clouds,
    evaluate + 30-iteration ICP from the inverse predicted pose + final Chamfer -- fully
    against the oracle (icp.py:64-117)."""
    from imagesequenceregistrationfor6dposeestimationlabeling_b200 import api, synth
    upper = synth.make_cloud(100000, seed=1, half="upper")
    lower = synth.make_cloud(100000, seed=2, half="lower")
    cad = synth.make_cloud(50000, seed=3).astype(np.float64)
    R_GT, t_GT = synth.true_pose(3)
    R_pred = R_GT @ synth.rotvec_to_matrix(np.deg2rad(2.0) * np.array([0.6, 0.0, 0.8]))
    t_pred = t_GT + np.array([1.2, -1.0, 1.2])
    ev_o, reg_o, ch_o = oracle.icp_script(upper, lower, R_GT, t_GT, R_pred, t_pred, cad)
    actual_upper = upper.dot(R_GT.T) + t_GT
    init = np.linalg.inv(api.pose_from_Rt(R_pred, t_pred))
    ev = gpu.evaluate_registration(actual_upper, lower, 20.0, init)
    reg = gpu.icp(actual_upper, lower, init, 20.0)
    assert ev.n_corr == len(ev_o.correspondence_set) and ev.fitness == ev_o.fitness
    np.testing.assert_allclose(ev.inlier_rmse, ev_o.inlier_rmse, rtol=1e-9)
    assert reg.iterations == reg_o.iterations
    assert reg.n_corr == len(reg_o.correspondence_set) and reg.fitness == reg_o.fitness
    np.testing.assert_allclose(reg.transformation, reg_o.transformation, rtol=1e-7, atol=1e-7)
    np.testing.assert_allclose(reg.inlier_rmse, reg_o.inlier_rmse, rtol=1e-8)
    np.testing.assert_array_equal(reg.correspondence_set, reg_o.correspondence_set)
    T = reg.transformation
    merged = np.concatenate([actual_upper @ T[:3, :3].T + T[:3, 3], lower.astype(np.float64)])
    np.testing.assert_allclose(float(gpu.chamfer_distance(merged, cad)), ch_o, rtol=1e-6)


def test_config2_verification_1k_candidates_100k_points(gpu):
    """Config 2: 1000 candidates x 100k points.  Oracle on a seeded subset of 120 candidates
    (incl. the planted one and the GPU's top 8; ~12 s of KD-tree on the box's host cores);
    planted candidate selected; shuffling the candidates permutes the losses."""
    from imagesequenceregistrationfor6dposeestimationlabeling_b200 import synth
    cloud = synth.make_cloud(100000, seed=1)
    R_true, _ = synth.true_pose(3)
    Rs, _, k0 = synth.make_candidates(1000, seed=10, R_true=R_true, t_true=np.zeros(3))
    Mq, Mt = synth.verification_matrices(Rs, R_true)
    res = gpu.verify_poses(cloud, Mq, Mt, mode="chamfer")
    losses = res.losses.cpu().numpy()
    assert res.best_index == k0 == int(np.argmin(losses))
    rng = np.random.default_rng(0)
    top8 = np.argsort(losses, kind="stable")[:8]
    sub = np.unique(np.concatenate([[k0], top8, rng.choice(1000, size=120, replace=False)]))
    assert len(sub) >= 100
    ref, _ = oracle.verify_matrices(cloud, cloud, Mq[sub], Mt[sub])
    np.testing.assert_allclose(losses[sub], ref, rtol=1e-5)
    assert int(sub[np.argmin(ref)]) == k0
    perm = rng.permutation(1000)[:64]
    res2 = gpu.verify_poses(cloud, Mq[perm], Mt[perm], mode="chamfer")
    np.testing.assert_array_equal(res2.losses.cpu().numpy(), losses[perm])   # bit-identical


def test_config3_sweep_10k_candidates_100k_points(gpu):
    """Config 3 on one GPU (the multi-GPU form shards the same call by candidate,
    tests/test_dist_gpu.py): 10 000 candidates x 100k points.  Planted candidate selected;
    the oracle checks a seeded subset of 256 candidates plus the planted one plus the GPU's top 8
    (~30 s of KD-tree on the box's host cores); a block of the sweep
    scored on its own gives bit-identical losses (chunking and launch order do not leak into
    results)."""
    from imagesequenceregistrationfor6dposeestimationlabeling_b200 import synth
    cloud = synth.make_cloud(100000, seed=1)
    R_true, _ = synth.true_pose(3)
    Rs, _, k0 = synth.make_candidates(10000, seed=11, R_true=R_true, t_true=np.zeros(3))
    Mq, Mt = synth.verification_matrices(Rs, R_true)
    res = gpu.verify_poses(cloud, Mq, Mt, mode="chamfer")
    losses = res.losses.cpu().numpy()
    assert losses.shape == (10000,) and np.isfinite(losses).all()
    assert res.best_index == k0 == int(np.argmin(losses))
    rng = np.random.default_rng(3)
    top8 = np.argsort(losses, kind="stable")[:8]
    sub = np.unique(np.concatenate([[k0], top8, rng.choice(10000, size=256, replace=False)]))
    assert len(sub) >= 256
    ref, _ = oracle.verify_matrices(cloud, cloud, Mq[sub], Mt[sub])
    np.testing.assert_allclose(losses[sub], ref, rtol=1e-5)
    assert int(sub[np.argmin(ref)]) == k0
    blk = slice(4100, 4100 + 300)
    res2 = gpu.verify_poses(cloud, Mq[blk], Mt[blk], mode="chamfer")
    np.testing.assert_array_equal(res2.losses.cpu().numpy(), losses[blk])


def test_config2_adds_mode_100k_surface(gpu):
    from imagesequenceregistrationfor6dposeestimationlabeling_b200 import synth
    surface = synth.make_cloud(100000, seed=1)
    verts = synth.make_cloud(20000, seed=3)
    R_true, t_true = synth.true_pose(3)
    Rs, ts, k0 = synth.make_candidates(48, seed=10, R_true=R_true, t_true=t_true)
    got = gpu.adds(verts, np.tile(R_true, (48, 1, 1)), np.tile(t_true, (48, 1)), Rs, ts, surface).cpu().numpy()
    best = int(np.argmin(got))
    for k in sorted({0, k0, 17, 47, best}):
        ref = oracle.ADDS(verts.astype(np.float64), R_true, t_true, Rs[k], ts[k], surface.astype(np.float64))
        np.testing.assert_allclose(got[k], ref, rtol=1e-5)
    assert np.all(np.isfinite(got)) and got.shape == (48,)
    # the same pairs with the surface prepared once (the vote's path) and their sphere bounds
    from imagesequenceregistrationfor6dposeestimationlabeling_b200 import api
    Pg = np.stack([synth.pose_matrix(R_true, t_true)] * 48)
    Pp = np.stack([synth.pose_matrix(Rs[k], ts[k]) for k in range(48)])
    rig = api.adds_rigid(verts, Pg, Pp, surface).losses.cpu().numpy()
    np.testing.assert_allclose(rig, got, rtol=1e-6)
    cen = api.centroid_of(surface)
    tgt = api.prepare_cloud(surface, centroid=cen, perm=api.spatial_order(surface), stage_centroids=True)
    lo, hi = api.adds_bounds(verts, api.rigid_relative(Pg, Pp), tgt)
    lo, hi = lo.cpu().numpy(), hi.cpu().numpy()
    assert np.all(lo <= got) and np.all(got <= hi)
    assert np.mean((hi < 12.0) | (lo >= 12.0)) > 0.5      # most pairs are decided at 0.1 x 120 mm


def test_nn_100k_x_100k_exact_indices_sampled(gpu):
    """Full-size search; 4000 seeded queries checked against the float64 brute force, and the
    whole result against the self-consistency property d2[i] == |q_i - t_idx[i]|^2."""
    from imagesequenceregistrationfor6dposeestimationlabeling_b200 import synth
    t = synth.make_cloud(100000, seed=1)
    q = (synth.make_cloud(100000, seed=2).astype(np.float64) @ synth.rotvec_to_matrix([0.01, 0.02, -0.01]).T
         ).astype(np.float32)
    res = gpu.nearest_neighbors(q, t)
    idx, d2 = res.idx.cpu().numpy(), res.d2.cpu().numpy()
    rng = np.random.default_rng(1)
    s = rng.choice(len(q), size=4000, replace=False)
    rd2, ridx = c_oracle.nn_f64(q[s], t)
    np.testing.assert_array_equal(idx[s], ridx)
    own = ((q.astype(np.float64) - t[idx].astype(np.float64)) ** 2).sum(1)
    np.testing.assert_allclose(d2, own, rtol=1.2e-7)
    assert idx.min() >= 0 and idx.max() < len(t)


def test_config4_dense_icp_1m_x_1m(gpu):
    """Config 4: 1M x 1M, forced iterations (criteria 0).  Three iterations against the
    oracle (1M-point KD-tree), then the GPU's 50-iteration result must be a fixed point:
    one more evaluation from it reproduces fitness / rmse, and fitness stays 1."""
    from imagesequenceregistrationfor6dposeestimationlabeling_b200 import synth
    src, tgt, Tm = synth.icp_pair(1000000, 1000000, 4, 5)
    r3 = gpu.icp(src, tgt, np.eye(4), 20.0, max_iteration=3, relative_fitness=0.0, relative_rmse=0.0)
    o3 = oracle.registration_icp(src, tgt, 20.0, np.eye(4), max_iteration=3, relative_fitness=0.0,
                                 relative_rmse=0.0)
    assert r3.iterations == o3.iterations == 3 and r3.fitness == o3.fitness
    np.testing.assert_allclose(r3.transformation, o3.transformation, rtol=1e-8, atol=1e-8)
    np.testing.assert_allclose(r3.inlier_rmse, o3.inlier_rmse, rtol=1e-9)
    r50 = gpu.icp(src, tgt, np.eye(4), 20.0, max_iteration=50, relative_fitness=0.0, relative_rmse=0.0)
    assert r50.iterations == 50 and r50.fitness == 1.0
    ev = gpu.evaluate_registration(src, tgt, 20.0, r50.transformation)
    assert ev.fitness == r50.fitness
    np.testing.assert_allclose(ev.inlier_rmse, r50.inlier_rmse, rtol=1e-12)
    assert r50.inlier_rmse <= r3.inlier_rmse
    # point-to-point ICP slides slowly along the surface: after 50 iterations the synthetic
    # motion is only partly undone, but strictly better than at the start and than after 3
    err = lambda T: np.linalg.norm(T[:3, :3] @ Tm[:3, :3] - np.eye(3))
    assert err(r50.transformation) < err(r3.transformation) < err(np.eye(4))


def test_config5_multistart_64_x_250k(gpu):
    """Config 5: 64 symmetry-seeded starts x 250k points, batched, Chamfer-ranked.  Batched
    starts equal single runs bit for bit; two starts are checked against the oracle."""
    from imagesequenceregistrationfor6dposeestimationlabeling_b200 import synth
    src, tgt, _ = synth.icp_pair(250000, 250000, 6, 7)
    inits = np.stack([synth.pose_matrix(synth.rotvec_to_matrix([0, 0, 2 * np.pi * k / 64]), [0, 0, 0])
                      for k in range(64)])
    ms = gpu.multistart_icp(src, tgt, inits, 20.0, max_iteration=30)
    assert len(ms.results) == 64 and ms.order[0] == int(np.argmin(ms.chamfer))
    assert ms.order[0] in (0, 32) or ms.chamfer[ms.order[0]] <= ms.chamfer[0]
    for k in (0, 21):
        single = gpu.icp(src, tgt, inits[k], 20.0, max_iteration=30)
        np.testing.assert_array_equal(ms.results[k].transformation, single.transformation)
        assert ms.results[k].iterations == single.iterations
    o = oracle.registration_icp(src, tgt, 20.0, inits[0], max_iteration=30)
    assert ms.results[0].iterations == o.iterations
    np.testing.assert_allclose(ms.results[0].transformation, o.transformation, rtol=1e-7, atol=1e-7)
    ch0 = oracle.chamfer(oracle.transform(src, ms.results[0].transformation), tgt)
    np.testing.assert_allclose(ms.chamfer[0], ch0, rtol=1e-5)
