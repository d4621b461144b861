"""Seeded synthetic point clouds and candidate poses.

The reference ships no data (bop/*/dataset.zip are empty), so every parity test and
bench line runs on these generators.  Shapes and magnitudes follow the reference's
input contract: NeRF surface clouds of <= 80k float32 points in millimetres
(genFeat.py:199-228), object diameter ~120 mm, camera distance ~700 mm (T-LESS /
RU-APC range), two overlapping half-object clouds for ICP (icp.py:47-50).

numpy only -- this module runs on the host and is shared by tests, bench and smoke.
"""
from __future__ import annotations

import numpy as np

SEMI_AXES = np.array([1.0, 0.7, 0.5])


def make_cloud(n: int, seed: int, diameter: float = 120.0, sigma: float = 0.3,
               half: str | None = None) -> np.ndarray:
    """Noisy ellipsoid-surface cloud, float32 [n,3] in mm.

    half='upper' keeps z > -0.2*a_z, half='lower' keeps z < +0.2*a_z (the two
    overlapping half-NeRFs of icp.py:47-50).
    """
    rng = np.random.default_rng(seed)
    axes = SEMI_AXES * diameter / 2.0
    out = np.empty((0, 3), dtype=np.float64)
    while len(out) < n:
        m = int((n - len(out)) * (2.2 if half else 1.0)) + 16
        u = rng.normal(size=(m, 3))
        u /= np.linalg.norm(u, axis=1, keepdims=True)
        p = u * axes + rng.normal(scale=sigma, size=(m, 3))
        if half == "upper":
            p = p[p[:, 2] > -0.2 * axes[2]]
        elif half == "lower":
            p = p[p[:, 2] < 0.2 * axes[2]]
        elif half is not None:
            raise ValueError("half must be None, 'upper' or 'lower'")
        out = np.concatenate([out, p], axis=0)
    return out[:n].astype(np.float32)


def rotvec_to_matrix(w) -> np.ndarray:
    """Rodrigues formula, float64."""
    w = np.asarray(w, dtype=np.float64)
    th = np.linalg.norm(w)
    if th < 1e-300:
        return np.eye(3)
    k = w / th
    K = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * (K @ K)


def random_rotation(rng) -> np.ndarray:
    """Haar-random rotation from a normalised quaternion."""
    q = rng.normal(size=4)
    q /= np.linalg.norm(q)
    w, x, y, z = q
    return np.array([
        [1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
        [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
        [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)],
    ])


def pose_matrix(R, t) -> np.ndarray:
    T = np.eye(4)
    T[:3, :3] = R
    T[:3, 3] = np.asarray(t, dtype=np.float64).reshape(3)
    return T


def true_pose(seed: int):
    """(R*, t*) with t* ~ (0, 0, 700) mm."""
    rng = np.random.default_rng(seed)
    R = random_rotation(rng)
    t = np.array([0.0, 0.0, 700.0]) + rng.normal(scale=10.0, size=3)
    return R, t


def make_candidates(b: int, seed: int, R_true=None, t_true=None, mix: str = "default"):
    """PnP+RANSAC-like candidate set around a true pose.

    90% are R* . Exp(w), |w| ~ U(0, 30 deg); 10% are Haar-random (RANSAC failures);
    t_k = t* + N(0, 5^2 mm); exactly one index k0 has |w| <= 0.1 deg.
    mix="aligned" / "haar" makes every candidate (but k0) of the first / second kind -- the
    pruned search's work depends on how well the two clouds of a candidate are aligned.
    Returns (R [b,3,3], t [b,3], k0).
    """
    frac_haar = {"default": 0.1, "aligned": 0.0, "haar": 1.0}[mix]
    rng = np.random.default_rng(seed)
    if R_true is None:
        R_true, t_true = true_pose(seed + 7919)
    Rs = np.empty((b, 3, 3))
    ts = t_true + rng.normal(scale=5.0, size=(b, 3))
    k0 = int(rng.integers(0, b))
    for k in range(b):
        axis = rng.normal(size=3)
        axis /= np.linalg.norm(axis)
        if k == k0:
            ang = np.deg2rad(rng.uniform(0.0, 0.1))
            Rs[k] = R_true @ rotvec_to_matrix(axis * ang)
        elif rng.uniform() < frac_haar:
            Rs[k] = random_rotation(rng)
        else:
            ang = np.deg2rad(rng.uniform(0.5, 30.0))
            Rs[k] = R_true @ rotvec_to_matrix(axis * ang)
    return Rs, ts, k0


def verification_matrices(Rs, R_true):
    """Candidate scoring matrices in the convention of verfication.py:83-85.

    The reference right-multiplies: pcgt = (pc1 . R1pred^T) . R_rel and pcpred = pc1 . R2pred,
    translations off.  Here candidate k plays R2pred and the true pose plays the
    ground-truth-rotated copy, so column-vector matrices (p' = M p) are
    Mq[k] = R_k^T (pcpred) and Mt[k] = R_true^T (pcgt).  Returns (Mq, Mt) [b,4,4] float64.
    """
    b = len(Rs)
    Mq = np.tile(np.eye(4), (b, 1, 1))
    Mt = np.tile(np.eye(4), (b, 1, 1))
    Mq[:, :3, :3] = np.transpose(Rs, (0, 2, 1))
    Mt[:, :3, :3] = np.asarray(R_true).T
    return Mq, Mt


def icp_pair(n_src: int, n_tgt: int, seed_src: int, seed_tgt: int, angle_deg: float = 1.5,
             shift_mm: float = 1.5, half: bool = False):
    """Source/target pair for ICP: target cloud, and a source cloud of the same object
    moved by a small rigid motion.  Returns (src f32, tgt f32, T_move 4x4 that was applied
    to the source, so ICP should recover ~inv(T_move))."""
    tgt = make_cloud(n_tgt, seed_tgt, half="lower" if half else None)
    src0 = make_cloud(n_src, seed_src, half="upper" if half else None)
    rng = np.random.default_rng(seed_src * 1000003 + seed_tgt)
    axis = rng.normal(size=3)
    axis /= np.linalg.norm(axis)
    R = rotvec_to_matrix(axis * np.deg2rad(angle_deg))
    d = rng.normal(size=3)
    d *= shift_mm / np.linalg.norm(d)
    T = pose_matrix(R, d)
    src = (src0.astype(np.float64) @ R.T + d).astype(np.float32)
    return src, tgt, T
