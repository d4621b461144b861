"""CPU oracle for the registration-and-verification hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module.  The product package
(``imagesequenceregistrationfor6dposeestimationlabeling_b200``) never does: it has no
CPU fallback and fails loudly without its CUDA library.

PARITY STATUS
-------------
* The sklearn half (``ADD``, ``ADDS``, ``compute_rel_poses``, ``calculate_relative_pose``)
  is PINNED: ``tests/golden/make_golden.py`` executes the reference's own function
  bodies (extracted with ``ast`` from /root/reference at generation time) and the
  committed vectors in ``tests/golden/*.npz`` are checked against this restatement.
* The Open3D half (``compute_point_cloud_distance``, ``evaluate_registration``,
  ``registration_icp``, ``PointCloud.transform``) is **parity unpinned**: Open3D is not
  installable in this environment (no wheel, no network, no Eigen/nanoflann headers), its
  version is unpinned by the reference (requirements.txt:14), and the reference holds
  no tests or golden vectors.  What follows restates Open3D's published algorithm
  (KDTreeFlann 1-NN, RegistrationICP loop, Eigen::umeyama without scaling) and anchors
  on the reference's call sites (verfication.py:97-102, icp.py:96-117).

All arithmetic is float64 (the reference's Open3D path is double: Vector3dVector).
Nearest neighbours use scipy ``cKDTree`` (exact 1-NN, multi-threaded) as the stand-in
for Open3D's nanoflann tree; ``ADDS`` uses sklearn ``KDTree(leaf_size=2)`` exactly as
choosePose.py:21-22 calls it.
"""
from __future__ import annotations

import dataclasses

import numpy as np
from scipy.spatial import cKDTree


# --------------------------------------------------------------------------------------
# pose algebra  (verfication.py:9-19, choosePose.py:18-22,43-51)
# --------------------------------------------------------------------------------------
def calculate_relative_pose(R1, T1, R2, T2):
    """Rel = [R2|T2] . inv([R1|T1]); returns (Rel[:3,:3], Rel[:3,3]).  verfication.py:9-19."""
    A = np.eye(4)
    A[:3, :3] = np.asarray(R1, dtype=np.float64)
    A[:3, 3] = np.asarray(T1, dtype=np.float64).reshape(3)
    B = np.eye(4)
    B[:3, :3] = np.asarray(R2, dtype=np.float64)
    B[:3, 3] = np.asarray(T2, dtype=np.float64).reshape(3)
    rel = B @ np.linalg.inv(A)
    return rel[:3, :3], rel[:3, -1]


def compute_rel_poses(R1, t1, R2, t2):
    """(R1^T R2, t2 - t1) -- deliberately NOT an SE(3) composition.  choosePose.py:43-51."""
    return np.dot(np.asarray(R1).T, R2), np.asarray(t2) - np.asarray(t1)


def rel_pose_table(RList, TList):
    """All-pairs n x n x 4 x 4 table.  choosePose.py:98-107."""
    n = len(TList)
    out = np.zeros((n, n, 4, 4))
    for i in range(n):
        for j in range(n):
            r, t = compute_rel_poses(RList[i], TList[i], RList[j], TList[j])
            m = np.eye(4)
            m[:3, :3] = r
            m[:3, 3:4] = np.asarray(t).reshape(3, 1)
            out[i][j] = m
    return out


def ADD(verts, gtR1, gtT1, R1, T1):
    """mean || (V gtR^T + gtT) - (V R^T + T) ||.  choosePose.py:18-19."""
    return np.linalg.norm(verts.dot(gtR1.T) + gtT1 - verts.dot(R1.T) - T1, axis=-1).mean()


def ADDS(verts, gtR1, gtT1, R1, T1, surfacePointsScaled):
    """One-directional mean 1-NN distance, sklearn KDTree(leaf_size=2).  choosePose.py:20-22.

    The reference reads ``surfacePointsScaled`` from a module global; it is explicit here.
    """
    from sklearn.neighbors import KDTree

    treeTr = KDTree(surfacePointsScaled.dot(R1.T) + T1, leaf_size=2)
    return treeTr.query(verts.dot(gtR1.T) + gtT1, k=1)[0].mean()


def choose_image(pred_rel_poses, gt_rel_poses, modelVerts, surfacePointsScaled, diameter):
    """ADD-S vote matrix, argmax row-sum, top-50.  choosePose.py:121-151."""
    n0, n1 = pred_rel_poses.shape[:2]
    error = np.zeros((n0, n1))
    for i in range(n0):
        for j in range(n1):
            gtR = gt_rel_poses[i][j][:3, :3]
            gtT = np.squeeze(gt_rel_poses[i][j][:3, 3:4])
            pR = pred_rel_poses[i][j][:3, :3]
            pT = np.squeeze(pred_rel_poses[i][j][:3, 3:4])
            e = ADDS(modelVerts, gtR, gtT, pR, pT, surfacePointsScaled)
            if e < 0.1 * diameter:
                error[i][j] = 1
    votes = np.sum(error, axis=1)
    image_id = int(np.argmax(votes))
    top = np.argsort(-votes)[:50]
    return error, image_id, top


# --------------------------------------------------------------------------------------
# nearest neighbour / Chamfer  (verfication.py:97-102, icp.py:113-117; Open3D upstream)
# --------------------------------------------------------------------------------------
def nearest(src, tgt, workers=-1):
    """Exact float64 1-NN of every src point in tgt -> (dist (N,), idx (N,) int64)."""
    src = np.ascontiguousarray(src, dtype=np.float64)
    tgt = np.ascontiguousarray(tgt, dtype=np.float64)
    tree = cKDTree(tgt)
    d, i = tree.query(src, k=1, workers=workers)
    return d, i.astype(np.int64)


def compute_point_cloud_distance(src, tgt, workers=-1):
    """Open3D PointCloud::ComputePointCloudDistance: Euclidean (sqrt) distance of each src
    point to its nearest tgt point; an empty target yields zeros (upstream)."""
    src = np.asarray(src, dtype=np.float64)
    if len(tgt) == 0:
        return np.zeros(len(src))
    if len(src) == 0:
        return np.zeros(0)
    return nearest(src, tgt, workers)[0]


def chamfer(a, b, workers=-1):
    """(mean(d a->b) + mean(d b->a)) / 2, unsquared.  verfication.py:97-101, icp.py:113-117."""
    da = compute_point_cloud_distance(a, b, workers)
    db = compute_point_cloud_distance(b, a, workers)
    return (np.mean(da) + np.mean(db)) / 2


def verify_chamfer(pc1, gt_R, gt_T, pred_R, pred_T, workers=-1):
    """The hot loop of verfication.py:61-108 on in-memory pose lists.

    gt_R/gt_T and pred_R/pred_T are per-image lists (len n); pair i uses images i, i+1.
    Returns (chamferdis list, min_index, min_chamfer) with first-minimum selection.
    """
    chamferdis = []
    for i in range(len(pred_R) - 1):
        R_rel, _ = calculate_relative_pose(gt_R[i], gt_T[i], gt_R[i + 1], gt_T[i + 1])
        R1pred = np.asarray(pred_R[i]).reshape(3, 3)
        R2pred = np.asarray(pred_R[i + 1]).reshape(3, 3)
        pc1p1 = pc1.dot(R1pred.T)  # translations disabled, verfication.py:83
        pcgt = pc1p1.dot(R_rel)  # :84
        pcpred = pc1.dot(R2pred)  # :85
        chamferdis.append(chamfer(pcpred, pcgt, workers))
    mn = min(chamferdis)
    return chamferdis, chamferdis.index(mn), mn


def verify_matrices(pc_q, pc_t, Mq, Mt, bidirectional=True, workers=-1):
    """Candidate scoring in its batched form: candidate k compares
    X = pc_q . Mq[k][:3,:3]^T + Mq[k][:3,3]  against  Y = pc_t . Mt[k][:3,:3]^T + Mt[k][:3,3].
    bidirectional -> Chamfer (verfication.py:97-101); else one-directional mean (ADD-S,
    choosePose.py:20-22).  Returns (losses float64 [B], first-min index)."""
    losses = []
    for k in range(len(Mq)):
        X = transform(pc_q, Mq[k])
        Y = transform(pc_t, Mt[k])
        if bidirectional:
            losses.append(chamfer(X, Y, workers))
        else:
            losses.append(np.mean(compute_point_cloud_distance(X, Y, workers)))
    losses = np.asarray(losses)
    return losses, int(np.argmin(losses))


# --------------------------------------------------------------------------------------
# ICP  (icp.py:88-117; Open3D RegistrationICP / Eigen::umeyama upstream)
# --------------------------------------------------------------------------------------
def radius_count(points, radius, target=None, chunk=512):
    """Neighbours with d^2 < radius^2 (strict, float64; the point itself included when the
    target is the cloud) -- what Open3D's remove_radius_outlier counts (upstream: KDTreeFlann
    SearchRadius -> nanoflann RadiusResultSet keeps `dist < radius`, both squared;
    generateCors.py:254-258, trainPose.py:343-347).  Brute force in chunks: the definition."""
    q = np.asarray(points, dtype=np.float64)
    t = q if target is None else np.asarray(target, dtype=np.float64)
    r2 = float(radius) * float(radius)
    out = np.zeros(len(q), dtype=np.int32)
    for lo in range(0, len(q), chunk):
        d = q[lo:lo + chunk, None, :] - t[None, :, :]
        out[lo:lo + chunk] = ((d * d).sum(-1) < r2).sum(1)
    return out


def remove_radius_outlier(points, nb_points, radius):
    """-> (kept points, kept indices): keep point i iff its radius count > nb_points
    (upstream PointCloud::RemoveRadiusOutliers)."""
    cnt = radius_count(points, radius)
    ind = np.nonzero(cnt > nb_points)[0]
    return np.asarray(points)[ind], ind


def pnp_inliers(p3d, p2d, cam, R, t, reperr=2.0):
    """Consensus set of ONE PnP hypothesis as cv2.solvePnPRansac scores it (choosePose.py:23-33
    calls it with reprojectionError=2, distCoeffs=None).  OpenCV (calib3d: PnPRansacCallback::
    computeError, RANSACPointSetRegistrator::findInliers), restated from the published source
    and PINNED against cv2 4.13 itself (tests/golden/reference_pnp_cv2.npz, made by
    tests/golden/make_golden_cv2.py; test_pnp_inliers_match_cv2_golden): points and image points are float32, the
    pinhole projection is evaluated in float64 and stored as float32 (z == 0 -> 1),
    err = float32(double(du)^2 + double(dv)^2), inlier iff err <= reperr^2 (double).
    Returns the boolean mask [n]."""
    P = np.asarray(p3d, dtype=np.float32).astype(np.float64)
    uv = np.asarray(p2d, dtype=np.float32)
    K = np.asarray(cam, dtype=np.float64).reshape(3, 3)
    R = np.asarray(R, dtype=np.float64).reshape(3, 3)
    t = np.asarray(t, dtype=np.float64).reshape(3)
    x = ((R[0, 0] * P[:, 0] + R[0, 1] * P[:, 1]) + R[0, 2] * P[:, 2]) + t[0]
    y = ((R[1, 0] * P[:, 0] + R[1, 1] * P[:, 1]) + R[1, 2] * P[:, 2]) + t[1]
    z = ((R[2, 0] * P[:, 0] + R[2, 1] * P[:, 1]) + R[2, 2] * P[:, 2]) + t[2]
    with np.errstate(divide="ignore"):
        iz = np.where(z != 0.0, 1.0 / z, 1.0)
    xn, yn = x * iz, y * iz
    pu = (K[0, 0] * xn + K[0, 1] * yn + K[0, 2]).astype(np.float32)
    pv = (K[1, 1] * yn + K[1, 2]).astype(np.float32)
    du, dv = (uv[:, 0] - pu).astype(np.float64), (uv[:, 1] - pv).astype(np.float64)
    err = (du * du + dv * dv).astype(np.float32)
    return err.astype(np.float64) <= float(reperr) * float(reperr)


def transform(points, T):
    """PointCloud.transform: p <- T[:3,:3] p + T[:3,3] (float64).  icp.py:22,110."""
    T = np.asarray(T, dtype=np.float64)
    return np.asarray(points, dtype=np.float64) @ T[:3, :3].T + T[:3, 3]


@dataclasses.dataclass
class RegistrationResult:
    transformation: np.ndarray
    fitness: float = 0.0
    inlier_rmse: float = 0.0
    correspondence_set: np.ndarray = dataclasses.field(
        default_factory=lambda: np.zeros((0, 2), dtype=np.int32))
    iterations: int = 0  # number of updates applied (not part of Open3D's struct)


def _result_and_correspondences(src_t, tree, n_src, max_dist, T, workers):
    """GetRegistrationResultAndCorrespondences (upstream): 1-NN per source point, keep
    d^2 < max_dist^2 (strict; SearchHybrid uses lower_bound on radius^2)."""
    res = RegistrationResult(np.array(T, dtype=np.float64))
    if max_dist <= 0.0 or n_src == 0:
        return res
    d, idx = tree.query(src_t, k=1, workers=workers)
    d2 = d * d
    keep = d2 < max_dist * max_dist
    k = int(keep.sum())
    if k == 0:
        return res
    res.correspondence_set = np.stack(
        [np.nonzero(keep)[0], idx[keep]], axis=1).astype(np.int32)
    res.fitness = k / float(n_src)
    res.inlier_rmse = float(np.sqrt(d2[keep].sum() / k))
    return res


def umeyama(src, dst):
    """Eigen::umeyama(src, dst, with_scaling=false): 4x4 least-squares rigid motion
    dst ~ R src + t.  Sigma = (1/n) (dst - mu_d)(src - mu_s)^T, SVD, det-sign fix."""
    src = np.asarray(src, dtype=np.float64)
    dst = np.asarray(dst, dtype=np.float64)
    n = len(src)
    mu_s = src.mean(axis=0)
    mu_d = dst.mean(axis=0)
    sigma = (dst - mu_d).T @ (src - mu_s) / n
    U, _, Vt = np.linalg.svd(sigma)
    S = np.ones(3)
    if np.linalg.det(U) * np.linalg.det(Vt) < 0:
        S[2] = -1.0
    R = U @ np.diag(S) @ Vt
    T = np.eye(4)
    T[:3, :3] = R
    T[:3, 3] = mu_d - R @ mu_s
    return T


def evaluate_registration(source, target, max_correspondence_distance, transformation=None,
                          workers=-1):
    """o3d.pipelines.registration.evaluate_registration.  icp.py:97-98."""
    T = np.eye(4) if transformation is None else np.asarray(transformation, dtype=np.float64)
    src_t = transform(source, T)
    tree = cKDTree(np.ascontiguousarray(target, dtype=np.float64))
    return _result_and_correspondences(src_t, tree, len(source), max_correspondence_distance,
                                       T, workers)


def registration_icp(source, target, max_correspondence_distance, init=None,
                     max_iteration=30, relative_fitness=1e-6, relative_rmse=1e-6, workers=-1):
    """o3d.pipelines.registration.registration_icp with TransformationEstimationPointToPoint
    and default ICPConvergenceCriteria (30 / 1e-6 / 1e-6).  icp.py:101-103.

    Loop (upstream RegistrationICP): T=init; pcd=source.Transform(init);
    res=evaluate; repeat { U=umeyama(corr); T=U T; pcd.Transform(U); prev=res;
    res=evaluate; break if |dfitness|<rf and |drmse|<rr }.  The returned result is the
    re-evaluation AFTER the last update; an empty correspondence set gives U=I.
    """
    T = np.eye(4) if init is None else np.array(init, dtype=np.float64)
    tgt = np.ascontiguousarray(target, dtype=np.float64)
    tree = cKDTree(tgt)
    pcd = transform(source, T)
    n_src = len(source)
    res = _result_and_correspondences(pcd, tree, n_src, max_correspondence_distance, T, workers)
    iters = 0
    for _ in range(max_iteration):
        cs = res.correspondence_set
        if len(cs) == 0:
            U = np.eye(4)
        else:
            U = umeyama(pcd[cs[:, 0]], tgt[cs[:, 1]])
        T = U @ T
        pcd = transform(pcd, U)  # incremental, as upstream
        iters += 1
        prev = res
        res = _result_and_correspondences(pcd, tree, n_src, max_correspondence_distance, T,
                                          workers)
        if (abs(prev.fitness - res.fitness) < relative_fitness
                and abs(prev.inlier_rmse - res.inlier_rmse) < relative_rmse):
            break
    res.iterations = iters
    return res


def icp_script(upper, lower, R_GT, t_GT, R_pred, t_pred, cad, threshold=20.0, workers=-1,
               **criteria):
    """The flow of icp.py:64-117 on in-memory arrays.  Returns (evaluation, reg_p2p,
    final chamfer of (transformed source + target) vs cad)."""
    upper = np.asarray(upper, dtype=np.float32)
    lower = np.asarray(lower, dtype=np.float32)
    actual_upper = upper.dot(np.asarray(R_GT).T) + np.asarray(t_GT)  # icp.py:68
    M = np.eye(4)
    M[:3, :3] = R_pred
    M[:3, 3] = t_pred
    init = np.linalg.inv(M)  # icp.py:88-92
    ev = evaluate_registration(actual_upper, lower, threshold, init, workers)
    reg = registration_icp(actual_upper, lower, threshold, init, workers=workers, **criteria)
    merged = np.concatenate([transform(actual_upper, reg.transformation),
                             np.asarray(lower, dtype=np.float64)], axis=0)  # :110-111
    return ev, reg, chamfer(merged, cad, workers)


def p3p(P, uv, cam):
    """All solutions of the perspective three-point problem for object points P [3,3] seen at
    pixels uv [3,2] (Grunert's quartic as in Haralick et al. 1994): the minimal solver behind
    ``cv2.solvePnPRansac(..., flags=cv2.SOLVEPNP_P3P)`` (choosePose.py:23).  Pinned against
    cv2.solveP3P on the committed 3-point sets of tests/golden/reference_pnp_cv2.npz.
    Returns a list of (R, t) with camera = R object + t."""
    P = np.asarray(P, dtype=np.float64)
    K = np.asarray(cam, dtype=np.float64).reshape(3, 3)
    f = (np.linalg.inv(K) @ np.concatenate([np.asarray(uv, dtype=np.float64), np.ones((3, 1))], 1).T).T
    f /= np.linalg.norm(f, axis=1, keepdims=True)
    a2 = np.sum((P[1] - P[2]) ** 2); b2 = np.sum((P[0] - P[2]) ** 2); c2 = np.sum((P[0] - P[1]) ** 2)
    ca, cb, cg = f[1] @ f[2], f[0] @ f[2], f[0] @ f[1]
    q1, q2, q3, q4 = (a2 - c2) / b2, (a2 + c2) / b2, (b2 - c2) / b2, (b2 - a2) / b2
    coef = [(q1 - 1) ** 2 - 4 * c2 / b2 * ca * ca,
            4 * (q1 * (1 - q1) * cb - (1 - q2) * ca * cg + 2 * c2 / b2 * ca * ca * cb),
            2 * (q1 * q1 - 1 + 2 * q1 * q1 * cb * cb + 2 * q3 * ca * ca - 4 * q2 * ca * cb * cg + 2 * q4 * cg * cg),
            4 * (-q1 * (1 + q1) * cb + 2 * a2 / b2 * cg * cg * cb - (1 - q2) * ca * cg),
            (1 + q1) ** 2 - 4 * a2 / b2 * cg * cg]

    def frame(X):
        e1 = X[1] - X[0]
        e1 = e1 / np.linalg.norm(e1)
        e3 = np.cross(e1, X[2] - X[0])
        e3 = e3 / np.linalg.norm(e3)
        return np.stack([e1, np.cross(e3, e1), e3], 1)

    out = []
    for v in np.roots(coef):
        if abs(v.imag) > 1e-6 * max(1.0, abs(v.real)) or v.real <= 0:
            continue
        v = v.real
        den = 2 * (cg - v * ca)
        if abs(den) < 1e-12:
            continue
        u = ((q1 - 1) * v * v - 2 * q1 * cb * v + 1 + q1) / den
        s1sq = b2 / (1 + v * v - 2 * v * cb)
        if u <= 0 or s1sq <= 0:
            continue
        s1 = np.sqrt(s1sq)
        Q = np.stack([s1 * f[0], u * s1 * f[1], v * s1 * f[2]])
        R = frame(Q) @ frame(P).T
        out.append((R, Q[0] - R @ P[0]))
    return out


def estimate_normals(points, neighborhood_size=50, disambiguate_directions=True):
    """pytorch3d.ops.estimate_pointcloud_normals(points[None], neighborhood_size,
    disambiguate_directions)[0] restated (generateCors.py:200-215 calls it with 400 neighbours on
    1000 farthest-point samples and negates the result).  pytorch3d is not installable here ->
    parity unpinned for this function; it follows upstream's published algorithm
    (points_normals.py): k nearest neighbours of the same cloud (self included), covariance about
    the neighbourhood mean (mean of outer products), eigenvector of the smallest eigenvalue of
    eigh, flipped when fewer than k / 2 neighbours have a positive projection on it.
    Returns float64 [n, 3]."""
    P = np.asarray(points, dtype=np.float64)
    k = int(neighborhood_size)
    _, idx = cKDTree(P).query(P, k=k)
    idx = idx.reshape(len(P), k)
    nb = P[idx]                                   # [n, k, 3]
    cen = nb - nb.mean(axis=1, keepdims=True)
    cov = np.einsum("nki,nkj->nij", cen, cen) / k
    w, v = np.linalg.eigh(cov)
    nrm = v[:, :, 0].copy()
    if disambiguate_directions:
        proj = np.einsum("nki,ni->nk", nb - P[:, None, :], nrm)
        flip = (proj > 0).sum(axis=1) < 0.5 * k
        nrm[flip] *= -1.0
    return nrm
