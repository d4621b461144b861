"""The reference scripts' helper functions, same names / argument order / returns.

  calculate_relative_pose   verfication.py:9-19
  compute_rel_poses         choosePose.py:43-51
  relative_pose_table       choosePose.py:98-107   (the n x n x 4 x 4 table)
  ADD, ADDS                 choosePose.py:18-22, inference.py:116-120
  choose_image              choosePose.py:121-151  (ADD-S vote, argmax, top-50; device-resident)
  choose_image_from_poses   choosePose.py:98-107 + 121-151 in one pass (tables built on the device)
  pnp                       choosePose.py:23-33    (cv2.solvePnPRansac with P3P, on the device)
  draw_registration_result, vp   verfication.py:21-31, icp.py:8-27  (GUI; no-ops here)

Pose algebra is host-side float64 numpy (3x3 / 4x4, negligible work); everything that
touches a point cloud goes to the CUDA library through ``api``.
"""
from __future__ import annotations

import numpy as np

from . import api

#: the reference's ADDS reads this module global (choosePose.py:21,160); assign it, or
#: pass ``surface_points=`` explicitly.
surfacePointsScaled = None


def calculate_relative_pose(R1, T1, R2, T2):
    """Rel = [R2|T2] . inv([R1|T1]) -> (Rel[:3,:3], Rel[:3,3])."""
    T1_column = np.asarray(T1, dtype=np.float64).reshape(-1, 1)
    RT1 = np.vstack([np.hstack((np.asarray(R1, dtype=np.float64), T1_column)), [0, 0, 0, 1]])
    T2_column = np.asarray(T2, dtype=np.float64).reshape(-1, 1)
    RT2 = np.vstack([np.hstack((np.asarray(R2, dtype=np.float64), T2_column)), [0, 0, 0, 1]])
    Rel = np.dot(RT2, np.linalg.inv(RT1))
    return Rel[:3, :3], Rel[:3, -1]


def compute_rel_poses(R1, t1, R2, t2):
    """(R1^T R2, t2 - t1).  Not an SE(3) composition -- kept exactly as the reference."""
    return np.dot(np.asarray(R1).T, R2), np.asarray(t2) - np.asarray(t1)


def relative_pose_table(RList, TList) -> np.ndarray:
    """relative_poses[i][j] = 4x4 of compute_rel_poses(R_i, t_i, R_j, t_j); vectorised."""
    R = np.asarray(RList, dtype=np.float64).reshape(-1, 3, 3)
    t = np.asarray(TList, dtype=np.float64).reshape(-1, 3)
    n = len(t)
    out = np.zeros((n, n, 4, 4))
    out[:, :, :3, :3] = np.einsum("iba,jbc->ijac", R, R)  # R_i^T R_j
    out[:, :, :3, 3] = t[None, :, :] - t[:, None, :]
    out[:, :, 3, 3] = 1.0
    return out


def ADD(verts, gtR1, gtT1, R1, T1):
    """mean ||(V gtR^T + gtT) - (V R^T + T)|| on the device (K1 twice + FP64 mean)."""
    import torch

    V = api._points(verts, api._device())
    P = np.stack([api.pose_from_Rt(gtR1, gtT1), api.pose_from_Rt(R1, T1)])
    out = api.transform_points(V, P).to(torch.float64)
    return float(torch.linalg.norm(out[0] - out[1], dim=-1).mean().item())


def ADDS(verts, gtR1, gtT1, R1, T1, surface_points=None):
    """mean 1-NN distance from verts.gtR^T+gtT into surface.R^T+T (one direction)."""
    S = surfacePointsScaled if surface_points is None else surface_points
    if S is None:
        raise NameError("surfacePointsScaled is not set (choosePose.py:160 defines it as a global)")
    return float(api.adds(verts, gtR1, gtT1, R1, T1, S).item())


def _vote_result(losses_dev, n0, n1, diameter):
    """Vote + selection of choosePose.py:135-151.  error / votes / first-argmax come from the
    device (isr_vote); the top-50 list is numpy's own ``argsort(-votes)[:50]`` on the n vote
    counts, so that ties are ordered exactly as the reference's (unstable) sort orders them."""
    error, votes, best = api.vote(losses_dev.view(n0, n1), 0.1 * float(diameter))
    votes_h = votes.cpu().numpy().astype(np.float64)   # (np.sum of a float64 0/1 matrix)
    image_id = int(best[0].item())
    assert image_id == int(np.argmax(votes_h))
    return error.cpu().numpy().astype(np.float64), image_id, np.argsort(-votes_h)[:50]


class _PairScorer:
    """ADD-S of pose pairs against ONE surface cloud, for the vote: the surface is prepared once;
    every chunk of pairs first gets rigorous bounds from the surface's tile spheres
    (isr_adds_bounds: lower <= ADDS <= upper); pairs that the bounds already place on one side of
    the threshold keep that bound as their "loss", the others are scored exactly at the end
    (isr_adds_fixed_target) -- one host synchronisation in total, to count them."""

    def __init__(self, modelVerts, surface, n_pairs, threshold, use_bounds=True):
        import torch

        self.dev = api._device()
        self.V = api._points(modelVerts, self.dev)
        if use_bounds and self.V.shape[0] > 0:
            # curve order: 32 consecutive vertices are one small patch (isr_adds_bounds tests the
            # target's stage spheres once per warp); a mean over the vertices does not depend on it
            self.V = self.V[api.spatial_order(self.V).to(torch.int64)].contiguous()
        self.S = api._points(surface, self.dev)
        self.thr = float(threshold)
        self.losses = torch.empty((n_pairs,), dtype=torch.float64, device=self.dev)
        self.M = torch.empty((n_pairs, 4, 4), dtype=torch.float64, device=self.dev) if use_bounds else None
        self.target = None
        if use_bounds:
            cen = api.centroid_of(self.S)
            self.target = api.prepare_cloud(self.S, centroid=cen, perm=api.spatial_order(self.S),
                                            stage_centroids=True)
        self.undecided = 0

    def add(self, k0, poses_gt, poses_pred):
        """pairs k0 .. k0 + len: ADDS(verts, gt, pred) (choosePose.py:131-134)."""
        import torch

        c = poses_gt.shape[0]
        M = api.rigid_relative(poses_gt, poses_pred)
        bounds = api.adds_bounds(self.V, M, self.target, presorted=True) if self.target is not None else None
        if bounds is None:
            self.losses[k0:k0 + c] = api.adds_fixed(self.V, M, self.S).losses
            return
        lo, hi = bounds
        self.M[k0:k0 + c] = M
        # decided pairs: upper < thr (votes) keeps upper, lower >= thr (no vote) keeps lower; NaN marks the rest
        nan = torch.full_like(lo, float("nan"))
        self.losses[k0:k0 + c] = torch.where(hi < self.thr, hi, torch.where(lo >= self.thr, lo, nan))

    def finish(self):
        import torch

        if self.target is not None:
            idx = torch.nonzero(torch.isnan(self.losses), as_tuple=False)[:, 0]   # the one host sync
            self.undecided = int(idx.numel())
            for i0 in range(0, self.undecided, 65536):
                sel = idx[i0:i0 + 65536]
                self.losses[sel] = api.adds_fixed(self.V, self.M[sel], self.S).losses
        return self.losses


def choose_image(pred_rel_poses, gt_rel_poses, modelVerts, diameter, surface_points=None,
                 chunk: int = 65536, use_bounds: bool = True, stats=None):
    """ADD-S vote over all pose pairs -> (error n x n, image_id, top-50 indices), from the two
    n x n x 4 x 4 tables that ``--rel_poses`` saves (choosePose.py:121-151).  The tables are
    uploaded once; scoring (_PairScorer), the 0.1 x diameter test, the row sums and the argmax
    run on the device.  `stats` (a dict) receives the number of pairs that needed the exact search."""
    import torch

    S = surfacePointsScaled if surface_points is None else surface_points
    if S is None:
        raise NameError("surfacePointsScaled is not set")
    dev = api._device()
    pred = pred_rel_poses if isinstance(pred_rel_poses, torch.Tensor) else np.asarray(pred_rel_poses, dtype=np.float64)
    gt = gt_rel_poses if isinstance(gt_rel_poses, torch.Tensor) else np.asarray(gt_rel_poses, dtype=np.float64)
    n0, n1 = pred.shape[:2]
    P = api._poses(pred.reshape(-1, 4, 4), dev)
    G = api._poses(gt.reshape(-1, 4, 4), dev)
    sc = _PairScorer(modelVerts, S, n0 * n1, 0.1 * float(diameter), use_bounds)
    for k0 in range(0, n0 * n1, chunk):
        sc.add(k0, G[k0:k0 + chunk], P[k0:k0 + chunk])
    losses = sc.finish()
    if stats is not None:
        stats.update(pairs=n0 * n1, exact=sc.undecided if use_bounds and sc.target is not None else n0 * n1)
    return _vote_result(losses, n0, n1, diameter)


def choose_image_from_poses(pred_R, pred_t, gt_R, gt_t, modelVerts, diameter, surface_points=None,
                            rows_per_chunk: int = 64, use_bounds: bool = True, stats=None):
    """``--rel_poses`` and ``--choose_image`` in one pass (choosePose.py:98-107, 121-151): the two
    relative-pose tables are never materialised -- each chunk of table rows is built on the
    device (isr_rel_pose_table) and scored at once.  Same return as `choose_image`."""
    import torch

    S = surfacePointsScaled if surface_points is None else surface_points
    if S is None:
        raise NameError("surfacePointsScaled is not set")
    dev = api._device()
    Rp = api._to_dev(np.asarray(pred_R, dtype=np.float64).reshape(-1, 9), torch.float64, dev)
    tp = api._to_dev(np.asarray(pred_t, dtype=np.float64).reshape(-1, 3), torch.float64, dev)
    Rg = api._to_dev(np.asarray(gt_R, dtype=np.float64).reshape(-1, 9), torch.float64, dev)
    tg = api._to_dev(np.asarray(gt_t, dtype=np.float64).reshape(-1, 3), torch.float64, dev)
    n = Rp.shape[0]
    if not (tp.shape[0] == Rg.shape[0] == tg.shape[0] == n):
        raise ValueError("predicted and ground-truth pose lists must have the same length")
    sc = _PairScorer(modelVerts, S, n * n, 0.1 * float(diameter), use_bounds)
    step = max(1, int(rows_per_chunk)) * n
    for k0 in range(0, n * n, step):
        c = min(step, n * n - k0)
        sc.add(k0, api.relative_pose_table(Rg, tg, k0, c), api.relative_pose_table(Rp, tp, k0, c))
    losses = sc.finish()
    if stats is not None:
        stats.update(pairs=n * n, exact=sc.undecided if use_bounds and sc.target is not None else n * n)
    return _vote_result(losses, n, n, diameter)


def select_pnp_hypothesis(h3d, h2d, cam, Rs, ts, reperr: float = 2.0):
    """Score PnP hypotheses (Rs [B,3,3], ts [B,3]) on the device and return the reference's
    `pnp` triple for the winner (choosePose.py:23-33): (R, t, inlier indices) of the first
    hypothesis with the largest consensus, or the failure sentinel (1, 1, 1) -- which
    `verify_poses(valid_mask=...)` maps to +inf -- when no hypothesis has an inlier.  The
    hypothesis GENERATION (cv2's P3P + RNG) stays with the caller; this replaces the consensus
    loop, so that candidates go to the verification without a host round trip per pose."""
    Rs = np.asarray(Rs, dtype=np.float64).reshape(-1, 3, 3)
    ts = np.asarray(ts, dtype=np.float64).reshape(-1, 3)
    P = np.tile(np.eye(4), (len(Rs), 1, 1))
    P[:, :3, :3], P[:, :3, 3] = Rs, ts
    counts, flags = api.score_pnp_hypotheses(h3d, h2d, cam, P, reperr, return_inliers=True)
    counts = counts.cpu().numpy()
    if len(counts) == 0 or counts.max() == 0:
        return 1, 1, 1
    k = int(np.argmax(counts))  # first maximum
    return Rs[k].copy(), ts[k].copy(), np.nonzero(flags[k].cpu().numpy())[0]


def pnp(h3d, h2d, cam, itr=100, reperr=2, flag=None, gtR=None, gtT=None, spts=None, ret=None, seed=0):
    """choosePose.py:23-33 with the same signature and returns: ``(rotmat 3x3, tvec (3,), inlier
    indices)`` or the failure sentinel ``(1, 1, 1)``.  The RANSAC loop (P3P hypotheses, consensus,
    refit) runs on the device (api.pnp_ransac); `flag` is accepted for compatibility -- the
    reference only ever passes cv2.SOLVEPNP_P3P -- and gtR / gtT / spts / ret are unused, as in
    the reference."""
    r = api.pnp_ransac(h3d, h2d, cam, iterations=int(itr), reprojection_error=float(reperr), seed=seed)
    if not r.ok:
        print("pose could not be estimated with these correspondences")
        return 1, 1, 1
    return r.R, r.t, r.inliers


def draw_registration_result(source, target, transformation):
    """Open3D GUI in the reference (blocking); a no-op here."""
    return None


def vp(finalV):
    """Point-cloud viewer in the reference; a no-op here."""
    return None
