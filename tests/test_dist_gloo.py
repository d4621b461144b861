"""Multi-rank host logic on CPU: world_size 2 over gloo (127.0.0.1).  The compute back end
is injected (oracle-backed), so this covers exactly the rank arithmetic and the collectives
that the NCCL path uses: block sharding, exact first-minimum argmin, loss gather, and the
17-double all-reduce inside the ICP loop."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as td
import torch.multiprocessing as mp

from imagesequenceregistrationfor6dposeestimationlabeling_b200 import dist, synth
from oracle import oracle


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _oracle_scorer(cloud_q, poses_q, poses_t, cloud_t, mode, valid_mask):
    ct = cloud_q if cloud_t is None else cloud_t
    losses, _ = oracle.verify_matrices(cloud_q, ct, poses_q, poses_t, bidirectional=(mode == "chamfer"),
                                       workers=1)
    if valid_mask is not None:
        losses = np.where(np.asarray(valid_mask, dtype=bool), losses, np.inf)
    best = int(np.argmin(losses))
    return (torch.from_numpy(losses), torch.tensor([best], dtype=torch.int64),
            torch.tensor([losses[best]], dtype=torch.float64))


class _OracleIcpBackend:
    """numpy restatement of accumulate / solve for one source shard (tests only)."""

    def __init__(self, source_shard, target, inits):
        from scipy.spatial import cKDTree
        self.src = np.asarray(source_shard, dtype=np.float64)
        self.tgt = np.asarray(target, dtype=np.float64)
        self.tree = cKDTree(self.tgt)
        self.T = [np.array(t) for t in inits]
        self.state = [dict(fitness=0.0, rmse=0.0, evals=0, iters=0, done=False) for _ in inits]

    def accumulate(self, max_dist):
        sums = np.zeros((len(self.T), 17))
        for k, T in enumerate(self.T):
            if self.state[k]["done"] or len(self.src) == 0:
                continue
            s = oracle.transform(self.src, T)
            d, j = self.tree.query(s, k=1)
            keep = d * d < max_dist * max_dist
            s, t, d2 = s[keep], self.tgt[j[keep]], (d * d)[keep]
            sums[k, 0:3] = s.sum(0)
            sums[k, 3:6] = t.sum(0)
            sums[k, 6:15] = (t.T @ s).reshape(9)
            sums[k, 15] = d2.sum()
            sums[k, 16] = keep.sum()
        self.sums = torch.from_numpy(sums)
        return self.sums

    def solve(self, sums, ns_total, rf, rr, final_eval):
        S = sums.numpy()
        for k, st in enumerate(self.state):
            if st["done"]:
                continue
            cnt = S[k, 16]
            fit = cnt / ns_total
            rmse = np.sqrt(S[k, 15] / cnt) if cnt > 0 else 0.0
            had, pf, pr = st["evals"] > 0, st["fitness"], st["rmse"]
            st.update(fitness=fit, rmse=rmse, evals=st["evals"] + 1)
            if (had and abs(pf - fit) < rf and abs(pr - rmse) < rr) or final_eval:
                st["done"] = True
                continue
            st["iters"] += 1
            if cnt == 0:
                continue
            ms, mt = S[k, 0:3] / cnt, S[k, 3:6] / cnt
            sig = S[k, 6:15].reshape(3, 3) / cnt - np.outer(mt, ms)
            U, _, Vt = np.linalg.svd(sig)
            D = np.diag([1, 1, -1.0 if np.linalg.det(U) * np.linalg.det(Vt) < 0 else 1.0])
            R = U @ D @ Vt
            Uu = np.eye(4)
            Uu[:3, :3], Uu[:3, 3] = R, mt - R @ ms
            self.T[k] = Uu @ self.T[k]

    def results(self):
        return [dict(T=self.T[k], **self.state[k]) for k in range(len(self.T))]


class _OracleIcpTargetBackend(_OracleIcpBackend):
    """Target-sharded counterpart: all source points against one target slice (tests only)."""

    def search(self):
        idx = np.zeros((len(self.T), len(self.src)), dtype=np.int32)
        D = np.full((len(self.T), len(self.src)), np.inf)
        self._s = [None] * len(self.T)
        for k, T in enumerate(self.T):
            if self.state[k]["done"]:
                continue
            s = oracle.transform(self.src, T)
            _, j = self.tree.query(s, k=1)
            idx[k] = j
            D[k] = ((s - self.tgt[j]) ** 2).sum(1)
            self._s[k] = s
        return torch.from_numpy(idx), torch.from_numpy(D)

    def accumulate(self, local_idx, max_dist):
        sums = np.zeros((len(self.T), 17))
        li = local_idx.numpy()
        for k in range(len(self.T)):
            if self.state[k]["done"]:
                continue
            own = li[k] >= 0
            s, t = self._s[k][own], self.tgt[li[k][own]]
            d2 = ((s - t) ** 2).sum(1)
            keep = d2 < max_dist * max_dist
            s, t, d2 = s[keep], t[keep], d2[keep]
            sums[k, 0:3], sums[k, 3:6] = s.sum(0), t.sum(0)
            sums[k, 6:15] = (t.T @ s).reshape(9)
            sums[k, 15], sums[k, 16] = d2.sum(), keep.sum()
        return torch.from_numpy(sums)


def _oracle_multistart(source, target, inits, max_dist, max_iteration, rf, rr):
    Ts, fit, rmse, iters, ch = [], [], [], [], []
    for T0 in inits:
        o = oracle.registration_icp(source, target, max_dist, T0, max_iteration=max_iteration,
                                    relative_fitness=rf, relative_rmse=rr)
        Ts.append(o.transformation); fit.append(o.fitness); rmse.append(o.inlier_rmse)
        iters.append(o.iterations)
        ch.append(oracle.chamfer(oracle.transform(source, o.transformation), target))
    return Ts, fit, rmse, iters, np.array(ch)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    td.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    out = {}
    try:
        # exact first-min argmin, including a cross-rank tie (lowest global index wins)
        loss = torch.tensor([2.5 if rank == 0 else 2.5], dtype=torch.float64)
        idx = torch.tensor([7 if rank == 0 else 3], dtype=torch.int64)
        m, i = dist.global_first_argmin(loss, idx)
        out["tie"] = (float(m), int(i))
        loss = torch.tensor([float("inf") if rank == 0 else 1.25], dtype=torch.float64)
        idx = torch.tensor([0 if rank == 0 else 9], dtype=torch.int64)
        m, i = dist.global_first_argmin(loss, idx)
        out["inf"] = (float(m), int(i))

        cloud = synth.make_cloud(1500, seed=1)
        R_true, _ = synth.true_pose(3)
        Rs, _, k0 = synth.make_candidates(9, seed=10, R_true=R_true, t_true=np.zeros(3))
        Rs = np.concatenate([Rs, Rs[k0:k0 + 1]])        # duplicate best candidate, index 9
        Mq, Mt = synth.verification_matrices(Rs, R_true)
        bi, bl, losses = dist.verify_poses_sharded(cloud, Mq, Mt, gather_losses=True,
                                                   scorer=_oracle_scorer)
        out["verify"] = (int(bi), float(bl), losses.numpy().copy(), k0)
        valid = np.ones(len(Mq), dtype=bool)
        valid[k0] = False
        bi2, _, _ = dist.verify_poses_sharded(cloud, Mq, Mt, valid_mask=valid, scorer=_oracle_scorer)
        out["verify_masked"] = int(bi2)
        bi3, bl3, l3 = dist.verify_poses_sharded(cloud, Mq[:1], Mt[:1], gather_losses=True,
                                                 scorer=_oracle_scorer)  # fewer candidates than ranks
        out["verify_one"] = (int(bi3), len(l3))

        src, tgt, _ = synth.icp_pair(3001, 3500, 6, 7)
        res = dist.icp_sharded(src, tgt, np.eye(4), 20.0, max_iteration=12,
                               backend_factory=_OracleIcpBackend)
        out["icp"] = res[0]
        # the north star's variant: target rows sharded, per-point MIN all-reduce; the target
        # holds a duplicated block so that equal-distance neighbours sit on different ranks
        tgt2 = np.concatenate([tgt, tgt[:700]])
        res = dist.icp_sharded(src, tgt2, np.eye(4), 20.0, max_iteration=12, shard="target",
                               backend_factory=_OracleIcpTargetBackend)
        out["icp_target"] = res[0]
        # config 5: starts split across ranks (3 starts over 2 ranks: 2 + 1), one duplicated
        # start so that the best Chamfer value exists on both ranks (lowest index must win)
        s5, t5, _ = synth.icp_pair(1201, 1500, 6, 7)
        inits = np.stack([synth.pose_matrix(synth.rotvec_to_matrix([0, 0, a]), [0, 0, 0])
                          for a in (0.0, 2.0, 0.0)])
        out["multistart"] = dist.multistart_icp_sharded(s5, t5, inits, 20.0, max_iteration=6,
                                                        runner=_oracle_multistart)
        # no GPU here: creating the peer-memory exchange fails on every rank; the ranks must
        # walk through its collectives together and agree on "unavailable" (no hang, no raise)
        out["peer_unavailable"] = dist.peer_exchange(required=False) is None
    finally:
        td.destroy_process_group()
    q.put((rank, out))


def test_world_size_2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=300) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for r in (0, 1):
        assert got[r]["peer_unavailable"] is True
        assert got[r]["tie"] == (2.5, 3)
        assert got[r]["inf"] == (1.25, 9)
    cloud = synth.make_cloud(1500, seed=1)
    R_true, _ = synth.true_pose(3)
    Rs, _, k0 = synth.make_candidates(9, seed=10, R_true=R_true, t_true=np.zeros(3))
    Rs = np.concatenate([Rs, Rs[k0:k0 + 1]])
    Mq, Mt = synth.verification_matrices(Rs, R_true)
    ref, ref_best = oracle.verify_matrices(cloud, cloud, Mq, Mt)
    for r in (0, 1):
        bi, bl, losses, kk = got[r]["verify"]
        assert bi == ref_best == k0 == kk
        assert bl == ref[ref_best]
        np.testing.assert_array_equal(losses, ref)
        assert got[r]["verify_masked"] == 9          # the duplicate takes over
        assert got[r]["verify_one"] == (0, 1)
    src, tgt, _ = synth.icp_pair(3001, 3500, 6, 7)
    o = oracle.registration_icp(src, tgt, 20.0, np.eye(4), max_iteration=12)
    for r in (0, 1):
        res = got[r]["icp"]
        np.testing.assert_allclose(res["T"], o.transformation, rtol=1e-9, atol=1e-9)
        assert res["iters"] == o.iterations
        np.testing.assert_allclose(res["fitness"], o.fitness, rtol=1e-12)
        np.testing.assert_allclose(res["rmse"], o.inlier_rmse, rtol=1e-9)
    np.testing.assert_array_equal(got[0]["icp"]["T"], got[1]["icp"]["T"])   # bit-identical ranks
    o2 = oracle.registration_icp(src, np.concatenate([tgt, tgt[:700]]), 20.0, np.eye(4), max_iteration=12)
    for r in (0, 1):
        res = got[r]["icp_target"]
        np.testing.assert_allclose(res["T"], o2.transformation, rtol=1e-9, atol=1e-9)
        assert res["iters"] == o2.iterations
        np.testing.assert_allclose(res["fitness"], o2.fitness, rtol=1e-12)
        np.testing.assert_allclose(res["rmse"], o2.inlier_rmse, rtol=1e-9)
    np.testing.assert_array_equal(got[0]["icp_target"]["T"], got[1]["icp_target"]["T"])
    s5, t5, _ = synth.icp_pair(1201, 1500, 6, 7)
    inits = np.stack([synth.pose_matrix(synth.rotvec_to_matrix([0, 0, a]), [0, 0, 0]) for a in (0.0, 2.0, 0.0)])
    Ts, fit, rmse, iters, ch = _oracle_multistart(s5, t5, inits, 20.0, 6, 1e-6, 1e-6)
    for r in (0, 1):
        m = got[r]["multistart"]
        np.testing.assert_array_equal(m["transformations"], np.stack(Ts))
        np.testing.assert_array_equal(m["chamfer"], ch)
        np.testing.assert_array_equal(m["iterations"], iters)
        assert ch[0] == ch[2] and m["best"] == int(np.argmin(ch)) == m["order"][0]
        assert list(m["order"]) == list(np.argsort(ch, kind="stable"))
