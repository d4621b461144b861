"""Multi-rank parity of the sharded paths against the oracle and the single-GPU results.

  test_two_ranks_match_oracle[gloo]   world 2 on ONE device: two processes share cuda:0, gloo
                              carries the host-side collectives, and the kernel-fused exchange of the
                              sharded ICP runs through CUDA IPC between the two processes -- so
                              isr_icp_run_sharded, target-sharded ICP, the candidate split and the
                              multi-start split are parity-checked on the driver's 1-GPU box
  test_two_ranks_match_oracle[nccl]   the same with NCCL, one process per GPU (generated only where
                              >= 2 GPUs are visible: `gpurun --gpus 2 -- python -m pytest tests -m gpu`)"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from oracle import oracle

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q, backend):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world), LOCAL_RANK=str(rank if backend == "nccl" else 0))
    import torch.distributed as td
    from imagesequenceregistrationfor6dposeestimationlabeling_b200 import dist, synth
    dist.init_from_env(backend)
    out = {}
    try:
        cloud = synth.make_cloud(6000, seed=1)
        R_true, _ = synth.true_pose(3)
        Rs, _, k0 = synth.make_candidates(21, seed=10, R_true=R_true, t_true=np.zeros(3))
        Mq, Mt = synth.verification_matrices(Rs, R_true)
        bi, bl, losses = dist.verify_poses_sharded(cloud, Mq, Mt, gather_losses=True)
        out["verify"] = (int(bi.item()), float(bl.item()), losses.cpu().numpy(), k0)
        src, tgt, _ = synth.icp_pair(9001, 9500, 6, 7)
        res = dist.icp_sharded(src, tgt, np.eye(4), 20.0, max_iteration=12)
        out["icp"] = (res[0].transformation, res[0].fitness, res[0].inlier_rmse, res[0].iterations)
        res = dist.icp_sharded(src, tgt, np.eye(4), 20.0, max_iteration=12, exchange="nccl")
        out["icp_nccl"] = (res[0].transformation, res[0].fitness, res[0].inlier_rmse, res[0].iterations)
        inits = np.stack([synth.pose_matrix(synth.rotvec_to_matrix([0, 0, 2 * np.pi * k / 5]), [0, 0, 0])
                          for k in range(5)])
        res = dist.icp_sharded(src, tgt, inits, 20.0, max_iteration=30, relative_fitness=1e-4,
                               relative_rmse=1e-3)
        out["icp_multi"] = [(r.transformation, r.fitness, r.inlier_rmse, r.iterations) for r in res]
        res = dist.icp_sharded(src, tgt, inits, 20.0, max_iteration=30, relative_fitness=1e-4,
                               relative_rmse=1e-3, exchange="nccl")
        out["icp_multi_nccl"] = [(r.transformation, r.fitness, r.inlier_rmse, r.iterations) for r in res]
        out["multistart"] = dist.multistart_icp_sharded(src, tgt, inits, 20.0, max_iteration=15)
        res = dist.icp_sharded(src, tgt, np.eye(4), 20.0, max_iteration=12, shard="target")
        out["icp_target"] = (res[0].transformation, res[0].fitness, res[0].inlier_rmse, res[0].iterations)
        torch.cuda.synchronize()
    finally:
        dist.close_peer_exchanges()
        td.destroy_process_group()
    q.put((rank, out))


# gloo: two ranks share cuda:0 (runs on every GPU box); nccl: one rank per GPU -- NCCL refuses two
# ranks on one device, so that form only exists where two GPUs are visible
_BACKENDS = ["gloo"] + (["nccl"] if torch.cuda.device_count() >= 2 else [])


@pytest.mark.parametrize("backend", _BACKENDS)
def test_two_ranks_match_oracle(backend):
    _run_two_ranks(backend)


def _run_two_ranks(backend):
    from imagesequenceregistrationfor6dposeestimationlabeling_b200 import synth
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q, backend)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=600) for _ in procs)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    cloud = synth.make_cloud(6000, seed=1)
    R_true, _ = synth.true_pose(3)
    Rs, _, k0 = synth.make_candidates(21, seed=10, R_true=R_true, t_true=np.zeros(3))
    Mq, Mt = synth.verification_matrices(Rs, R_true)
    ref, ref_best = oracle.verify_matrices(cloud, cloud, Mq, Mt)
    src, tgt, _ = synth.icp_pair(9001, 9500, 6, 7)
    o = oracle.registration_icp(src, tgt, 20.0, np.eye(4), max_iteration=12)
    for r in (0, 1):
        bi, bl, losses, kk = got[r]["verify"]
        assert bi == ref_best == k0
        np.testing.assert_allclose(losses, ref, rtol=1e-5)
        np.testing.assert_allclose(bl, ref[ref_best], rtol=1e-5)
        for key in ("icp", "icp_nccl", "icp_target"):
            T, fit, rmse, it = got[r][key]
            np.testing.assert_allclose(T, o.transformation, rtol=1e-7, atol=1e-7)
            assert it == o.iterations and abs(fit - o.fitness) < 1e-12
            np.testing.assert_allclose(rmse, o.inlier_rmse, rtol=1e-7)
    np.testing.assert_array_equal(got[0]["icp"][0], got[1]["icp"][0])   # bit-identical ranks
    # the kernel-fused peer exchange (default) adds the two ranks' sums in rank order, as the
    # 2-rank all-reduce does: same bits
    np.testing.assert_array_equal(got[0]["icp"][0], got[0]["icp_nccl"][0])
    for r in (0, 1):
        assert len(got[r]["icp_multi"]) == 5
        for a, b, c in zip(got[r]["icp_multi"], got[r]["icp_multi_nccl"], got[0]["icp_multi"]):
            np.testing.assert_array_equal(a[0], c[0])
            np.testing.assert_allclose(a[0], b[0], rtol=1e-9, atol=1e-9)
            assert a[3] == b[3] and a[1] == b[1]
    np.testing.assert_array_equal(got[0]["icp_target"][0], got[1]["icp_target"][0])
    # config 5 split by starts (3 + 2): equals the single-GPU batched run, on both ranks
    from imagesequenceregistrationfor6dposeestimationlabeling_b200 import api
    inits = np.stack([synth.pose_matrix(synth.rotvec_to_matrix([0, 0, 2 * np.pi * k / 5]), [0, 0, 0])
                      for k in range(5)])
    ms = api.multistart_icp(src, tgt, inits, 20.0, max_iteration=15)
    for r in (0, 1):
        m = got[r]["multistart"]
        np.testing.assert_array_equal(m["transformations"], np.stack([x.transformation for x in ms.results]))
        np.testing.assert_array_equal(m["chamfer"], ms.chamfer)
        assert list(m["order"]) == list(ms.order) and m["best"] == int(ms.order[0])
    np.testing.assert_array_equal(got[0]["verify"][2], got[1]["verify"][2])
