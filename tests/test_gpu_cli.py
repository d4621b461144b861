"""Drop-in CLIs on synthetic copies of the reference's file layout, against the oracle."""
import json
import os

import numpy as np
import pytest

from oracle import oracle

pytestmark = pytest.mark.gpu


def _write_ply(path, v):
    with open(path, "wb") as f:
        f.write(("ply\nformat binary_little_endian 1.0\nelement vertex %d\nproperty float x\n"
                 "property float y\nproperty float z\nend_header\n" % len(v)).encode())
        f.write(np.asarray(v, dtype="<f4").tobytes())


def test_verfication_cli(gpu, tmp_path):
    from imagesequenceregistrationfor6dposeestimationlabeling_b200 import synth
    from imagesequenceregistrationfor6dposeestimationlabeling_b200.cli import verfication
    rng = np.random.default_rng(0)
    n = 7
    gt_R = [synth.random_rotation(rng) for _ in range(n)]
    gt_T = [rng.normal(scale=50, size=3) + [0, 0, 700] for _ in range(n)]
    pred_R = [gt_R[i] @ synth.rotvec_to_matrix(rng.normal(scale=0.05, size=3)) for i in range(n)]
    pc1 = synth.make_cloud(5000, seed=4)
    os.makedirs(tmp_path / "bop/Tless/train/000015")
    os.makedirs(tmp_path / "Tless/15poseEst_UH0")
    json.dump({str(i * 3): [{"cam_R_m2c": gt_R[i].reshape(-1).tolist(), "cam_t_m2c": gt_T[i].tolist()}]
               for i in range(n)}, open(tmp_path / "bop/Tless/train/000015/scene_gt.json", "w"))
    json.dump({str(i * 3): [{"R": pred_R[i].reshape(-1).tolist(), "T": gt_T[i].tolist()}] for i in range(n)},
              open(tmp_path / "Tless/15poseEst_UH0/pred6d.json", "w"))
    np.save(tmp_path / "Tless/15poseEst_UH0/vert1_scaled.npy", pc1)
    lst, idx, mn = verfication.main(["--objid", "15", "--root", str(tmp_path)])
    ref, ridx, rmn = oracle.verify_chamfer(pc1, gt_R, gt_T, pred_R, gt_T)
    np.testing.assert_allclose(lst, ref, rtol=1e-5)
    assert idx == ridx
    np.testing.assert_allclose(mn, rmn, rtol=1e-5)


def test_icp_and_choose_pose_cli(gpu, tmp_path):
    from imagesequenceregistrationfor6dposeestimationlabeling_b200 import synth
    from imagesequenceregistrationfor6dposeestimationlabeling_b200.cli import choose_pose, icp
    rng = np.random.default_rng(1)
    upper = synth.make_cloud(6000, seed=1, half="upper")
    lower = synth.make_cloud(6000, seed=2, half="lower")
    cad = synth.make_cloud(3000, seed=3)
    n = 5
    R = [synth.random_rotation(rng) for _ in range(n)]
    t = [rng.normal(scale=20, size=3) + [0, 0, 700] for _ in range(n)]
    Rp = [R[i] @ synth.rotvec_to_matrix(rng.normal(scale=0.02, size=3)) for i in range(n)]
    tp = [t[i] + rng.normal(scale=1.0, size=3) for i in range(n)]
    tp[3] = tp[3] + np.array([0, 0, 90.0])            # image 3's prediction is wrong
    for uh in ("0", "1"):
        os.makedirs(tmp_path / f"{uh}_ruapc_obj_1/1poseEst")
    os.makedirs(tmp_path / "bop/ruapc/models")
    os.makedirs(tmp_path / "bop/ruapc/train/000001")
    np.save(tmp_path / "1_ruapc_obj_1/1poseEst/vert1_scaled.npy", upper)
    np.save(tmp_path / "0_ruapc_obj_1/1poseEst/vert1_scaled.npy", lower)
    np.save(tmp_path / "0_ruapc_obj_1/1pred_R.npy", np.stack(Rp))
    np.save(tmp_path / "0_ruapc_obj_1/1pred_t.npy", np.stack(tp))
    json.dump({"1": {"diameter": 120.0}}, open(tmp_path / "bop/ruapc/models/models_info.json", "w"))
    json.dump({str(i): [{"cam_R_m2c": R[i].reshape(-1).tolist(), "cam_t_m2c": t[i].tolist()}] for i in range(n)},
              open(tmp_path / "bop/ruapc/train/000001/scene_gt.json", "w"))
    _write_ply(tmp_path / "bop/ruapc/models/obj_000001.ply", cad)
    _write_ply(tmp_path / "bop/ruapc/models/obj_000001.ply".replace("obj_000001", "obj_000001"), cad)
    # choosePose reads obj_0000<id:02d>.ply
    _write_ply(tmp_path / "bop/ruapc/models/obj_000001.ply", cad)

    root = ["--root", str(tmp_path), "--dataset", "ruapc", "--objid", "1"]
    out = choose_pose.main(root + ["--rel_poses", "1", "--cal_GT", "1", "--cal_pred", "1"])
    np.testing.assert_allclose(out["gt_relative_poses"], oracle.rel_pose_table(R, t), atol=1e-12)
    out = choose_pose.main(root + ["--choose_image", "1"])
    err, image_id, top = oracle.choose_image(oracle.rel_pose_table(Rp, tp), oracle.rel_pose_table(R, t),
                                             cad.astype(np.float64), lower.astype(np.float64), 120.0)
    np.testing.assert_array_equal(out["error"], err)
    assert out["image_id"] == image_id and list(out["top"]) == list(top)
    chosen = [int(x) for x in open(tmp_path / "0_ruapc_obj_1/1top_50_choices.txt")]
    assert chosen == list(top) and chosen[-1] == 3

    ev, reg, ch = icp.main(root)
    i0 = chosen[0]
    ev_o, reg_o, ch_o = oracle.icp_script(upper, lower, R[i0], t[i0], Rp[i0], tp[i0], cad.astype(np.float64))
    assert ev.fitness == ev_o.fitness
    np.testing.assert_allclose(ev.inlier_rmse, ev_o.inlier_rmse, rtol=1e-7)
    np.testing.assert_allclose(reg.transformation, reg_o.transformation, rtol=1e-6, atol=1e-6)
    assert reg.iterations == reg_o.iterations
    np.testing.assert_allclose(ch, ch_o, rtol=1e-5)
