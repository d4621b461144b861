"""verfication.py as one batched call: Chamfer-verify consecutive predicted relative poses.

Files (verfication.py:40-56): bop/Tless/train/<objid:06d>/scene_gt.json,
Tless/<objid>poseEst_UH0/pred6d.json, Tless/<objid>poseEst_UH0/vert1_scaled.npy."""
from __future__ import annotations

import argparse
import json
import os

import numpy as np

from .. import api, helpers


def build_matrices(data, pred6d):
    """The loop of verfication.py:61-85 as two pose arrays (rotations only, translations off)."""
    keysgt = sorted(data.keys(), key=lambda x: int(x))
    keyspred = sorted(pred6d.keys(), key=lambda x: int(x))
    n = len(keyspred) - 1
    Mq = np.tile(np.eye(4), (n, 1, 1))
    Mt = np.tile(np.eye(4), (n, 1, 1))
    for i in range(n):
        R1 = np.array(data[keysgt[i]][0]["cam_R_m2c"]).reshape(3, 3)
        T1 = np.array(data[keysgt[i]][0]["cam_t_m2c"])
        R2 = np.array(data[keysgt[i + 1]][0]["cam_R_m2c"]).reshape(3, 3)
        T2 = np.array(data[keysgt[i + 1]][0]["cam_t_m2c"])
        R_rel, _ = helpers.calculate_relative_pose(R1, T1, R2, T2)           # :75
        R1pred = np.array(pred6d[keyspred[i]][0]["R"]).reshape(3, 3)         # :77
        R2pred = np.array(pred6d[keyspred[i + 1]][0]["R"]).reshape(3, 3)     # :79
        Mq[i, :3, :3] = R2pred.T             # pcpred = pc1.dot(R2pred)                  :85
        Mt[i, :3, :3] = R_rel.T @ R1pred     # pcgt = pc1.dot(R1pred.T).dot(R_relative)  :83-84
    return Mq, Mt


def main(argv=None):
    ap = argparse.ArgumentParser(description="Train a Linemod")  # the reference's own text
    ap.add_argument("--objid", dest="objid", default="15")
    ap.add_argument("--root", default=".", help="directory holding bop/ and Tless/")
    args = ap.parse_args(argv)
    objid, UH = str(args.objid), "0"
    with open(os.path.join(args.root, "bop/Tless/train", objid.zfill(6), "scene_gt.json")) as f:
        data = json.load(f)
    base = os.path.join(args.root, "Tless", objid + "poseEst_UH" + UH)
    with open(os.path.join(base, "pred6d.json")) as f:
        pred6d = json.load(f)
    pc1 = np.load(os.path.join(base, "vert1_scaled.npy"))
    Mq, Mt = build_matrices(data, pred6d)
    res = api.verify_poses(pc1, Mq, Mt, mode="chamfer")
    chamferdis = res.losses.cpu().numpy().tolist()
    min_index, min_chamfer = res.best_index, res.best_loss
    print("best image", min_index)
    print("min chamfer distance", min_chamfer)
    return chamferdis, min_index, min_chamfer


if __name__ == "__main__":
    main()
