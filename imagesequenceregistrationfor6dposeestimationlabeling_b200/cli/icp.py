"""icp.py: ICP refinement of the second-sequence cloud + final Chamfer against the CAD model.

Files (icp.py:37-65,112): 0_<ds>_obj_<id>/<id>top_50_choices.txt, bop/<ds>/models/models_info.json,
{0,1}_<ds>_obj_<id>/<id>poseEst/vert1_scaled.npy, 0_<ds>_obj_<id>/<id>pred_{R,t}.npy,
bop/<ds>/train/<id:06d>/scene_gt.json, bop/<ds>/models/obj_<id:06d>.ply."""
from __future__ import annotations

import argparse
import json
import os

import numpy as np

from .. import api
from ..o3d_compat.io import read_ply_vertices


def main(argv=None):
    ap = argparse.ArgumentParser(description="Train a Linemod")
    ap.add_argument("--objid", dest="objid", default="1")
    ap.add_argument("--dataset", dest="dataset", default="ruapc")
    ap.add_argument("--root", default=".")
    args = ap.parse_args(argv)
    ds, oid, root = str(args.dataset), str(args.objid), args.root
    d0 = os.path.join(root, "0_" + ds + "_obj_" + oid)
    d1 = os.path.join(root, "1_" + ds + "_obj_" + oid)
    with open(os.path.join(d0, oid + "top_50_choices.txt")) as f:
        id_chosen = [int(line.strip()) for line in f][0]                      # icp.py:37-39
    with open(os.path.join(root, "bop", ds, "models", "models_info.json")) as f:
        diam = json.load(f)[oid]["diameter"]
    upper = np.load(os.path.join(d1, oid + "poseEst", "vert1_scaled.npy")).astype("float32")
    lower = np.load(os.path.join(d0, oid + "poseEst", "vert1_scaled.npy")).astype("float32")
    R_pred = np.load(os.path.join(d0, oid + "pred_R.npy"), allow_pickle=True)[id_chosen]
    t_pred = np.load(os.path.join(d0, oid + "pred_t.npy"), allow_pickle=True)[id_chosen]
    with open(os.path.join(root, "bop", ds, "train", oid.zfill(6), "scene_gt.json")) as f:
        gt = json.load(f)
    R_GT = np.array(gt[str(id_chosen)][0]["cam_R_m2c"]).reshape(3, 3)
    t_GT = np.array(gt[str(id_chosen)][0]["cam_t_m2c"])
    actual_upper = upper.dot(R_GT.T) + t_GT                                   # :68 (float64)
    init = np.linalg.inv(api.pose_from_Rt(np.asarray(R_pred, dtype=np.float64), t_pred))  # :88-92
    threshold = 20
    print("Initial alignment")
    evaluation = api.evaluate_registration(actual_upper, lower, threshold, init)  # :97-98
    print(evaluation)
    print("Apply point-to-point ICP")
    reg = api.icp(actual_upper, lower, init, threshold)                       # :101-103
    print(reg)
    print("Transformation is:")
    print(reg.transformation)
    T = reg.transformation
    merged = np.concatenate([actual_upper @ T[:3, :3].T + T[:3, 3], lower.astype(np.float64)])  # :110-111
    cad = read_ply_vertices(os.path.join(root, "bop", ds, "models", "obj_" + oid.zfill(6) + ".ply"))
    chamfer_distance = float(api.chamfer_distance(merged, cad))                # :113-117
    print("diameter", diam)
    print("Chamfer Distance(final):", chamfer_distance)
    print("final transformation matrix between first and second sequence is: \n", T)
    return evaluation, reg, chamfer_distance


if __name__ == "__main__":
    main()
