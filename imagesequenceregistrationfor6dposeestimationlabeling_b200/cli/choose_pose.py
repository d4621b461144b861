"""choosePose.py --rel_poses / --choose_image: all-pairs relative poses and the ADD-S vote.

Files (choosePose.py:69-72,95-96,111-120,141-150): bop/<ds>/models/models_info.json,
bop/<ds>/models/obj_0000<id:02d>.ply, <UH>_<ds>_obj_<id>/<id>pred_{R,t}.npy,
.../<id>{gt,pred}_relative_poses.npy, .../<id>poseEst/vert1_scaled.npy (surface points);
writes agreedposes.npy, error.npy and <id>top_50_choices.txt.  Ground-truth poses for
--cal_GT come from scene_gt.json in image-id order (the reference walks the depth/ folder,
choosePose.py:77-90)."""
from __future__ import annotations

import argparse
import json
import os

import numpy as np

from .. import helpers
from ..o3d_compat.io import read_ply_vertices


def _table(RList, TList):
    """The n x n x 4 x 4 table of choosePose.py:98-107, built on the device (isr_rel_pose_table)
    and brought back for np.save."""
    from .. import api

    n = len(TList)
    return api.relative_pose_table(np.asarray(RList, dtype=np.float64), np.asarray(TList, dtype=np.float64)
                                   ).cpu().numpy().reshape(n, n, 4, 4)


def main(argv=None):
    ap = argparse.ArgumentParser(description="Train a Linemod")
    ap.add_argument("--objid", dest="objid", default="2")
    ap.add_argument("--cal_GT", dest="cal_GT", default=0)
    ap.add_argument("--cal_pred", dest="cal_pred", default=0)
    ap.add_argument("--rel_poses", dest="rel_poses", default=0)
    ap.add_argument("--choose_image", dest="choose_image", default=0)
    ap.add_argument("--dataset", dest="dataset", default="tless")
    ap.add_argument("--UH", dest="UH", default=0)
    ap.add_argument("--root", default=".")
    ap.add_argument("--limit", type=int, default=1280, help="images used (choosePose.py:81)")
    args = ap.parse_args(argv)
    oid, ds, root = str(args.objid), str(args.dataset), args.root
    exp = os.path.join(root, str(args.UH) + "_" + ds + "_obj_" + oid)
    out = {}
    if int(args.rel_poses):
        if int(args.cal_GT):
            with open(os.path.join(root, "bop", ds, "train", oid.zfill(6), "scene_gt.json")) as f:
                gt = json.load(f)
            keys = sorted(gt.keys(), key=lambda x: int(x))[:args.limit]
            RList = [np.asarray(gt[k][0]["cam_R_m2c"]).reshape(3, 3) for k in keys]
            TList = [np.asarray(gt[k][0]["cam_t_m2c"]) for k in keys]
            rel = _table(RList, TList)                                         # :98-107
            np.save(os.path.join(exp, oid + "gt_relative_poses.npy"), rel)
            out["gt_relative_poses"] = rel
        if int(args.cal_pred):
            RList = np.load(os.path.join(exp, oid + "pred_R.npy"), allow_pickle=True)
            TList = np.load(os.path.join(exp, oid + "pred_t.npy"), allow_pickle=True)
            rel = _table(np.stack(list(RList)), np.stack(list(TList)))
            np.save(os.path.join(exp, oid + "pred_relative_poses.npy"), rel)
            out["pred_relative_poses"] = rel
    if int(args.choose_image):
        with open(os.path.join(root, "bop", ds, "models", "models_info.json")) as f:
            diameter = json.load(f)[oid]["diameter"]
        modelVerts = read_ply_vertices(os.path.join(root, "bop", ds, "models", "obj_0000" + oid.zfill(2) + ".ply"))
        surface = np.load(os.path.join(exp, oid + "poseEst", "vert1_scaled.npy"))
        pred = np.load(os.path.join(exp, oid + "pred_relative_poses.npy"))
        gtr = np.load(os.path.join(exp, oid + "gt_relative_poses.npy"))
        error, image_id, top = helpers.choose_image(pred, gtr, modelVerts, diameter, surface_points=surface)
        agreed = np.argwhere(error > 0)
        np.save(os.path.join(exp, oid + "agreedposes.npy"), agreed)
        np.save(os.path.join(exp, oid + "error.npy"), error)
        with open(os.path.join(exp, oid + "top_50_choices.txt"), "w") as f:
            for item in top:
                f.write(str(item) + "\n")
        print("image which should be chosen is ", image_id)
        out.update(error=error, image_id=image_id, top=top)
    return out


if __name__ == "__main__":
    main()
