"""Generate golden vectors by EXECUTING THE REFERENCE'S OWN FUNCTION BODIES.

Run in the build container only (needs /root/reference; the GPU box has none):
    python tests/golden/make_golden.py

The reference scripts cannot be imported (module-level argparse, missing dep.* modules,
open3d / trimesh / pytorch3d absent), so the four pure helper functions on the hot path
are lifted out of the source with ``ast`` at generation time and executed unchanged:

    choosePose.py : ADD (18-19), ADDS (20-22), compute_rel_poses (43-51)
    verfication.py: calculate_relative_pose (9-19)

Nothing is copied into this repository -- only the inputs and the reference's outputs are
stored, in tests/golden/reference_helpers.npz.  The Open3D calls of the hot path cannot be
executed here (parity unpinned for that half; see oracle/oracle.py).
"""
import ast
import os
import sys

import numpy as np
from sklearn.neighbors import KDTree

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))


def lift(path, names):
    src = open(path).read()
    tree = ast.parse(src)
    ns = {"np": np, "KDTree": KDTree}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in names:
            exec(compile(ast.Module(body=[node], type_ignores=[]), path, "exec"), ns)
    missing = [n for n in names if n not in ns]
    assert not missing, missing
    return ns


def main():
    from imagesequenceregistrationfor6dposeestimationlabeling_b200 import synth

    cp = lift(os.path.join(REF, "choosePose.py"), ["ADD", "ADDS", "compute_rel_poses"])
    vf = lift(os.path.join(REF, "verfication.py"), ["calculate_relative_pose"])
    rng = np.random.default_rng(2024)
    surface = synth.make_cloud(2000, seed=21).astype(np.float64)
    verts = synth.make_cloud(500, seed=22).astype(np.float64)
    cp["surfacePointsScaled"] = surface  # the module global ADDS reads (choosePose.py:21,160)
    n = 6
    R = np.stack([synth.random_rotation(rng) for _ in range(n)])
    t = rng.normal(scale=30.0, size=(n, 3)) + np.array([0, 0, 700.0])
    out = {"surface": surface, "verts": verts, "R": R, "t": t}
    add, adds, relR, relt, calR, calT = [], [], [], [], [], []
    for i in range(n):
        for j in range(n):
            add.append(cp["ADD"](verts, R[i], t[i], R[j], t[j]))
            adds.append(cp["ADDS"](verts, R[i], t[i], R[j], t[j]))
            r, tt = cp["compute_rel_poses"](R[i], t[i], R[j], t[j])
            relR.append(r)
            relt.append(tt)
            r, tt = vf["calculate_relative_pose"](R[i], t[i], R[j], t[j])
            calR.append(r)
            calT.append(tt)
    out.update(ADD=np.array(add).reshape(n, n), ADDS=np.array(adds).reshape(n, n),
               rel_R=np.array(relR).reshape(n, n, 3, 3), rel_t=np.array(relt).reshape(n, n, 3),
               cal_R=np.array(calR).reshape(n, n, 3, 3), cal_T=np.array(calT).reshape(n, n, 3))
    # small perturbation poses too (ADD-S near the 0.1*diameter decision, choosePose.py:135)
    Rp = np.stack([R[0] @ synth.rotvec_to_matrix(rng.normal(scale=s, size=3))
                   for s in (0.001, 0.01, 0.05, 0.1, 0.3)])
    tp = t[0] + rng.normal(scale=1.0, size=(5, 3))
    out.update(Rp=Rp, tp=tp,
               ADDS_p=np.array([cp["ADDS"](verts, R[0], t[0], Rp[k], tp[k]) for k in range(5)]),
               ADD_p=np.array([cp["ADD"](verts, R[0], t[0], Rp[k], tp[k]) for k in range(5)]))
    path = os.path.join(HERE, "reference_helpers.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
