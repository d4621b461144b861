"""Host-side logic and the C-ABI surface, without a GPU: helper functions against the
reference's golden vectors, the Open3D/PLY shim pieces that never touch the device, the
shard arithmetic, and that libisr.so loads and exports every symbol include/isr.h declares.
No compute call is made."""
import copy
import ctypes
import os
import re

import numpy as np
import pytest

import imagesequenceregistrationfor6dposeestimationlabeling_b200 as isr
from imagesequenceregistrationfor6dposeestimationlabeling_b200 import _lib, dist, helpers, synth
import imagesequenceregistrationfor6dposeestimationlabeling_b200.o3d_compat as o3d
from oracle import oracle

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_helpers.npz"))


def test_helpers_match_reference_golden():
    n = len(GOLD["R"])
    for i in range(n):
        for j in range(n):
            r, t = helpers.compute_rel_poses(GOLD["R"][i], GOLD["t"][i], GOLD["R"][j], GOLD["t"][j])
            np.testing.assert_array_equal(r, GOLD["rel_R"][i, j])
            np.testing.assert_array_equal(t, GOLD["rel_t"][i, j])
            r, t = helpers.calculate_relative_pose(GOLD["R"][i], GOLD["t"][i], GOLD["R"][j], GOLD["t"][j])
            np.testing.assert_array_equal(r, GOLD["cal_R"][i, j])
            np.testing.assert_array_equal(t, GOLD["cal_T"][i, j])


def test_relative_pose_table_vectorised_equals_loop():
    tab = helpers.relative_pose_table(GOLD["R"], GOLD["t"])
    ref = oracle.rel_pose_table(GOLD["R"], GOLD["t"])
    np.testing.assert_allclose(tab, ref, rtol=0, atol=1e-15)
    np.testing.assert_array_equal(tab[:, :, :3, 3], GOLD["rel_t"])


def test_header_symbols_all_exported_and_bound():
    header = open(_lib.HEADER_PATH).read()
    declared = set(re.findall(r"\b(isr_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations found in include/isr.h"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in isr.h but not exported by libisr.so"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert _lib.load().isr_version() == 100


def test_abi_constants_and_struct_layout():
    lib = _lib.load()
    assert lib.isr_soa_padded_len(1) == 1024 and lib.isr_soa_padded_len(1024) == 1024
    assert lib.isr_soa_padded_len(100000) == 100352 == _lib.soa_padded_len(100000)
    assert _lib.ICP_STATE_DTYPE.itemsize == 184
    assert _lib.ICP_STATE_DTYPE.fields["done"][1] == 176
    header = open(_lib.HEADER_PATH).read()
    assert "#define ISR_SOA_TILE 1024" in header and "#define ISR_ICP_NSUMS 17" in header
    assert lib.isr_verify_workspace_bytes(100000, 100000, 1000, 1) > 2 * 3 * 100352 * 4
    assert lib.isr_icp_workspace_bytes(1000, 1000, 1) % 256 == 0


def test_argument_validation_needs_no_gpu():
    lib = _lib.load()
    assert lib.isr_nn_soa(None, 5, 1024, 0, None, 5, 1024, 0, 1, None, None, None, 0, None, 0, None) == -1
    assert b"null pointer" in lib.isr_last_error()
    assert lib.isr_nn_soa(None, 5, 1000, 0, None, 0, 1024, 0, 1, None, None, None, 0, None, 0, None) == -2
    assert lib.isr_transform_points(None, -1, None, 1, None, None) == -2
    assert lib.isr_verify_poses(None, 0, None, 5, None, None, None, 1, 1, None, None, None, 0, None) == -2
    assert lib.isr_icp_run(None, 1, None, None, None, 1, None, None, None, 20.0, -1, 0.0, 0.0, None, None,
                           None, None, 0, None) == -1
    assert lib.isr_nn2(None, None, 1, 1, None, None, None, 0, None, 0, None) == -1
    empty = _lib.IsrCloud(None, 5, 1024, 0, None, None, None, None)
    assert lib.isr_nn2(ctypes.byref(empty), ctypes.byref(empty), 1, 1, None, None, None, 0, None, 0,
                       None) == -1
    assert lib.isr_prepare_cloud(None, None, None, 5, None, 16, None, 16, None, 2, None, 1024, None, 0,
                                 None) == -1
    assert lib.isr_spatial_order(None, 5, None, None, 0, None) == -1
    assert lib.isr_spatial_order_workspace_bytes(100000) >= 131072 * 8
    # stage spheres + one chunk sphere per 32 stages
    assert [lib.isr_stage_sphere_count(1024 * k) for k in (1, 32, 33, 977)] == [2, 33, 35, 977 + 31]
    assert lib.isr_pnp_score(None, None, 5, None, None, 3, 2.0, None, None, None) == -1
    assert lib.isr_pnp_score(None, None, -1, None, None, 3, 2.0, None, None, None) == -2
    assert lib.isr_peer_create(3, 2, None, None) == -1
    assert lib.isr_icp_run_sharded(None, 1, None, None, None, 1, 1, None, None, None, 20.0, 1, 0.0, 0.0, None,
                                   None, None, None, 0, None, None) == -1
    with pytest.raises(_lib.IsrError):
        _lib.check(-3)


def test_product_path_has_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        isr.chamfer_distance(np.zeros((4, 3)), np.zeros((4, 3)))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        isr.icp(np.zeros((4, 3)), np.zeros((4, 3)))
    pc = o3d.geometry.PointCloud(np.zeros((4, 3)))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        pc.compute_point_cloud_distance(pc)


def test_product_never_imports_the_oracle():
    root = os.path.dirname(_lib.__file__)
    for dp, _, files in os.walk(root):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert "oracle" not in src.replace("the CPU oracle", "").replace("oracle-backed", "") \
                    .replace("CPU oracle", "") or f in ("dist.py",), f


def test_pointcloud_shim_host_semantics():
    a = np.arange(12, dtype=np.float64).reshape(4, 3)
    pc = o3d.geometry.PointCloud()
    pc.points = o3d.utility.Vector3dVector(a)
    T = synth.pose_matrix(synth.rotvec_to_matrix([0.1, 0.2, 0.3]), [1, 2, 3])
    out = pc.transform(T)
    assert out is pc                                              # in place, returns self
    np.testing.assert_allclose(np.asarray(pc.points), oracle.transform(a, T), rtol=1e-15)
    other = o3d.geometry.PointCloud(np.ones((3, 3)))
    s = pc + other
    assert len(s) == 7 and len(pc) == 4
    c = copy.deepcopy(pc)
    c.paint_uniform_color([1, 0.706, 0]).transform(np.eye(4))
    assert np.asarray(c.colors).shape == (4, 3)
    np.testing.assert_array_equal(np.asarray(c.points), np.asarray(pc.points))
    assert o3d.visualization.draw_geometries([pc]) is None
    assert helpers.draw_registration_result(pc, other, np.eye(4)) is None and helpers.vp(a) is None
    with pytest.raises(RuntimeError):
        o3d.utility.Vector3dVector(np.zeros((3, 2)))
    e = o3d.geometry.PointCloud()
    assert len(np.asarray(e.compute_point_cloud_distance(pc))) == 0
    assert np.all(np.asarray(pc.compute_point_cloud_distance(e)) == 0)      # empty target -> zeros
    r = o3d.pipelines.registration.registration_icp(e, pc, 20, np.eye(4))
    assert r.fitness == 0 and r.inlier_rmse == 0 and len(r.correspondence_set) == 0
    crit = o3d.pipelines.registration.ICPConvergenceCriteria()
    assert (crit.max_iteration, crit.relative_fitness, crit.relative_rmse) == (30, 1e-6, 1e-6)


@pytest.mark.parametrize("fmt", ["ascii", "binary_little_endian", "binary_big_endian"])
def test_ply_reader(tmp_path, fmt):
    rng = np.random.default_rng(0)
    v = rng.normal(size=(17, 3)).astype(np.float32)
    nrm = rng.normal(size=(17, 3)).astype(np.float32)
    path = tmp_path / f"obj_{fmt}.ply"
    hdr = (f"ply\nformat {fmt} 1.0\ncomment test\nelement vertex 17\nproperty float x\nproperty float y\n"
           "property float z\nproperty float nx\nproperty float ny\nproperty float nz\n"
           "element face 1\nproperty list uchar int vertex_indices\nend_header\n")
    with open(path, "wb") as f:
        f.write(hdr.encode())
        if fmt == "ascii":
            for a, b in zip(v, nrm):
                f.write((" ".join(repr(float(x)) for x in (*a, *b)) + "\n").encode())
            f.write(b"3 0 1 2\n")
        else:
            end = "<" if fmt.endswith("little_endian") else ">"
            rec = np.concatenate([v, nrm], axis=1).astype(end + "f4")
            f.write(rec.tobytes())
            f.write(np.array([3], dtype="u1").tobytes() + np.array([0, 1, 2], dtype=end + "i4").tobytes())
    pc = o3d.io.read_point_cloud(str(path))
    np.testing.assert_array_equal(np.asarray(pc.points), v.astype(np.float64))
    with pytest.raises(RuntimeError):
        o3d.io.read_point_cloud(str(tmp_path / "x.pcd"))


def test_shard_bounds_cover_exactly_once():
    for n in (0, 1, 7, 1000, 10000):
        for w in (1, 2, 3, 4, 8):
            seen = []
            for r in range(w):
                lo, hi = dist.shard_bounds(n, r, w)
                assert 0 <= lo <= hi <= n
                seen.extend(range(lo, hi))
            assert seen == list(range(n))


def test_synthetic_generators_are_seeded_and_shaped():
    a, b = synth.make_cloud(5000, 1), synth.make_cloud(5000, 1)
    np.testing.assert_array_equal(a, b)
    assert a.dtype == np.float32 and a.shape == (5000, 3)
    assert 100 < np.ptp(a[:, 0]) < 125 and 50 < np.ptp(a[:, 2]) < 65
    up = synth.make_cloud(3000, 2, half="upper")
    lo = synth.make_cloud(3000, 2, half="lower")
    assert up[:, 2].min() > -6.0 and lo[:, 2].max() < 6.0 and len(up) == len(lo) == 3000
    Rs, ts, k0 = synth.make_candidates(50, 10)
    assert Rs.shape == (50, 3, 3) and ts.shape == (50, 3) and 0 <= k0 < 50
    np.testing.assert_allclose(Rs @ np.transpose(Rs, (0, 2, 1)), np.tile(np.eye(3), (50, 1, 1)), atol=1e-12)
    src, tgt, T = synth.icp_pair(100, 120, 4, 5)
    assert src.shape == (100, 3) and tgt.shape == (120, 3) and T.shape == (4, 4)


def test_integration_doc_names_every_export():
    """INTEGRATION.md's entry-point map covers the whole header (a *_workspace_bytes companion counts
    as named when its function is)."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    header = open(_lib.HEADER_PATH).read()
    doc = open(os.path.join(root, "INTEGRATION.md")).read()
    declared = set(re.findall(r"\b(isr_[a-z0-9_]+)\s*\(", header))
    missing = sorted(n for n in declared
                     if n not in doc and n.replace("_workspace_bytes", "") not in doc)
    assert not missing, missing


def test_committed_bench_line_keeps_the_contract():
    """The bench line of the final build under profiles/ parses and carries the contract's keys."""
    import json
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    line = json.load(open(os.path.join(root, "profiles", "r02d_bench.json")))
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "clocks",
                "roofline", "cpu_baseline"):
        assert key in line, key
    assert line["config"]["workload"] and line["gpu_launches"] > 0 and line["warmup"] >= 3
    assert set(("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step")) <= set(line["e2e"])
    assert line["e2e"]["h2d_bytes_per_step"] > 0 and line["e2e"]["d2h_bytes_per_step"] > 0
    r = line["roofline"]
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert set(("value", "unit", "cores", "kind", "sample")) <= set(line["cpu_baseline"])
    assert not set(line["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
