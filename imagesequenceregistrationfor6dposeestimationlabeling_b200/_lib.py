"""ctypes binding of csrc/libisr.so (the C ABI declared in include/isr.h).

There is no CPU fallback: if the library is missing or a call fails, an exception is
raised.  PyTorch only supplies device memory and the CUDA stream; every data pointer
handed to the library is a ``tensor.data_ptr()``.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
#: (ISR_LIBISR_PATH: developer A/B runs against another build of the same ABI)
LIB_PATH = os.environ.get("ISR_LIBISR_PATH") or os.path.join(_HERE, "csrc", "libisr.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "isr.h")

ISR_SOA_TILE = 1024
ISR_SUB_TILE = 64
ISR_PAD_COORD = np.float32(1.0e18)
ISR_ICP_NSUMS = 17
ISR_PEER_MAX_RANKS = 8
ISR_PEER_MAX_STARTS = 64
ISR_PEER_HANDLE_BYTES = 64

#: numpy mirror of ``struct IsrIcpState`` (include/isr.h)
ICP_STATE_DTYPE = np.dtype([
    ("T", np.float64, (16,)),
    ("fitness", np.float64),
    ("inlier_rmse", np.float64),
    ("prev_fitness", np.float64),
    ("prev_rmse", np.float64),
    ("n_corr", np.int64),
    ("iters", np.int32),
    ("evals", np.int32),
    ("done", np.int32),
    ("reserved", np.int32),
])
assert ICP_STATE_DTYPE.itemsize == 184


class IsrCloud(ctypes.Structure):
    """ctypes mirror of ``struct IsrCloud`` (include/isr.h)."""
    _fields_ = [("soa7", ctypes.c_void_p), ("n", ctypes.c_int64), ("npad", ctypes.c_int64),
                ("bstride", ctypes.c_int64), ("stage_c", ctypes.c_void_p), ("perm", ctypes.c_void_p),
                ("sub_c", ctypes.c_void_p), ("hint", ctypes.c_void_p), ("sub_box", ctypes.c_void_p)]


class IsrError(RuntimeError):
    """A libisr call returned a negative status."""

    def __init__(self, status: int, message: str):
        super().__init__(f"libisr error {status}: {message}")
        self.status = status


_P = ctypes.c_void_p
_I64 = ctypes.c_int64
_I = ctypes.c_int
_D = ctypes.c_double
_SZ = ctypes.c_size_t

# name -> (restype, argtypes); must list every function declared in include/isr.h
SIGNATURES = {
    "isr_version": (_I, []),
    "isr_last_error": (ctypes.c_char_p, []),
    "isr_device_info": (_I, [_P, _P, _P]),
    "isr_launch_count": (ctypes.c_uint64, []),
    "isr_reset_launch_count": (None, []),
    "isr_profile_enable": (_I, [_I]),
    "isr_profile_collect": (_I, [_P, _P]),
    "isr_soa_padded_len": (_I64, [_I64]),
    "isr_transform_points": (_I, [_P, _I64, _P, _I64, _P, _P]),
    "isr_transform_points_f64": (_I, [_P, _I64, _P, _P, _P]),
    "isr_transform_points_soa": (_I, [_P, _I64, _P, _I64, _I64, _P, _I64, _P, _I64, _P]),
    "isr_nn_workspace_bytes": (_SZ, [_I64, _I64, _I64]),
    "isr_nn_soa": (_I, [_P, _I64, _I64, _I64, _P, _I64, _I64, _I64, _I64, _P, _P, _P, _I64, _P,
                        _SZ, _P]),
    "isr_centroid": (_I, [_P, _I64, _P, _P]),
    "isr_spatial_order_workspace_bytes": (_SZ, [_I64]),
    "isr_spatial_order": (_I, [_P, _I64, _P, _P, _SZ, _P]),
    "isr_prepare_cloud": (_I, [_P, _P, _P, _I64, _P, _I64, _P, _I64, _P, _I64, _P, _I64, _P, _I64, _P]),
    "isr_stage_sphere_count": (_I64, [_I64]),
    "isr_tile_spheres": (_I, [_P, _I64, _I64, _I64, _I64, _P, _P, _P, _P]),
    "isr_set_nn_pruning": (_I, [_I]),
    "isr_get_nn_pruning": (_I, []),
    "isr_profile_nn_pairs": (_I, [_P, _P]),
    "isr_profile_nn_counters": (_I, [_P]),
    "isr_nn2_workspace_bytes": (_SZ, [_I64, _I64, _I64]),
    "isr_nn2": (_I, [_P, _P, _I64, _I, _P, _P, _P, _I64, _P, _SZ, _P]),
    "isr_mean_sqrt": (_I, [_P, _I64, _I64, _P, _P]),
    "isr_verify_workspace_bytes": (_SZ, [_I64, _I64, _I64, _I]),
    "isr_verify_poses": (_I, [_P, _I64, _P, _I64, _P, _P, _P, _I64, _I, _P, _P, _P, _SZ, _P]),
    "isr_rel_pose_table": (_I, [_P, _P, _I64, _I64, _I64, _P, _P]),
    "isr_rigid_relative": (_I, [_P, _P, _I64, _P, _P]),
    "isr_adds_fixed_target_workspace_bytes": (_SZ, [_I64, _I64, _I64]),
    "isr_adds_fixed_target": (_I, [_P, _I64, _P, _I64, _P, _P, _I64, _P, _P, _P, _SZ, _P]),
    "isr_adds_bounds": (_I, [_P, _I64, _P, _I64, _P, _P, _I64, _P, _P, _P, _P]),
    "isr_vote": (_I, [_P, _I64, _I64, _D, _P, _P, _P, _P]),
    "isr_icp_workspace_bytes": (_SZ, [_I64, _I64, _I64]),
    "isr_icp_accumulate": (_I, [_P, _I64, _P, _P, _P, _I64, _P, _P, _P, _D, _P, _P, _P, _P, _SZ, _P]),
    "isr_icp_search": (_I, [_P, _I64, _P, _P, _P, _I64, _P, _P, _P, _P, _SZ, _P]),
    "isr_icp_corr_dist": (_I, [_P, _I64, _P, _P, _I64, _P, _P, _P, _P]),
    "isr_icp_accumulate_corr": (_I, [_P, _I64, _P, _P, _P, _I64, _P, _I64, _P, _D, _P, _P, _P, _SZ, _P]),
    "isr_icp_solve": (_I, [_P, _I64, _P, _I64, _D, _D, _I, _P]),
    "isr_icp_run": (_I, [_P, _I64, _P, _P, _P, _I64, _P, _P, _P, _D, _I, _D, _D, _P, _P, _P, _P, _SZ,
                         _P]),
    "isr_peer_create": (_I, [_I, _I, _P, _P]),
    "isr_peer_connect": (_I, [_P, _P]),
    "isr_peer_set_timeout": (_I, [_P, _D]),
    "isr_peer_destroy": (_I, [_P]),
    "isr_icp_run_sharded": (_I, [_P, _I64, _P, _P, _P, _I64, _I64, _P, _P, _P, _D, _I, _D, _D, _P, _P, _P,
                                 _P, _SZ, _P, _P]),
    "isr_radius_count": (_I, [_P, _P, _D, _P, _P]),
    "isr_knn_normals": (_I, [_P, _I64, _I64, _I, _P, _P]),
    "isr_pnp_score": (_I, [_P, _P, _I64, _P, _P, _I64, _D, _P, _P, _P]),
    "isr_first_max": (_I, [_P, _I64, _P, _P]),
    "isr_p3p_solve": (_I, [_P, _P, _P, _I64, _P, _P, _P]),
    "isr_pnp_ransac_workspace_bytes": (_SZ, [_I64, _I64]),
    "isr_pnp_ransac": (_I, [_P, _P, _I64, _P, _I64, ctypes.c_uint64, _D, _I, _P, _P, _P, _P, _SZ, _P]),
    "isr_bench_ffma": (_I, [_I, _I, _I, _P, _P, _P]),
    "isr_debug_cta_log": (_I, [_P, _I64]),
}

_lib = None


def load() -> ctypes.CDLL:
    """Load libisr.so (once).  Raises if it has not been built -- there is no fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `make -C {os.path.dirname(LIB_PATH)}` "
                "or `python -c 'import __graft_entry__ as g; g.build()'`. "
                "This package has no CPU fallback.")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            if os.environ.get("ISR_LIBISR_PATH") and not hasattr(lib, name):
                continue  # (an older build in a developer A/B run)
            fn = getattr(lib, name)  # AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(status: int) -> None:
    if status != 0:
        msg = load().isr_last_error()
        raise IsrError(status, msg.decode("utf-8", "replace") if msg else "")


def soa_padded_len(n: int) -> int:
    return max(ISR_SOA_TILE, (int(n) + ISR_SOA_TILE - 1) // ISR_SOA_TILE * ISR_SOA_TILE)


def launch_count() -> int:
    return int(load().isr_launch_count())


def reset_launch_count() -> None:
    load().isr_reset_launch_count()
