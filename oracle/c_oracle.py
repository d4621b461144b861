"""ctypes loader for the C oracle (oracle/oracle_nn.c).  TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "oracle_nn.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _SO


def _load():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = ctypes.CDLL(_SO)
        i64, p = ctypes.c_int64, ctypes.c_void_p
        _lib.oracle_nn_f64.argtypes = [p, i64, p, i64, p, p]
        _lib.oracle_nn_f64.restype = None
        _lib.oracle_nn_f32_fma.argtypes = [p, i64, p, i64, p, p]
        _lib.oracle_nn_f32_fma.restype = None
    return _lib


def _threaded(fn, q, t, d2, idx):
    """Split the queries over host threads (ctypes drops the GIL during the call)."""
    n = len(q)
    nthr = max(1, min(len(os.sched_getaffinity(0)), (n + 255) // 256))
    bounds = np.linspace(0, n, nthr + 1).astype(np.int64)

    def run(k):
        a, b = int(bounds[k]), int(bounds[k + 1])
        if b > a:
            fn(q[a:b].ctypes.data, b - a, t.ctypes.data, len(t), d2[a:b].ctypes.data,
               idx[a:b].ctypes.data)

    if nthr == 1:
        run(0)
    else:
        with ThreadPoolExecutor(nthr) as ex:
            list(ex.map(run, range(nthr)))


def nn_f64(q, t):
    """Exact float64 brute-force 1-NN, lowest index on ties -> (d2 f64, idx i64)."""
    q = np.ascontiguousarray(q, dtype=np.float64)
    t = np.ascontiguousarray(t, dtype=np.float64)
    d2 = np.empty(len(q), dtype=np.float64)
    idx = np.empty(len(q), dtype=np.int64)
    _threaded(_load().oracle_nn_f64, q, t, d2, idx)
    return d2, idx


def nn_f32_fma(q, t):
    """float32 direct-difference 1-NN with the CUDA kernel's rounding sequence
    -> (d2 f32, idx i32); the GPU result must match bit for bit."""
    q = np.ascontiguousarray(q, dtype=np.float32)
    t = np.ascontiguousarray(t, dtype=np.float32)
    d2 = np.empty(len(q), dtype=np.float32)
    idx = np.empty(len(q), dtype=np.int32)
    _threaded(_load().oracle_nn_f32_fma, q, t, d2, idx)
    return d2, idx
