"""ctypes wrapper around oracle/topk_ref.c (test infrastructure; see that file for the contract)."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libebsd_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    srcs = [os.path.join(_HERE, f) for f in ("topk_ref.c", "hnsw_ref.c")]
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < max(os.path.getmtime(s) for s in srcs):
        subprocess.run(["make", "-C", _HERE, "-B", "_build/libebsd_oracle.so"], check=True, capture_output=True)
    return _LIB_PATH


def _load():
    global _lib
    if _lib is None:
        build()
        lib = ctypes.CDLL(_LIB_PATH)
        fp = ctypes.POINTER(ctypes.c_float)
        ip = ctypes.POINTER(ctypes.c_int64)
        lib.ebsd_oracle_normalize_rows.argtypes = [fp, ctypes.c_int64, ctypes.c_int]
        lib.ebsd_oracle_normalize_rows.restype = None
        lib.ebsd_oracle_topk.argtypes = [fp, ctypes.c_int64, ctypes.c_int64, fp, ctypes.c_int64, ctypes.c_int,
                                         ctypes.c_int, fp, ip, ctypes.c_int]
        lib.ebsd_oracle_topk.restype = None
        lib.ebsd_oracle_topk_merge.argtypes = [fp, ip, ctypes.c_int, ctypes.c_int64, ctypes.c_int, fp, ip]
        lib.ebsd_oracle_topk_merge.restype = None
        _lib = lib
    return _lib


def _fp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def _ip(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_int64))


def normalize_rows(x: np.ndarray) -> np.ndarray:
    """Return a normalised float32 copy of x [n, d] (canonical arithmetic)."""
    out = np.ascontiguousarray(x, dtype=np.float32).copy()
    if out.ndim != 2:
        raise ValueError("x must be [n, d]")
    _load().ebsd_oracle_normalize_rows(_fp(out), out.shape[0], out.shape[1])
    return out


def topk(dict_hat: np.ndarray, queries_hat: np.ndarray, k: int, index_base: int = 0, nthreads: int = 0):
    """Exact top-k of normalised queries against normalised rows. Returns (dot [Q,k] f32, idx [Q,k] i64)."""
    d = np.ascontiguousarray(dict_hat, dtype=np.float32)
    q = np.ascontiguousarray(queries_hat, dtype=np.float32)
    if d.ndim != 2 or q.ndim != 2 or d.shape[1] != q.shape[1]:
        raise ValueError("shape mismatch")
    out_dot = np.empty((q.shape[0], k), dtype=np.float32)
    out_idx = np.empty((q.shape[0], k), dtype=np.int64)
    _load().ebsd_oracle_topk(_fp(d), d.shape[0], index_base, _fp(q), q.shape[0], d.shape[1], k, _fp(out_dot),
                             _ip(out_idx), nthreads)
    return out_dot, out_idx


def topk_merge(dots: np.ndarray, idx: np.ndarray):
    """Merge partial lists [R,Q,k] -> ([Q,k], [Q,k])."""
    dots = np.ascontiguousarray(dots, dtype=np.float32)
    idx = np.ascontiguousarray(idx, dtype=np.int64)
    r, q, k = dots.shape
    out_dot = np.empty((q, k), dtype=np.float32)
    out_idx = np.empty((q, k), dtype=np.int64)
    _load().ebsd_oracle_topk_merge(_fp(dots), _ip(idx), r, q, k, _fp(out_dot), _ip(out_idx))
    return out_dot, out_idx


def pack_candidates(dots: np.ndarray, idx: np.ndarray) -> np.ndarray:
    """(dot f32, global row i64, -1 = empty) -> one int64 word per candidate: (float bits << 32) | (row + 1), the
    exchange format of ebsd_topk_pack / ebsd_topk_merge_packed (include/ebsd_b200.h)."""
    bits = np.ascontiguousarray(dots, dtype=np.float32).view(np.uint32).astype(np.uint64)
    row1 = np.where(idx < 0, 0, idx + 1).astype(np.uint64)
    return ((bits << np.uint64(32)) | row1).view(np.int64)


def unpack_candidates(packed: np.ndarray):
    w = np.ascontiguousarray(packed).view(np.uint64)
    dots = (w >> np.uint64(32)).astype(np.uint32).view(np.float32)
    idx = (w & np.uint64(0xFFFFFFFF)).astype(np.int64) - 1
    return dots, idx
