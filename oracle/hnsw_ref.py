"""ctypes wrapper around oracle/hnsw_ref.c: the reference's APPROXIMATE Chroma/HNSW search restated on the CPU.

Test / bench infrastructure only (parity unpinned -- see the header of hnsw_ref.c).  Used by bench.py to report the
recall of the reference's default index (latice/index/chroma_db.py:124-130, 254-258) next to the exact search.
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import topk_ref

# chromadb 0.6.3 HnswParams defaults (hnsw:M, hnsw:construction_ef, hnsw:search_ef); hnswlib's own seed
CHROMA_M, CHROMA_EF_CONSTRUCTION, CHROMA_EF_SEARCH, HNSWLIB_SEED = 16, 100, 10, 100

_configured = False


def _lib():
    global _configured
    lib = topk_ref._load()
    if not _configured:
        fp = ctypes.POINTER(ctypes.c_float)
        ip = ctypes.POINTER(ctypes.c_int64)
        lib.ebsd_oracle_hnsw_build.argtypes = [fp, ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_uint]
        lib.ebsd_oracle_hnsw_build.restype = ctypes.c_void_p
        lib.ebsd_oracle_hnsw_free.argtypes = [ctypes.c_void_p]
        lib.ebsd_oracle_hnsw_free.restype = None
        lib.ebsd_oracle_hnsw_search.argtypes = [ctypes.c_void_p, fp, ctypes.c_int64, ctypes.c_int, ctypes.c_int, fp, ip,
                                                ctypes.c_int]
        lib.ebsd_oracle_hnsw_search.restype = None
        lib.ebsd_oracle_hnsw_stats.argtypes = [ctypes.c_void_p, ip]
        lib.ebsd_oracle_hnsw_stats.restype = None
        _configured = True
    return lib


class HnswIndex:
    """hnswlib-style cosine index over L2-normalised rows [N, d] (float32); single-threaded insertion in row order."""

    def __init__(self, rows_hat: np.ndarray, M: int = CHROMA_M, ef_construction: int = CHROMA_EF_CONSTRUCTION,
                 seed: int = HNSWLIB_SEED) -> None:
        self._rows = np.ascontiguousarray(rows_hat, dtype=np.float32)   # the C side keeps a pointer into this array
        if self._rows.ndim != 2 or self._rows.shape[0] == 0:
            raise ValueError("rows_hat must be a non-empty [N, d] array")
        self._lib = _lib()
        self._h = self._lib.ebsd_oracle_hnsw_build(topk_ref._fp(self._rows), self._rows.shape[0], self._rows.shape[1], M,
                                                   ef_construction, seed)
        if not self._h:
            raise ValueError("hnsw build refused its arguments")
        self.M = M

    def search(self, queries_hat: np.ndarray, k: int, ef: int = CHROMA_EF_SEARCH, nthreads: int = 1):
        """(distance [Q,k] f32 = 1 - dot, ascending; row [Q,k] i64, -1 where fewer than k rows were reached)."""
        q = np.ascontiguousarray(queries_hat, dtype=np.float32)
        if q.ndim != 2 or q.shape[1] != self._rows.shape[1]:
            raise ValueError("queries must be [Q, d]")
        dist = np.empty((q.shape[0], k), dtype=np.float32)
        idx = np.empty((q.shape[0], k), dtype=np.int64)
        self._lib.ebsd_oracle_hnsw_search(self._h, topk_ref._fp(q), q.shape[0], k, ef, topk_ref._fp(dist), topk_ref._ip(idx),
                                          nthreads)
        return dist, idx

    def stats(self) -> dict:
        out = np.zeros(7, dtype=np.int64)
        self._lib.ebsd_oracle_hnsw_stats(self._h, topk_ref._ip(out))
        keys = ("max_level", "entry_point", "max_degree_level0", "max_degree_upper", "self_links", "bad_links", "rows_above_level0")
        return dict(zip(keys, (int(v) for v in out)))

    def close(self) -> None:
        if self._h:
            self._lib.ebsd_oracle_hnsw_free(self._h)
            self._h = None

    def __del__(self):  # noqa: D105
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass


def recall_at_k(approx_idx: np.ndarray, exact_idx: np.ndarray) -> float:
    """Mean fraction of the exact top-k rows present in the approximate lists."""
    hit = (approx_idx[:, :, None] == exact_idx[:, None, :]).any(axis=1)
    return float(hit.mean())
