"""Golden vectors for the reference's FAISS search path: tests/golden/faiss_query.npz.

Test infrastructure; runs only in the build container (needs /root/reference).  The UNMODIFIED reference class
``FaissLatentVectorDatabase`` (latice/index/faiss_db.py: add_vectors 161-193, query_similar 216-256) is driven end to
end -- its float32 cast, ``_l2_normalize`` of rows and query, the clamp of ``n_results`` to the row count, the
empty-index guard and the ``distances[0], indices[0]`` return -- with the one thing that cannot be installed here, the
``faiss`` wheel (faiss-cpu 1.10.0, uv.lock:813-814), replaced by a stand-in for the two calls the class makes:
``index_factory(d, "Flat", METRIC_INNER_PRODUCT)`` -> an object with ``add`` / ``search`` / ``ntotal`` that restates
IndexFlatIP as published: the inner products of the query with every stored row in float32 (numpy / BLAS sgemm, the
routine faiss itself calls for blocks of queries) and the k largest in descending order.  FAISS does not define the
order of equal scores; the stand-in breaks them on the lower row id, the rule this repository states.
What the fixture pins is therefore the reference's CALL-SITE semantics plus an exact float32 inner-product ranking --
not the faiss binary (which stays "parity unpinned").

    python oracle/make_golden_faiss_query.py        # rewrites tests/golden/faiss_query.npz
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


class FlatIP:
    """IndexFlatIP as published: exact inner products, k largest, descending (ties: lower id)."""

    def __init__(self, d: int) -> None:
        self.d = d
        self.rows = np.zeros((0, d), dtype=np.float32)

    @property
    def ntotal(self) -> int:
        return len(self.rows)

    def add(self, x) -> None:
        assert x.dtype == np.float32 and x.shape[1] == self.d
        self.rows = np.concatenate([self.rows, x])

    def search(self, q, k: int):
        assert q.dtype == np.float32
        sims = q @ self.rows.T
        order = np.lexsort((np.broadcast_to(np.arange(self.ntotal), sims.shape), -sims), axis=1)[:, :k]
        return np.take_along_axis(sims, order, axis=1), order.astype(np.int64)


def main() -> None:
    from oracle import refload

    refload.load()
    faiss = types.ModuleType("faiss")
    faiss.METRIC_INNER_PRODUCT = 0
    faiss.index_factory = lambda d, desc, metric: FlatIP(d)
    sys.modules["faiss"] = faiss
    import latice.index.faiss_db as faiss_db

    rng = np.random.default_rng(2025)
    n, q, k = 3000, 96, 10
    latents = (rng.normal(size=(n, 16)) * rng.uniform(0.2, 5.0, size=(n, 1))).astype(np.float32)   # un-normalised rows
    latents[1500] = 0.0                                      # a zero row: the reference divides it by 1
    latents[2000:2004] = latents[17]                         # exact duplicates: equal scores
    orientations = rng.uniform(0, 1, size=(n, 3)) * np.array([360.0, 180.0, 360.0])
    queries = np.concatenate([
        latents[rng.integers(0, n, 48)] + 0.05 * rng.normal(size=(48, 16)).astype(np.float32),   # near a row
        rng.normal(size=(46, 16)).astype(np.float32),                                             # anywhere
        latents[17:18] * 3.0,                                                                     # hits the duplicates
        np.zeros((1, 16), dtype=np.float32),                                                      # zero query
    ]).astype(np.float64)                                     # the reference casts queries to float32 itself

    cfg = faiss_db.FaissLatentVectorDatabaseConfig(npz_path=os.path.join("/tmp", "ebsd_golden_no_such_index.npz"))
    db = faiss_db.FaissLatentVectorDatabase(cfg)
    assert db.query_similar(queries[0], n_results=k)[0].size == 0           # empty-index guard (faiss_db.py:232-234)
    db.add_vectors(latents, orientations)
    sims = np.zeros((q, k), dtype=np.float32)
    idx = np.zeros((q, k), dtype=np.int64)
    for i in range(q):
        s, j = db.query_similar(queries[i], n_results=k)
        sims[i], idx[i] = s, j
    # fewer rows than n_results: the reference returns all of them (faiss_db.py:235-239)
    small = faiss_db.FaissLatentVectorDatabase(cfg)
    small.add_vectors(latents[:4], orientations[:4])
    s_small, i_small = small.query_similar(queries[0], n_results=k)
    assert len(i_small) == 4
    np.savez_compressed(os.path.join(GOLDEN, "faiss_query.npz"), latents=latents, orientations=orientations,
                        queries=queries, sims=sims, idx=idx, small_sims=s_small.astype(np.float32), small_idx=i_small)
    print("faiss_query.npz:", sims.shape, "first list", idx[0].tolist())


if __name__ == "__main__":
    main()
