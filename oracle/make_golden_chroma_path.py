"""Golden vectors for the reference's Chroma path, end to end: tests/golden/chroma_path.npz.

Test infrastructure; runs only in the build container (needs /root/reference).  The UNMODIFIED reference class
``ChromaLatentVectorDatabase`` (latice/index/chroma_db.py) is driven through ``add_vectors`` (144-208: batches, ids
``vec_{j + count}``, the metadata dictionaries), ``query_similar`` (231-259), ``find_best_orientation`` (261-342: the
REAL query feeding the consensus) and ``find_best_orientations_batch`` (377-410).  The one thing that cannot be
installed here, the ``chromadb`` wheel (0.6.3, uv.lock:553-554; its tests mock the collection too,
tests/index/test_chroma_db.py:267-291), is replaced by a stand-in ``Client`` / collection that keeps what ``add``
receives and answers ``query`` as a collection created with ``{"hnsw:space": "cosine"}`` is specified to: the
``n_results`` smallest cosine distances 1 - q.d / (|q||d|) in float32, ascending, in Chroma's result layout
(``{"ids": [[..]], "distances": [[..]], "metadatas": [[..]]}``) -- EXACTLY, where the real index is an approximate
HNSW graph (its recall is reported by bench.py from oracle/hnsw_ref.c).  Equal distances are ordered by insertion
order, the tie rule this repository states.  What the fixture pins is the reference's call-site semantics and its
consensus on real search output -- not the chromadb binary.

    python oracle/make_golden_chroma_path.py        # rewrites tests/golden/chroma_path.npz
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


class ExactCosineCollection:
    def __init__(self, name, metadata):
        assert metadata.get("hnsw:space") == "cosine"
        self.name, self.metadata = name, metadata
        self.ids, self.metas, self.rows = [], [], np.zeros((0, metadata["dimension"]), dtype=np.float32)

    def count(self) -> int:
        return len(self.ids)

    def add(self, embeddings, metadatas, ids) -> None:
        assert isinstance(embeddings, list) and len(embeddings) == len(metadatas) == len(ids)
        self.ids.extend(ids)
        self.metas.extend(metadatas)
        self.rows = np.concatenate([self.rows, np.asarray(embeddings, dtype=np.float32)])   # hnswlib stores float32

    def query(self, query_embeddings, n_results, include=None):
        q = np.asarray(query_embeddings, dtype=np.float32).reshape(1, -1)
        qn = q / max(float(np.linalg.norm(q)), 1e-30)
        norms = np.linalg.norm(self.rows, axis=1, keepdims=True)
        dn = self.rows / np.where(norms == 0, 1.0, norms)
        dist = (np.float32(1.0) - dn @ qn[0]).astype(np.float32)
        order = np.lexsort((np.arange(len(dist)), dist))[: min(n_results, len(dist))]
        out = {"ids": [[self.ids[i] for i in order]]}
        if include is None or "distances" in include:
            out["distances"] = [[float(dist[i]) for i in order]]
        if include is None or "metadatas" in include:
            out["metadatas"] = [[self.metas[i] for i in order]]
        return out


def main() -> None:
    class _Client:
        def __init__(self, *a, **k):
            self.collections = {}

        def get_collection(self, name):
            if name not in self.collections:
                raise ValueError(f"Collection {name} does not exist.")
            return self.collections[name]

        def create_collection(self, name, metadata=None):
            self.collections[name] = ExactCosineCollection(name, metadata or {})
            return self.collections[name]

    errs = types.ModuleType("chromadb.errors")
    errs.InvalidCollectionException = type("InvalidCollectionException", (Exception,), {})
    chroma = types.ModuleType("chromadb")
    chroma.Client, chroma.PersistentClient, chroma.errors = _Client, _Client, errs
    sys.modules["chromadb"], sys.modules["chromadb.errors"] = chroma, errs
    from oracle import refload

    ref = refload.load()
    chroma_db = ref.chroma_db
    rng = np.random.default_rng(77)
    n, k = 2500, 10
    # latents that are a smooth function of the orientation, so that near latents mean near orientations and the
    # consensus has something to agree on: cubic-symmetric quaternion features through a fixed random projection
    from scipy.spatial.transform import Rotation as R

    base = R.random(60, random_state=3)
    eul = []
    for b in base:
        jitter = R.from_rotvec(rng.normal(size=(n // 60 + 1, 3)) * 0.02)
        eul.append((jitter * b).as_euler("zxz", degrees=True))
    orientations = np.concatenate(eul)[:n]
    quat = R.from_euler("zxz", orientations, degrees=True).as_quat()
    quat *= np.sign(quat[:, 3:4] + 1e-12)
    proj = rng.normal(size=(4, 16))
    latents = (quat @ proj + 0.03 * rng.normal(size=(n, 16))).astype(np.float64)
    queries = latents[rng.integers(0, n, 40)] + 0.03 * rng.normal(size=(40, 16))
    queries[30:] = rng.normal(size=(10, 16))      # far from every row: candidates from unrelated clusters, no consensus

    db = chroma_db.ChromaLatentVectorDatabase(chroma_db.LatentVectorDatabaseConfig(persist_directory=None))
    db.add_vectors(latents[:1700], orientations[:1700], batch_size=1000)     # two calls: the id offset (chroma_db.py:160)
    db.add_vectors(latents[1700:], orientations[1700:], batch_size=512)
    assert db.get_count() == n and db.collection.ids[1700] == "vec_1700"
    q_idx = np.zeros((40, k), dtype=np.int64)
    q_dist = np.zeros((40, k))
    meta_str = []
    for i in range(40):
        res = db.query_similar(queries[i], n_results=k)
        q_idx[i] = [int(s.split("_")[1]) for s in res["ids"][0]]
        q_dist[i] = res["distances"][0]
        meta_str.append([m["orientation_str"] for m in res["metadatas"][0]])
        assert all(m["phi1"] == orientations[j][0] for m, j in zip(res["metadatas"][0], q_idx[i]))
    params = dict(top_n=k, orientation_threshold=0.03, min_required_matches=6, max_iterations=3)
    results = db.find_best_orientations_batch(queries, **params)
    success = np.array([r.success for r in results])
    mean = np.array([r.mean_orientation if r.success else [np.nan] * 3 for r in results])
    best = np.array([r.best_orientation for r in results])
    similar = np.zeros((40, k), dtype=bool)
    for i, r in enumerate(results):
        similar[i, r.similar_indices] = True
    one = db.find_best_orientation(queries[0], **params)
    assert one.success == results[0].success
    np.savez_compressed(os.path.join(GOLDEN, "chroma_path.npz"), latents=latents, orientations=orientations, queries=queries,
                        idx=q_idx, dist=q_dist, orientation_str=np.array(meta_str), success=success, mean=mean, best=best,
                        similar=similar, params=np.array([params["orientation_threshold"], params["min_required_matches"],
                                                          params["max_iterations"]]))
    print("chroma_path.npz: success", int(success.sum()), "of", len(success), "first ids", q_idx[0].tolist())


if __name__ == "__main__":
    main()
