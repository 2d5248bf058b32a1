"""oracle/hnsw_ref.c -- the reference's approximate Chroma/HNSW index restated for the recall figure in bench.py.

Parity is UNPINNED (chromadb / chroma-hnswlib are not installable here and the reference's tests mock the collection,
/root/reference/tests/index/test_chroma_db.py:267-291); what can be checked is the restatement against the exact oracle
and the structural invariants of hnswlib's graph (hnswalg.h).
"""
import math

import numpy as np
import pytest

from oracle import hnsw_ref, topk_ref


@pytest.fixture(scope="module")
def small_index():
    rng = np.random.default_rng(5)
    rows = topk_ref.normalize_rows(rng.normal(size=(6000, 16)).astype(np.float32))
    queries = topk_ref.normalize_rows(rows[:300] + 0.05 * rng.normal(size=(300, 16)).astype(np.float32))
    index = hnsw_ref.HnswIndex(rows)
    yield rows, queries, index
    index.close()


def test_graph_invariants(small_index):
    rows, _, index = small_index
    st = index.stats()
    assert st["self_links"] == 0 and st["bad_links"] == 0
    assert st["max_degree_level0"] <= 2 * index.M and st["max_degree_upper"] <= index.M   # maxM0 = 2 M, maxM = M
    # level = floor(-ln(U) / ln(M)): P(level > 0) = 1 / M
    frac = st["rows_above_level0"] / len(rows)
    assert abs(frac - 1.0 / index.M) < 4 * math.sqrt((1 / index.M) * (1 - 1 / index.M) / len(rows))
    assert 1 <= st["max_level"] <= 6


def test_lists_are_true_distances_in_ascending_order(small_index):
    rows, queries, index = small_index
    dist, idx = index.search(queries, 10, ef=hnsw_ref.CHROMA_EF_SEARCH)
    assert (idx >= 0).all() and (np.diff(dist, axis=1) >= 0).all()
    for r in range(idx.shape[0]):
        assert len(set(idx[r])) == 10
    want = 1.0 - np.einsum("qkd,qd->qk", rows[idx].astype(np.float64), queries.astype(np.float64))
    assert np.abs(dist - want).max() < 1e-6          # Chroma's cosine distance 1 - q.d of the rows it names


def test_recall_against_the_exact_oracle_grows_with_ef(small_index):
    rows, queries, index = small_index
    _, exact = topk_ref.topk(rows, queries, 10)
    recalls = []
    for ef in (10, 40, 400, 6000):
        _, idx = index.search(queries, 10, ef=ef, nthreads=2)
        recalls.append(hnsw_ref.recall_at_k(idx, exact))
    assert recalls[0] > 0.8                           # an approximate index: good, not exact, at Chroma's default ef
    assert all(b >= a - 1e-9 for a, b in zip(recalls, recalls[1:]))
    assert recalls[-1] == 1.0                         # a beam as wide as the dictionary is an exhaustive search


def test_search_is_deterministic_and_thread_count_independent(small_index):
    _, queries, index = small_index
    a = index.search(queries, 10, ef=10, nthreads=1)
    b = index.search(queries, 10, ef=10, nthreads=3)
    assert np.array_equal(a[1], b[1]) and np.array_equal(a[0], b[0])


def test_fewer_rows_than_k():
    rows = topk_ref.normalize_rows(np.eye(16, dtype=np.float32)[:4])
    index = hnsw_ref.HnswIndex(rows)
    dist, idx = index.search(rows[:1], 10)
    assert sorted(idx[0][:4]) == [0, 1, 2, 3] and (idx[0][4:] == -1).all() and np.isinf(dist[0][4:]).all()
    assert idx[0][0] == 0 and abs(dist[0][0]) < 1e-6
    with pytest.raises(ValueError):
        index.search(np.zeros((1, 8), dtype=np.float32), 3)
