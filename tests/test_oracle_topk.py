"""The canonical-fp32 top-k oracle (oracle/topk_ref.c) against an independent float64 brute force."""
import numpy as np

from oracle import topk_ref as T


def _data(n, q, seed, dup=0):
    rng = np.random.default_rng(seed)
    d = rng.normal(size=(n, 16)).astype(np.float32)
    if dup:
        src = rng.integers(0, n, size=dup)
        dst = rng.integers(0, n, size=dup)
        d[dst] = d[src]
    qs = d[rng.integers(0, n, size=q)] + 0.05 * rng.normal(size=(q, 16)).astype(np.float32)
    return d, qs.astype(np.float32)


def test_normalize_rows():
    d, _ = _data(1000, 1, 0)
    d[5] = 0.0
    dn = T.normalize_rows(d)
    np.testing.assert_array_equal(dn[5], 0.0)  # zero norm -> divided by 1 (faiss_db.py:111-112)
    ref = d.astype(np.float64) / np.maximum(np.linalg.norm(d.astype(np.float64), axis=1, keepdims=True), 1e-300)
    ref[5] = 0
    np.testing.assert_allclose(dn, ref, rtol=3e-7, atol=0)


def test_normalize_rows_equals_reference_l2_normalize(golden_dir):
    """The normalisation leg of the search oracle is pinned bit for bit to the reference's own
    FaissLatentVectorDatabase._l2_normalize (latice/index/faiss_db.py:109-113, fixture from oracle/make_golden.py)."""
    import os
    g = np.load(os.path.join(golden_dir, "l2_normalize.npz"))
    np.testing.assert_array_equal(T.normalize_rows(g["rows"]), g["normalized"])


def test_topk_sets_agree_with_sklearn_brute_force_cosine():
    """Independent cross-check of the search leg: scikit-learn's exact cosine k-NN returns the
    same neighbour SETS, except where the k-th and (k+1)-th cosine are within fp32 rounding of each other."""
    nn = __import__("pytest").importorskip("sklearn.neighbors")
    d, q = _data(30000, 200, 11, dup=50)
    dn, qn = T.normalize_rows(d), T.normalize_rows(q)
    dots, idx = T.topk(dn, qn, 10)
    sk = nn.NearestNeighbors(n_neighbors=10, metric="cosine", algorithm="brute").fit(dn.astype(np.float64))
    dist, nbrs = sk.kneighbors(qn.astype(np.float64))
    s = qn.astype(np.float64) @ dn.astype(np.float64).T
    n_diff = 0
    for i in range(len(q)):
        a, b = set(idx[i].tolist()), set(nbrs[i].tolist())
        if a != b:
            n_diff += 1
            for r in a ^ b:   # rows only one side lists tie with the k-th best
                assert abs(s[i, r] - np.sort(s[i])[-10]) < 4e-7
    assert n_diff <= 4
    np.testing.assert_allclose(1.0 - dots, dist, atol=4e-7)


def test_topk_agrees_with_float64_bruteforce_up_to_near_ties():
    d, q = _data(20000, 128, 1)
    dn, qn = T.normalize_rows(d), T.normalize_rows(q)
    dots, idx = T.topk(dn, qn, 10)
    s = qn.astype(np.float64) @ dn.astype(np.float64).T
    ref = np.argsort(-s, axis=1, kind="stable")[:, :10]
    assert (np.diff(dots, axis=1) <= 0).all()
    mism = idx != ref
    # where the ordered lists differ the float64 scores must be within fp32 rounding of each other
    for qi, pos in zip(*np.where(mism)):
        assert abs(s[qi, idx[qi, pos]] - s[qi, ref[qi, pos]]) < 4e-7
    assert mism.mean() < 0.01
    np.testing.assert_allclose(dots, np.take_along_axis(s, idx, 1), atol=3e-7)


def test_ties_break_on_lower_global_index_and_index_base():
    d, q = _data(5000, 32, 2, dup=500)
    d[100:110] = d[7]  # ten exact copies of one row
    q[0] = d[7]
    dn, qn = T.normalize_rows(d), T.normalize_rows(q)
    dots, idx = T.topk(dn, qn, 10, index_base=1000)
    assert idx[0, 0] == 1007 and list(idx[0, 1:]) == list(range(1100, 1109))
    assert (dots[0] == dots[0, 0]).all()
    # order inside every list: dot desc, idx asc
    for a in range(idx.shape[0]):
        for b in range(9):
            assert dots[a, b] > dots[a, b + 1] or (dots[a, b] == dots[a, b + 1] and idx[a, b] < idx[a, b + 1])


def test_small_dictionary_and_merge():
    d, q = _data(7, 3, 3)
    dn, qn = T.normalize_rows(d), T.normalize_rows(q)
    dots, idx = T.topk(dn, qn, 10)
    assert (idx[:, 7:] == -1).all() and np.isinf(dots[:, 7:]).all()
    assert sorted(idx[0, :7]) == list(range(7))
    # row-sharded search + merge == one search
    d2, q2 = _data(3001, 17, 4, dup=100)
    dn, qn = T.normalize_rows(d2), T.normalize_rows(q2)
    full = T.topk(dn, qn, 10)
    cuts = [0, 1000, 1001, 2500, 3001]
    parts = [T.topk(dn[a:b], qn, 10, index_base=a) for a, b in zip(cuts[:-1], cuts[1:])]
    md, mi = T.topk_merge(np.stack([p[0] for p in parts]), np.stack([p[1] for p in parts]))
    np.testing.assert_array_equal(mi, full[1])
    np.testing.assert_array_equal(md, full[0])


def test_threads_do_not_change_results():
    d, q = _data(4000, 64, 5)
    dn, qn = T.normalize_rows(d), T.normalize_rows(q)
    a = T.topk(dn, qn, 10, nthreads=1)
    b = T.topk(dn, qn, 10, nthreads=4)
    np.testing.assert_array_equal(a[1], b[1])
    np.testing.assert_array_equal(a[0], b[0])


def faiss_lists_agree(idx, dots, g_idx, g_sims, rows, queries, tol=2e-6):
    """Our lists against the golden FAISS-path lists: scores agree within ``tol`` position by position, and a row named
    by only one of the two lists scores within ``tol`` of the golden k-th (sgemm and the canonical fma chain round the
    last bits differently, so near-ties may swap).  Returns the number of queries whose row lists differ."""
    np.testing.assert_allclose(dots, g_sims, rtol=0, atol=tol)
    differing = 0
    for q in range(len(idx)):
        if np.array_equal(idx[q], g_idx[q]):
            continue
        differing += 1
        kth = float(g_sims[q][-1])
        for row in set(idx[q].tolist()) ^ set(g_idx[q].tolist()):
            score = float(rows[row].astype(np.float64) @ queries[q].astype(np.float64))
            assert abs(score - kth) <= tol, (q, row, score, kth)
    return differing


def test_topk_equals_the_reference_faiss_call_path(golden_dir):
    """tests/golden/faiss_query.npz: the unmodified FaissLatentVectorDatabase.add_vectors / query_similar
    (latice/index/faiss_db.py:161-193, 216-256) over an exact float32 inner-product stand-in for the faiss wheel
    (oracle/make_golden_faiss_query.py).  The oracle gives the same rows, scores within 2e-6, incl. the zero row, the
    zero query, exact duplicates (lower id first) and fewer rows than n_results."""
    import os
    g = np.load(os.path.join(golden_dir, "faiss_query.npz"))
    rows = T.normalize_rows(g["latents"])
    queries = T.normalize_rows(g["queries"].astype(np.float32))
    dots, idx = T.topk(rows, queries, 10)
    swaps = faiss_lists_agree(idx, dots, g["idx"], g["sims"], rows, queries)
    assert swaps <= 2
    dup = idx[94]                                     # the query that is 3 x row 17: rows 17, 2000..2003 tie at 1.0
    assert dup[:5].tolist() == [17, 2000, 2001, 2002, 2003] and g["idx"][94][:5].tolist() == dup[:5].tolist()
    assert idx[95].tolist() == list(range(10))        # zero query: every score is 0, lowest ids win
    d4, i4 = T.topk(rows[:4], queries[:1], 4)
    np.testing.assert_array_equal(i4[0], g["small_idx"])
    np.testing.assert_allclose(d4[0], g["small_sims"], atol=2e-6)
