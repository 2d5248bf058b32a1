"""GPU parity: ebsd_normalize_rows / ebsd_topk / ebsd_topk_merge against the canonical-fp32 C oracle (bit-exact)."""
import numpy as np
import pytest
import torch

from oracle import topk_ref as T

pytestmark = pytest.mark.gpu


def _db(latents, index_base=0):
    import ebsd_vae_b200 as E
    db = E.LatentVectorDatabase()
    db.index_base = index_base
    if len(latents):
        db.add_vectors(latents, np.zeros((len(latents), 3)))
    return db


def _data(n, q, seed, dup=0.0):
    rng = np.random.default_rng(seed)
    d = rng.normal(size=(n, 16)).astype(np.float32)
    nd = int(n * dup)
    if nd:
        d[rng.integers(0, n, size=nd)] = d[rng.integers(0, n, size=nd)]
    qs = d[rng.integers(0, max(n, 1), size=q)] if n else rng.normal(size=(q, 16))
    qs = (qs + 0.05 * rng.normal(size=(q, 16))).astype(np.float32)
    return d, qs


def _search(db, q, k):
    qh = db._prepare_queries(q)
    dot, idx, dist = db.search_device(qh, k)
    torch.cuda.synchronize()
    return qh.cpu().numpy(), dot.cpu().numpy(), idx.cpu().numpy(), dist.cpu().numpy()


def test_normalize_rows_bit_exact():
    d, _ = _data(10000, 1, 0)
    d[17] = 0.0
    d[18] *= 1e-20
    db = _db(d)
    got = db._latents[: db.get_count()].cpu().numpy()
    np.testing.assert_array_equal(got, T.normalize_rows(d))


def test_normalize_rows_equals_reference_l2_normalize(golden_dir):
    """Bit for bit the output of the reference's own FaissLatentVectorDatabase._l2_normalize
    (latice/index/faiss_db.py:109-113; fixture made by oracle/make_golden.py from the unmodified class)."""
    import os
    g = np.load(os.path.join(golden_dir, "l2_normalize.npz"))
    db = _db(g["rows"])
    np.testing.assert_array_equal(db._latents[: db.get_count()].cpu().numpy(), g["normalized"])


@pytest.mark.parametrize("n,q,k,dup", [
    (625, 1, 10, 0.0), (625, 1, 20, 0.0), (1000, 5, 20, 0.01), (5000, 200, 10, 0.01), (127, 3, 10, 0.0),
    (128, 16, 10, 0.0), (129, 17, 32, 0.0), (7, 4, 10, 0.0), (100_000, 300, 10, 0.001), (33_333, 1000, 1, 0.01),
    (20_000, 129, 10, 0.3),
    # K2q, the streaming kernel for a handful of queries over >= 32 768 rows (one lane per row, TQ = 1, 2, 4, 8)
    (32_768, 1, 10, 0.0), (200_013, 1, 10, 0.2), (150_001, 2, 32, 0.3), (99_999, 3, 10, 0.3), (400_000, 5, 1, 0.1),
    (65_537, 8, 20, 0.5),
])
def test_topk_bit_exact_vs_oracle(n, q, k, dup):
    d, qs = _data(n, q, seed=n + q + k, dup=dup)
    db = _db(d, index_base=1000)
    qh, dot, idx, dist = _search(db, qs, k)
    dn = db._latents[:n].cpu().numpy()
    np.testing.assert_array_equal(qh, T.normalize_rows(qs))
    odot, oidx = T.topk(dn, qh, k, index_base=1000, nthreads=8)
    np.testing.assert_array_equal(idx, oidx)
    np.testing.assert_array_equal(dot, odot)
    filled = oidx >= 0
    np.testing.assert_array_equal(dist[filled], (np.float32(1.0) - odot)[filled])


def test_exact_duplicates_tie_break_on_index():
    d, qs = _data(4096, 8, 3)
    d[1000:1012] = d[5]
    qs[0] = d[5]
    db = _db(d)
    _, dot, idx, _ = _search(db, qs, 10)
    assert idx[0, 0] == 5 and list(idx[0, 1:]) == list(range(1000, 1009))


def test_empty_dictionary_and_empty_queries():
    db = _db(np.zeros((0, 16), np.float32))
    _, dot, idx, _ = _search(db, np.ones((3, 16), np.float32), 10)
    assert (idx == -1).all() and np.isneginf(dot).all()
    d, _ = _data(300, 1, 1)
    db = _db(d)
    _, dot, idx, _ = _search(db, np.zeros((0, 16), np.float32), 10)
    assert idx.shape == (0, 10)


def test_merge_kernel_matches_oracle_and_sharded_equals_whole():
    from ebsd_vae_b200 import _native
    d, qs = _data(30_011, 257, 9, dup=0.01)
    whole = _db(d)
    qh, dot, idx, _ = _search(whole, qs, 10)
    cuts = [0, 10_000, 10_001, 25_000, 30_011]
    dn = whole._latents[: len(d)]
    dots, idxs = [], []
    for a, b in zip(cuts[:-1], cuts[1:]):
        shard = _db(np.zeros((0, 16), np.float32), index_base=a)
        shard._latents, shard._count, shard._capacity = dn[a:b].clone(), b - a, b - a
        qd = torch.from_numpy(qh).cuda()
        sd, si, _ = shard.search_device(qd, 10)
        dots.append(sd)
        idxs.append(si)
    dots_t, idx_t = torch.stack(dots).contiguous(), torch.stack(idxs).contiguous()
    out_d = torch.empty_like(dots[0])
    out_i = torch.empty_like(idxs[0])
    out_dist = torch.empty_like(dots[0])
    lib = _native.load()
    _native.check(lib.ebsd_topk_merge(dots_t.data_ptr(), idx_t.data_ptr(), 4, 257, 10, out_d.data_ptr(),
                                      out_i.data_ptr(), out_dist.data_ptr(),
                                      torch.cuda.current_stream().cuda_stream), "merge")
    torch.cuda.synchronize()
    np.testing.assert_array_equal(out_i.cpu().numpy(), idx)
    np.testing.assert_array_equal(out_d.cpu().numpy(), dot)
    od, oi = T.topk_merge(dots_t.cpu().numpy(), idx_t.cpu().numpy())
    np.testing.assert_array_equal(oi, idx)
    np.testing.assert_array_equal(od, dot)


def test_full_size_properties_1m_rows():
    """BASELINE config 3 size (N = 1M): sortedness, split-invariance, and a sampled bit-exact check."""
    g = torch.Generator(device="cuda").manual_seed(2024)
    d = torch.randn((1_000_000, 16), generator=g, device="cuda")
    d[::1000] = d[1::1000]  # 0.1 % exact duplicates
    import ebsd_vae_b200 as E
    db = E.LatentVectorDatabase()
    db.add_vectors(d, torch.zeros((len(d), 3), dtype=torch.float64, device="cuda"))
    q = d[torch.randint(0, len(d), (4096,), generator=g, device="cuda")] + 0.05 * torch.randn(
        (4096, 16), generator=g, device="cuda")
    qh = db._prepare_queries(q)
    dot, idx, _ = db.search_device(qh, 10)
    assert bool((dot[:, 1:] <= dot[:, :-1]).all())
    ties = dot[:, 1:] == dot[:, :-1]
    assert bool((idx[:, 1:][ties] > idx[:, :-1][ties]).all())
    assert int(ties.sum()) > 0
    # a different query batch size takes a different (query tile, split) plan: results must not change
    dot2, idx2, _ = db.search_device(qh[:100].contiguous(), 10)
    assert torch.equal(idx2, idx[:100]) and torch.equal(dot2, dot[:100])
    dot3, idx3, _ = db.search_device(qh[:7].contiguous(), 10)
    assert torch.equal(idx3, idx[:7]) and torch.equal(dot3, dot[:7])
    # sampled oracle check
    dn = db._latents[: len(d)].cpu().numpy()
    odot, oidx = T.topk(dn, qh[:48].cpu().numpy(), 10, nthreads=8)
    np.testing.assert_array_equal(idx[:48].cpu().numpy(), oidx)
    np.testing.assert_array_equal(dot[:48].cpu().numpy(), odot)


# the three seeding paths of the screen (csrc/topk.cu run_screen / screen_stages): 420 000 and 301 000 rows seed with the
# tiled exact kernel (832 / 608 rows), 1 200 000 rows with the 16-rows-per-lane selection kernel (320 rows), 600 000 with
# the 8-rows-per-lane one (256 rows)
@pytest.mark.parametrize("n,q", [(420_000, 2500), (301_000, 2100), (1_200_000, 2100), (600_000, 2300)])
def test_tensor_core_screen_equals_exact_kernel_incl_overflow_fallback(n, q):
    """for batched searches (N >= 65 536, Q >= 2048) ebsd_topk screens with tensor cores and re-ranks the survivors with the
    canonical arithmetic.  Its lists must equal the CUDA-core kernel's bit for bit -- also for queries whose k-th best
    dot is shared by hundreds of duplicate rows (survivor buffers overflow -> exact scan of the range) -- and the
    sampled oracle check pins both to the restatement."""
    g = torch.Generator(device="cuda").manual_seed(77)
    d = torch.randn((n, 16), generator=g, device="cuda")
    d[100_000:100_700] = d[5]                     # 700 exact copies of row 5
    d[300_000:300_040] = d[6] * 3.0               # 40 rescaled copies of row 6 (identical after normalisation)
    import ebsd_vae_b200 as E
    db = E.LatentVectorDatabase()
    db.add_vectors(d, torch.zeros((n, 3), dtype=torch.float64, device="cuda"))
    qs = d[torch.randint(0, n, (q,), generator=g, device="cuda")] + 0.03 * torch.randn((q, 16), generator=g, device="cuda")
    qs[0] = d[5]
    qs[1] = d[6]
    qh = db._prepare_queries(qs)
    dot_s, idx_s, _ = db.search_device(qh, 10)                       # screen path (N >= 65 536, Q >= 2048)
    dot_e, idx_e, _ = db.search_device(qh[:1000].contiguous(), 10)   # Q < 2048 -> CUDA-core kernel
    assert torch.equal(idx_s[:1000], idx_e) and torch.equal(dot_s[:1000], dot_e)
    assert idx_s[0].tolist() == [5] + list(range(100_000, 100_009))  # ties: lowest rows win
    assert idx_s[1].tolist()[:1] == [6] and set(idx_s[1].tolist()[1:]) <= set(range(300_000, 300_040))
    dn = db._latents[:n].cpu().numpy()
    odot, oidx = T.topk(dn, qh[:24].cpu().numpy(), 10, nthreads=8)
    np.testing.assert_array_equal(idx_s[:24].cpu().numpy(), oidx)
    np.testing.assert_array_equal(dot_s[:24].cpu().numpy(), odot)


@pytest.mark.parametrize("k", [1, 32])
def test_tensor_core_screen_other_k(k):
    """The screen path at the ends of the supported k range (EBSD_MAX_TOPK = 32 rides on CAP - 32 = 64 buffer entries),
    with 30 % duplicated rows so that ties at the k-th place are common: equal to the CUDA-core kernel and the oracle."""
    d, qs = _data(70_001, 2100, seed=5 + k, dup=0.3)
    db = _db(d, index_base=7)
    qh = db._prepare_queries(qs)
    dot_s, idx_s, _ = db.search_device(qh, k)                           # screen (N >= 65 536, Q >= 2048)
    dot_e, idx_e, _ = db.search_device(qh[:900].contiguous(), k)        # CUDA-core kernel
    torch.cuda.synchronize()
    assert torch.equal(idx_s[:900], idx_e) and torch.equal(dot_s[:900], dot_e)
    dn = db._latents[:70_001].cpu().numpy()
    odot, oidx = T.topk(dn, qh[:32].cpu().numpy(), k, index_base=7, nthreads=8)
    np.testing.assert_array_equal(idx_s[:32].cpu().numpy(), oidx)
    np.testing.assert_array_equal(dot_s[:32].cpu().numpy(), odot)


def test_screen_workspace_covers_every_query_chunk():
    """A long query batch runs through the screen in chunks of 131 072 queries and the short LAST chunk has its own
    plan (it splits the seeding search more ways): N = 4.2 M rows, Q = 131 072 + 118 928 (a 500 x 500 map) once wrote
    past a workspace sized for the first chunk only.  The guard bytes behind the workspace must stay untouched and the
    lists of the last chunk must equal a separate search of those queries."""
    import ebsd_vae_b200 as E
    from ebsd_vae_b200 import _native
    lib = _native.load()
    n, q, k = 4_200_000, 131_072 + 118_928, 20
    g = torch.Generator(device="cuda").manual_seed(3)
    d = torch.randn((n, 16), generator=g, device="cuda")
    db = E.LatentVectorDatabase()
    db.add_vectors(d, torch.zeros((n, 3), dtype=torch.float64, device="cuda"))
    qs = d[torch.randint(0, n, (q,), generator=g, device="cuda")] + 0.03 * torch.randn((q, 16), generator=g, device="cuda")
    qh = db._prepare_queries(qs)
    need = int(lib.ebsd_topk_workspace_bytes(n, q, k))
    guard = 64 << 20
    ws = torch.zeros(need + guard, dtype=torch.uint8, device="cuda")
    ws[need:] = 0xA5
    dot = torch.empty((q, k), dtype=torch.float32, device="cuda")
    idx = torch.empty((q, k), dtype=torch.int64, device="cuda")
    _native.check(lib.ebsd_topk(db._latents.data_ptr(), n, 0, qh.data_ptr(), q, k, dot.data_ptr(), idx.data_ptr(), None,
                                ws.data_ptr(), need, torch.cuda.current_stream().cuda_stream), "ebsd_topk")
    torch.cuda.synchronize()
    assert bool((ws[need:] == 0xA5).all()), "ebsd_topk wrote past the workspace it asked for"
    # one byte less must be refused, not overrun
    assert lib.ebsd_topk(db._latents.data_ptr(), n, 0, qh.data_ptr(), q, k, dot.data_ptr(), idx.data_ptr(), None,
                         ws.data_ptr(), need - 1, torch.cuda.current_stream().cuda_stream) == -4
    tail = qh[131_072:131_072 + 1500].contiguous()                     # Q < 2048 -> CUDA-core kernel
    dot_e, idx_e, _ = db.search_device(tail, k)
    assert torch.equal(idx[131_072:131_072 + 1500], idx_e) and torch.equal(dot[131_072:131_072 + 1500], dot_e)


def test_search_equals_the_reference_faiss_call_path(golden_dir):
    """tests/golden/faiss_query.npz = the unmodified FaissLatentVectorDatabase.add_vectors / query_similar
    (latice/index/faiss_db.py:161-193, 216-256) over an exact float32 inner-product stand-in for the faiss wheel:
    the GPU dictionary returns the same rows in the same order (duplicates: lower id first), inner products within
    2e-6, for un-normalised rows, a zero row, a zero query, and fewer rows than n_results."""
    import os

    import ebsd_vae_b200 as E

    g = np.load(os.path.join(golden_dir, "faiss_query.npz"))
    db = E.LatentVectorDatabase(E.LatentVectorDatabaseConfig(persist_directory=None, mode="faiss"))
    db.add_vectors(g["latents"], g["orientations"])
    _, dot, idx, dist = _search(db, g["queries"].astype(np.float32), 10)
    np.testing.assert_array_equal(idx, g["idx"])
    np.testing.assert_allclose(dot, g["sims"], rtol=0, atol=2e-6)
    # the batch surface in FAISS semantics carries the inner products as `distances` (faiss_db.py:241-256, 281-300)
    res = db.find_best_orientations_batch(g["queries"], top_n=10, orientation_threshold=3.0, min_required_matches=3)
    np.testing.assert_array_equal(res.indices, g["idx"])
    np.testing.assert_allclose(res.distances, g["sims"], rtol=0, atol=2e-6)
    np.testing.assert_array_equal(res.candidate_orientations[5], g["orientations"][g["idx"][5]])
    # single-query call, Chroma-shaped: ids name the same rows
    one = db.query_similar(g["queries"][3], n_results=10)
    assert one["ids"][0] == [f"vec_{i}" for i in g["idx"][3]]
    small = E.LatentVectorDatabase(E.LatentVectorDatabaseConfig(persist_directory=None, mode="faiss"))
    small.add_vectors(g["latents"][:4], g["orientations"][:4])
    got = small.query_similar(g["queries"][0], n_results=10)       # fewer rows than n_results: all of them
    assert got["ids"][0] == [f"vec_{i}" for i in g["small_idx"]]


def test_faiss_twin_class_equals_the_reference_faiss_call_path(golden_dir, tmp_path):
    """The drop-in FaissLatentVectorDatabase (reference: latice/index/faiss_db.py:92-496) against
    tests/golden/faiss_query.npz: query_similar returns the reference's (similarities, indices) tuple, clamps n_results
    to the row count, and the single .npz is reopened by a new object."""
    import os

    import ebsd_vae_b200 as E

    g = np.load(os.path.join(golden_dir, "faiss_query.npz"))
    cfg = E.FaissLatentVectorDatabaseConfig(npz_path=str(tmp_path / "faiss_index.npz"))
    db = E.FaissLatentVectorDatabase(cfg)
    db.add_vectors(g["latents"], g["orientations"])
    for i in (0, 50, 94, 95):
        sims, idx = db.query_similar(g["queries"][i], n_results=10)
        assert sims.dtype == np.float32 and idx.dtype == np.int64
        np.testing.assert_array_equal(idx, g["idx"][i])
        np.testing.assert_allclose(sims, g["sims"][i], rtol=0, atol=2e-6)
    with pytest.raises(ValueError, match="Expected query vector of dimension 16, got 8"):
        db.query_similar(np.zeros(8))
    res = db.find_best_orientation(g["queries"][3], top_n=10, orientation_threshold=30.0, min_required_matches=2)
    np.testing.assert_array_equal(res.candidate_orientations, g["orientations"][g["idx"][3]])
    np.testing.assert_allclose(res.distances, g["sims"][3], rtol=0, atol=2e-6)     # FAISS carries inner products
    assert res.query_vector.dtype == np.float64
    if res.success:
        np.testing.assert_array_equal(res.best_orientation, res.mean_orientation)  # faiss_db.py:338-342
    db.save()
    again = E.FaissLatentVectorDatabase(cfg)                                       # reopens the file (faiss_db.py:129-130)
    assert again.get_count() == len(g["latents"])
    np.testing.assert_array_equal(again.query_similar(g["queries"][7], 10)[1], g["idx"][7])
    small = E.FaissLatentVectorDatabase(E.FaissLatentVectorDatabaseConfig(npz_path=str(tmp_path / "small.npz")))
    small.add_vectors(g["latents"][:4], g["orientations"][:4])
    sims, idx = small.query_similar(g["queries"][0], n_results=10)
    np.testing.assert_array_equal(idx, g["small_idx"])
    np.testing.assert_allclose(sims, g["small_sims"], atol=2e-6)
    r4 = small.find_best_orientation(g["queries"][0], top_n=10)                    # top_n clamps to the 4 rows
    assert r4.candidate_orientations.shape == (4, 3)
    again.delete_persistence()
    assert again.get_count() == 0 and not (tmp_path / "faiss_index.npz").exists()
