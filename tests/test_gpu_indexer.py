"""GPU end-to-end: DiffractionPatternIndexer.build_dictionary / index_pattern against the oracle pipeline
(transform -> encoder -> exact top-k -> consensus) on the reference's sample-shaped inputs."""
import os

import numpy as np
import pytest
import torch

from oracle import consensus_ref as C
from oracle import encoder_ref as R
from oracle import topk_ref as T
from oracle import transform_ref as X

pytestmark = pytest.mark.gpu
N_PAT = 625   # the whole sample angle file (625 orientations), as BASELINE configs[0]


def _misorientation_deg(e1, e2):
    q1, q2 = C.quat_from_euler_zxz_deg(e1), C.quat_from_euler_zxz_deg(e2)
    return np.degrees(C.quat_angle(C.quat_mul(q1, C.quat_conj(q2))))


@pytest.fixture(scope="module")
def setup(tmp_path_factory, golden_dir):
    import ebsd_vae_b200 as E
    tmp = tmp_path_factory.mktemp("sample")
    base = R.synthetic_patterns(N_PAT, seed=99, size=144).numpy().astype(np.float64) / 255.0
    patterns = base[:, :140, :]  # (N, 140, 144): exercises the centre crop
    np.save(tmp / "sample_pattern.npy", patterns)
    sd = R.make_state_dict(42)
    torch.save(sd, tmp / "vae-best.pt")
    model = E.VariationalAutoEncoderRawData()
    model.load_state_dict(torch.load(tmp / "vae-best.pt", weights_only=True))
    cfg = E.IndexerConfig(pattern_path=tmp / "sample_pattern.npy",
                          angles_path=os.path.join(golden_dir, "anglefile_sample.txt"), device="cuda", top_n=10)
    indexer = E.DiffractionPatternIndexer(model, config=cfg)
    indexer.build_dictionary()
    # oracle pipeline
    u8 = np.stack([X.transform_u8(p) for p in patterns])
    mu, _ = R.encode(sd, R.u8_to_input(torch.from_numpy(u8)))
    angles = X.parse_rotation_angles(os.path.join(golden_dir, "anglefile_sample.txt"))[:N_PAT]
    return indexer, patterns, u8, mu.numpy(), angles, sd


def test_dictionary_contents(setup):
    indexer, patterns, u8, mu, angles, _ = setup
    db = indexer.db
    assert db.get_count() == N_PAT
    lat = db._latents[:N_PAT].cpu().numpy()
    want = T.normalize_rows(mu)
    rel = np.linalg.norm(lat - want, axis=1)
    assert rel.max() < 1e-3
    np.testing.assert_array_equal(db._eulers[:N_PAT].cpu().numpy(), angles)


def test_encode_apis_match_oracle(setup):
    indexer, patterns, u8, mu, _, _ = setup
    one = indexer.encode_pattern(patterns[3])
    assert isinstance(one, np.ndarray) and one.shape == (16,) and one.dtype == np.float32
    many = indexer.encode_patterns_batch(patterns[:10])
    assert many.shape == (10, 16)
    rel = np.linalg.norm(many - mu[:10], axis=1) / np.linalg.norm(mu[:10], axis=1)
    assert rel.max() < 1e-3
    np.testing.assert_allclose(one, many[3], rtol=0, atol=1e-5)  # plane sums are accumulated in tile order
    t = indexer.encode_pattern(torch.from_numpy(u8[3].astype(np.float32) / 255.0))  # tensors bypass the transform
    # float32 input takes conv0's plane statistics from an fp32 CUDA-core pass instead of the exact integer
    # autocorrelation of the uint8 path: ~1e-7 apart, which the fp8 correction operands can turn into ~1e-5 of |mu|
    np.testing.assert_allclose(t, one, rtol=0, atol=2e-4)


def test_index_pattern_reference_defaults(setup):
    indexer, patterns, u8, mu, angles, _ = setup
    res = indexer.index_pattern(patterns[5])
    # top_n=10 < min_required_matches=18 (chroma_db.py:266): the reference always fails here
    assert res.success is False and res.mean_orientation is None
    assert res.candidate_orientations.shape == (10, 3) and res.distances.shape == (10,)
    odot, oidx = T.topk(T.normalize_rows(mu), T.normalize_rows(mu[5:6]), 10)
    np.testing.assert_array_equal(res.candidate_orientations[0], angles[oidx[0, 0]])
    np.testing.assert_array_equal(res.best_orientation, angles[5])  # a dictionary pattern finds itself
    assert len(set(map(tuple, res.candidate_orientations)) & set(map(tuple, angles[oidx[0]]))) >= 9
    np.testing.assert_allclose(res.distances, 1.0 - odot[0], atol=2e-3)
    assert list(res.similar_indices) == list(range(10))  # radian threshold 3.0 lets every candidate through
    # the latent stays on the device inside index_pattern; the result still carries the host vector the reference
    # passes along (dp_indexer.py:210-214), and the two-call form gives the same answer
    lat = indexer.encode_pattern(patterns[5])
    assert isinstance(res.query_vector, np.ndarray) and res.query_vector.shape == (16,)
    np.testing.assert_array_equal(res.query_vector, lat)
    two = indexer.db.find_best_orientation(lat, top_n=10, orientation_threshold=indexer.config.orientation_threshold)
    np.testing.assert_array_equal(two.candidate_orientations, res.candidate_orientations)
    np.testing.assert_array_equal(two.distances, res.distances)


def test_search_on_reference_latents_is_bit_exact(setup):
    """North star: top-k indices on the REFERENCE's latents match exact brute force bit-exactly."""
    import ebsd_vae_b200 as E
    _, _, _, mu, angles, _ = setup
    db = E.LatentVectorDatabase()
    db.add_vectors(mu, angles)
    batch = db.find_best_orientations_batch(mu, top_n=10, min_required_matches=3, orientation_threshold=0.2)
    odot, oidx = T.topk(T.normalize_rows(mu), T.normalize_rows(mu), 10)
    np.testing.assert_array_equal(batch.indices, oidx)
    np.testing.assert_array_equal(batch.distances, np.float32(1.0) - odot)


def test_batch_indexing_consensus_matches_oracle(setup):
    indexer, patterns, _, _, _, _ = setup
    results = indexer.index_patterns_batch(patterns[:40], top_n=10, min_required_matches=3,
                                           orientation_threshold=0.2)
    assert len(results) == 40
    n_ok = 0
    for r in results:
        want = C.find_best_orientation(r.candidate_orientations, 0.2, 3, 3, mode="chroma")
        assert r.success == want.success
        np.testing.assert_array_equal(r.similar_indices, want.similar_indices)
        if r.success:
            n_ok += 1
            assert _misorientation_deg(r.mean_orientation, want.mean_orientation) < 0.1
        else:
            assert r.mean_orientation is None
        np.testing.assert_array_equal(r.best_orientation, r.candidate_orientations[0])
    assert n_ok > 0


def test_uint8_file_goes_through_float64_cast_like_the_reference(tmp_path, golden_dir, setup):
    """DPdataset casts to float64 before ToPILImage (data_module.py:132): uint8 files wrap to (-v) mod 256."""
    import ebsd_vae_b200 as E
    _, _, _, _, _, sd = setup
    raw = R.synthetic_patterns(4, seed=5).numpy()
    np.save(tmp_path / "u8.npy", raw)
    model = E.VariationalAutoEncoderRawData()
    model.load_state_dict(sd)
    cfg = E.IndexerConfig(pattern_path=tmp_path / "u8.npy",
                          angles_path=os.path.join(golden_dir, "anglefile_sample.txt"), device="cuda")
    idx = E.DiffractionPatternIndexer(model, config=cfg)
    idx.build_dictionary()
    u8 = np.stack([X.transform_u8(p.astype(np.float64)) for p in raw])
    mu, _ = R.encode(sd, R.u8_to_input(torch.from_numpy(u8)))
    lat = idx.db._latents[:4].cpu().numpy()
    assert np.linalg.norm(lat - T.normalize_rows(mu.numpy()), axis=1).max() < 1e-3


def test_dictionary_persistence_round_trip(tmp_path):
    """save -> load gives the same search results bit for bit; missing files raise like the FAISS class."""
    import ebsd_vae_b200 as E

    rng = np.random.default_rng(3)
    lat = rng.normal(size=(3000, 16)).astype(np.float32)
    lat[100] = lat[7]           # duplicate rows: the tie order must survive the round trip
    eul = rng.uniform(0, 360, size=(3000, 3))
    cfg = E.LatentVectorDatabaseConfig(collection_name="persist_test", persist_directory=str(tmp_path))
    db = E.LatentVectorDatabase(cfg)
    db.add_vectors(lat, eul)
    q = rng.normal(size=(50, 16)).astype(np.float32)
    q[0] = lat[7]
    before = db.find_best_orientations_batch(q, top_n=10, min_required_matches=3, orientation_threshold=3.0)
    path = db.save()
    assert path == tmp_path / "persist_test.npz" and path.exists()
    saved = np.load(path)
    assert saved["orientations"].shape == (3000, 3) and saved["orientations"].dtype == np.float64

    db2 = E.LatentVectorDatabase(cfg)
    assert db2.get_count() == 3000   # an existing store is reopened on construction (chroma_db.py:113-131, faiss_db.py:129)
    db2.load()
    assert db2.get_count() == 3000
    after = db2.find_best_orientations_batch(q, top_n=10, min_required_matches=3, orientation_threshold=3.0)
    np.testing.assert_array_equal(before.indices, after.indices)
    np.testing.assert_array_equal(before.distances, after.distances)
    np.testing.assert_array_equal(before.success, after.success)
    np.testing.assert_array_equal(before.mean_orientations, after.mean_orientations)
    assert list(before.indices[0][:2]) == [7, 100]

    db2.delete_persistence()
    assert not path.exists() and db2.get_count() == 0
    with pytest.raises(FileNotFoundError, match="NPZ file missing."):
        db2.load()


def test_faiss_mode_reports_inner_products_as_distances():
    """FaissLatentVectorDatabase carries the inner products in ``distances`` (faiss_db.py:216-256, 281-300), Chroma the
    cosine distance; ``get_top_n_orientations`` sorts ascending on whatever it is given, in both reference classes."""
    import ebsd_vae_b200 as E

    rng = np.random.default_rng(8)
    lat = rng.normal(size=(2000, 16)).astype(np.float32)
    eul = rng.uniform(0, 360, size=(2000, 3))
    out = {}
    for mode in ("chroma", "faiss"):
        db = E.LatentVectorDatabase(E.LatentVectorDatabaseConfig(mode=mode, persist_directory=None))
        db.add_vectors(lat, eul)
        out[mode] = db.find_best_orientation(lat[11], top_n=10, orientation_threshold=3.0, min_required_matches=3)
    np.testing.assert_array_equal(out["chroma"].candidate_orientations, out["faiss"].candidate_orientations)
    np.testing.assert_allclose(out["faiss"].distances, 1.0 - out["chroma"].distances, atol=1e-6)
    assert out["faiss"].distances[0] > 0.999 and (np.diff(out["faiss"].distances) <= 0).all()
    assert out["chroma"].distances[0] < 1e-3 and (np.diff(out["chroma"].distances) >= 0).all()
    with pytest.raises(ValueError, match="persist_directory is None"):
        db.save()


def test_database_reproduces_the_reference_chroma_path_end_to_end(golden_dir):
    """tests/golden/chroma_path.npz (oracle/make_golden_chroma_path.py): the unmodified ChromaLatentVectorDatabase --
    two add_vectors calls (id offset), query_similar, find_best_orientation and find_best_orientations_batch with the
    REAL query feeding the consensus (chroma_db.py:144-410) -- over an exact cosine stand-in for the chromadb
    collection.  The GPU dictionary gives the same ids, metadata, distances (1 - cos within 3e-7), success flags,
    similar indices, best and mean orientations."""
    import os

    import ebsd_vae_b200 as E

    g = np.load(os.path.join(golden_dir, "chroma_path.npz"))
    thr, mrm, mit = g["params"]
    db = E.LatentVectorDatabase(E.LatentVectorDatabaseConfig(persist_directory=None))
    db.add_vectors(g["latents"][:1700], g["orientations"][:1700], batch_size=1000)
    db.add_vectors(g["latents"][1700:], g["orientations"][1700:], batch_size=512)
    assert db.get_count() == len(g["latents"])
    for i in (0, 7, 31):
        res = db.query_similar(g["queries"][i], n_results=10)
        assert res["ids"][0] == [f"vec_{j}" for j in g["idx"][i]]
        np.testing.assert_allclose(res["distances"][0], g["dist"][i], rtol=0, atol=3e-7)
        assert [m["orientation_str"] for m in res["metadatas"][0]] == list(g["orientation_str"][i])
        assert [m["phi1"] for m in res["metadatas"][0]] == g["orientations"][g["idx"][i], 0].tolist()
    batch = db.find_best_orientations_batch(g["queries"], top_n=10, orientation_threshold=float(thr),
                                            min_required_matches=int(mrm), max_iterations=int(mit))
    np.testing.assert_array_equal(batch.indices, g["idx"])
    np.testing.assert_allclose(batch.distances, g["dist"], rtol=0, atol=3e-7)
    np.testing.assert_array_equal(batch.success, g["success"])
    np.testing.assert_array_equal(batch.best_orientations, g["best"])
    ok = g["success"]
    np.testing.assert_allclose(batch.mean_orientations[ok], g["mean"][ok], atol=1e-6)
    for i in range(len(batch)):
        r = batch[i]
        sim = np.zeros(10, bool)
        sim[r.similar_indices] = True
        np.testing.assert_array_equal(sim, g["similar"][i])
        assert (r.mean_orientation is None) == (not ok[i])
    one = db.find_best_orientation(g["queries"][2], top_n=10, orientation_threshold=float(thr),
                                   min_required_matches=int(mrm), max_iterations=int(mit))
    assert one.success == bool(ok[2])
    np.testing.assert_array_equal(one.candidate_orientations, g["orientations"][g["idx"][2]])


def test_indexer_reproduces_the_reference_end_to_end(tmp_path, golden_dir):
    """BASELINE configs[0] through the UNMODIFIED reference (tests/golden/indexer_path.npz, written by
    oracle/make_golden_indexer_path.py): DiffractionPatternIndexer.build_dictionary from a pattern .npy + angle file
    (DPDataModule, default transform, latice.model on the CPU), then index_pattern / index_patterns_batch.  The same
    script against this package gives the same dictionary (latents within 1e-3, angles bit for bit), the same candidate
    lists, distances within 2e-3, the same success flags and best orientations, mean orientations within 0.1 degree."""
    import ebsd_vae_b200 as E

    g = np.load(os.path.join(golden_dir, "indexer_path.npz"))
    thr, mrm, mit = g["params"]
    np.save(tmp_path / "sample_pattern.npy", (g["k_u8"].astype(np.float64) + 0.5) / 255.0)
    (tmp_path / "anglefile.txt").write_text(str(g["angle_text"]))
    model = E.VariationalAutoEncoderRawData()
    model.load_state_dict(R.make_state_dict(42))
    cfg = E.IndexerConfig(pattern_path=tmp_path / "sample_pattern.npy", angles_path=tmp_path / "anglefile.txt",
                          batch_size=16, device="cuda", top_n=10, orientation_threshold=float(thr))
    indexer = E.DiffractionPatternIndexer(model, db=E.LatentVectorDatabase(E.LatentVectorDatabaseConfig(persist_directory=None)),
                                          config=cfg)
    indexer.build_dictionary()
    n = len(g["k_u8"])
    assert indexer.db.get_count() == n
    np.testing.assert_array_equal(indexer.db._eulers[:n].cpu().numpy(), g["dict_angles"])
    want = g["dict_latents"] / np.linalg.norm(g["dict_latents"], axis=1, keepdims=True)
    got = indexer.db._latents[:n].cpu().numpy()
    assert np.linalg.norm(got - want, axis=1).max() < 1e-3

    queries = (g["k_u8"][:12].astype(np.float64) + 0.5) / 255.0
    lat = indexer.encode_pattern(queries[4])
    assert np.linalg.norm(lat - g["one_latent"]) / np.linalg.norm(g["one_latent"]) < 1e-3
    one = indexer.index_pattern(queries[4])                 # the reference's defaults: 18 matches of 10 -> no consensus
    assert one.success is False and bool(g["one_success"]) is False and one.mean_orientation is None
    np.testing.assert_array_equal(one.candidate_orientations, g["one_candidates"])
    np.testing.assert_allclose(one.distances, g["one_distances"], rtol=0, atol=2e-3)

    batch = indexer.index_patterns_batch(queries, top_n=10, orientation_threshold=float(thr),
                                         min_required_matches=int(mrm), max_iterations=int(mit))
    assert 0 < g["batch_success"].sum() < len(queries)
    np.testing.assert_array_equal(batch.success, g["batch_success"])
    np.testing.assert_array_equal(batch.candidate_orientations, g["batch_candidates"])
    np.testing.assert_array_equal(batch.best_orientations, g["batch_best"])
    np.testing.assert_allclose(batch.distances, g["batch_distances"], rtol=0, atol=2e-3)
    for i in np.flatnonzero(g["batch_success"]):
        assert _misorientation_deg(batch.mean_orientations[i], g["batch_mean"][i]) < 0.1
