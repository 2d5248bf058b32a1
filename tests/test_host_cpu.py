"""CPU-side tests: host logic of the drop-in API and the C-ABI surface (no GPU compute)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

import ebsd_vae_b200 as E
from ebsd_vae_b200 import _native, transform
from oracle import transform_ref
from oracle.make_golden import TRANSFORM_CASES, transform_input

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "ebsd_b200.h")).read()
    declared = set(re.findall(r"\b(ebsd_[a-z0-9_]+)\s*\(", header))
    declared -= {"ebsd_encoder"}
    assert declared == set(_native.SYMBOLS), declared ^ set(_native.SYMBOLS)
    lib = _native.load()
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.ebsd_abi_version() == 1


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the behaviour on a box without a GPU")
def test_calls_fail_loudly_without_a_gpu():
    lib = _native.load()
    rc = lib.ebsd_normalize_rows(None, 0, 16, None)
    assert rc != 0 and "CUDA" in _native.last_error()
    with pytest.raises(RuntimeError):
        E.LatentVectorDatabase().add_vectors(np.zeros((2, 16), np.float32), np.zeros((2, 3)))
    with pytest.raises(RuntimeError):
        E.DiffractionPatternIndexer(E.VariationalAutoEncoderRawData())
    with pytest.raises(RuntimeError):
        E.EncoderEngine(E.VariationalAutoEncoderRawData().state_dict(), "cpu")


def test_transform_matches_golden_and_oracle(golden_dir):
    g = np.load(os.path.join(golden_dir, "transform.npz"))
    for case, want in zip(TRANSFORM_CASES, g["outputs"]):
        x = transform_input(*case)
        got = transform.transform_batch_u8(x, (128, 128))
        assert got.shape == (1, 128, 128) and got.dtype == np.uint8
        np.testing.assert_array_equal(got[0], want, err_msg=str(case))
        np.testing.assert_array_equal(got[0], transform_ref.transform_u8(x, (128, 128)))


def test_transform_batch_and_bad_dtype():
    rng = np.random.default_rng(5)
    x = rng.random((6, 131, 140))
    got = transform.transform_batch_u8(x)
    for i in range(6):
        np.testing.assert_array_equal(got[i], transform_ref.transform_u8(x[i]))
    with pytest.raises(TypeError):
        transform.transform_batch_u8(np.zeros((2, 128, 128), dtype=np.int64))


def test_angle_parser(golden_dir, tmp_path):
    got = transform.parse_rotation_angles(os.path.join(golden_dir, "anglefile_sample.txt"))
    np.testing.assert_array_equal(got, np.load(os.path.join(golden_dir, "angles_sample.npy")))
    with pytest.raises(FileNotFoundError):
        transform.parse_rotation_angles(tmp_path / "missing.txt")
    # malformed files behave as pd.DataFrame(rows, columns=[z1, x, z2]).astype(float) does in the reference
    # (latice/data_module.py:100-110): short / blank rows are NaN-padded, too long or non-numeric rows raise
    pd = pytest.importorskip("pandas")
    cases = {"short.txt": "eu\n2\n0 1 0\n0 1\n", "blank.txt": "eu\n2\n0  1 0\n\n", "long.txt": "eu\n2\n0 1 0 5\n0 1 0\n",
             "allshort.txt": "eu\n2\n0 1\n0\n", "text.txt": "eu\n2\n0 a 0\n", "empty.txt": "eu\n2\n"}
    for name, text in cases.items():
        f = tmp_path / name
        f.write_text(text)
        rows = [[t for t in line.strip().split(" ") if t] for line in text.splitlines(keepends=True)[2:]]
        try:
            want = pd.DataFrame(rows, columns=["z1", "x", "z2"]).astype(float).to_numpy().reshape(-1, 3)
        except Exception:
            with pytest.raises(ValueError, match="Failed to parse rotation angles file"):
                transform.parse_rotation_angles(f)
            continue
        np.testing.assert_array_equal(transform.parse_rotation_angles(f), want)


def test_native_angle_parser_equals_python_restatement(tmp_path, monkeypatch):
    """ebsd_parse_angle_text (host C++, no GIL) takes the regular files and returns exactly what the Python restatement
    of latice/data_module.py:100-110 returns; anything irregular is left to that restatement (returns None)."""
    rng = np.random.default_rng(3)
    regular = "Euler angles\n2000\n" + "".join(
        f"{float(a)!r} {b:.4e}   +{abs(c):.3f}  \n" for a, b, c in rng.uniform(-400, 400, (2000, 3)))
    cases = {"regular": (regular, True), "crlf": ("h\r\nh\r\n1 2 3\r\n4 5 6\r\n", True),
             "no_final_newline": ("h\nh\n1 2 3\n4. .5 -6e-3", True), "only_header": ("h\nh\n", True),
             "padded": ("h\nh\n   1 2 3   \n", True), "blank_line": ("h\nh\n1 2 3\n\n", False),
             "short_row": ("h\nh\n1 2 3\n4 5\n", False), "nan": ("h\nh\nnan 2 3\n", False),
             "underscore": ("h\nh\n1_0 2 3\n", False), "overflow": ("h\nh\n1e999 1e-999 3\n", False),
             "lone_cr": ("h\nh\n1 2 3\r4 5 6\n", False), "non_ascii": ("h\nh\n1 2 \u00b53\n", None)}
    for name, (text, native) in cases.items():
        f = tmp_path / f"{name}.txt"
        with open(f, "w", newline="", encoding="utf-8") as fh:
            fh.write(text)
        fast = transform._parse_rotation_angles_native(f)
        if native is not None:
            assert (fast is not None) == native, name
        with monkeypatch.context() as m:
            m.setattr(transform, "_parse_rotation_angles_native", lambda p: None)
            try:
                want = transform.parse_rotation_angles(f)
            except ValueError:
                with pytest.raises(ValueError, match="Failed to parse rotation angles file"):
                    m.undo()
                    transform.parse_rotation_angles(f)
                continue
        np.testing.assert_array_equal(transform.parse_rotation_angles(f), want)
        if fast is not None:
            np.testing.assert_array_equal(fast, want)


def test_config_defaults_match_reference():
    c = E.IndexerConfig()
    assert (c.batch_size, c.latent_dim, c.random_seed, tuple(c.image_size), c.top_n, c.orientation_threshold) == (
        64, 16, 42, (128, 128), 20, 3.0)
    d = E.LatentVectorDatabaseConfig()
    assert (d.collection_name, d.persist_directory, d.dimension) == ("latent_vectors", ".chroma_db", 16)
    assert E.ChromaLatentVectorDatabase is E.LatentVectorDatabase


def test_db_validation_messages():
    db = E.LatentVectorDatabase()
    assert db.get_count() == 0 and db.collection_name == "latent_vectors" and db.dimension == 16
    with pytest.raises(ValueError, match="Number of latent vectors and orientations must match"):
        db.add_vectors(np.zeros((3, 16)), np.zeros((2, 3)))
    with pytest.raises(ValueError, match="Expected latent vectors of dimension 16, got 8"):
        db.add_vectors(np.zeros((3, 8)), np.zeros((3, 3)))
    with pytest.raises(ValueError, match="Expected query vector of dimension 16, got 8"):
        db.query_similar(np.zeros(8))
    with pytest.raises(ValueError, match="Expected query vector of dimension"):
        db.find_best_orientation(np.zeros(8))


def test_orientation_result_top_n():
    rng = np.random.default_rng(0)
    cand = rng.random((5, 3)) * 360
    r = E.OrientationResult(query_vector=rng.random(16), best_orientation=cand[0], candidate_orientations=cand,
                            distances=np.array([0.5, 0.1, 0.3, 0.2, 0.4]))
    np.testing.assert_array_equal(r.get_top_n_orientations(3), cand[[1, 3, 2]])
    assert r.get_top_n_orientations(10).shape == (5, 3)
    r.distances = None
    np.testing.assert_array_equal(r.get_top_n_orientations(2), cand[:2])


def test_weight_extraction_accepts_reference_layouts():
    from oracle import encoder_ref
    sd = encoder_ref.make_state_dict(3)
    full = dict(sd)
    full["decoder.1.0.weight"] = torch.zeros(1)  # extra keys of the full VAE are ignored
    hot = E.model.extract_hot_state_dict(full)
    assert list(hot) == list(E.model.HOT_KEYS)
    lightning = {"state_dict": {"model." + k: v for k, v in sd.items()}}
    assert torch.equal(E.model.extract_hot_state_dict(lightning)["mu.0.bias"], sd["mu.0.bias"])
    m = E.VariationalAutoEncoderRawData()
    m.load_state_dict(full)
    assert torch.equal(m.state_dict()["encoder.13.0.weight"], sd["encoder.13.0.weight"])
    with pytest.raises(KeyError):
        E.model.extract_hot_state_dict({"foo": torch.zeros(1)})


def test_copy_slices_cover_the_batch_with_a_short_first_slice():
    """DiffractionPatternIndexer._copy_slices: contiguous cover of [0, b), no empty slice, nothing longer than one
    encoder pass, and a first slice of at most half a pass once the batch is larger than that."""
    from ebsd_vae_b200.dp_indexer import DiffractionPatternIndexer as D

    for b in list(range(1, 40)) + [700, 740, 741, 1479, 1480, 1481, 2959, 2960, 2961, 8880, 10000, 12345, 100000]:
        sl = D._copy_slices(b)
        assert sl[0][0] == 0 and sl[-1][1] == b
        assert all(x[1] == y[0] for x, y in zip(sl, sl[1:]))
        assert all(0 < e - a <= D.ENCODE_PASS for a, e in sl)
        if b > D.ENCODE_PASS // 2:
            assert sl[0][1] - sl[0][0] <= (D.ENCODE_PASS + 1) // 2 + 1


def test_faiss_twin_names_defaults_and_empty_index(tmp_path):
    """FaissLatentVectorDatabase / FaissLatentVectorDatabaseConfig keep the reference's names, defaults and its
    empty-index behaviour (latice/index/faiss_db.py:34-46, 232-234, 280-291) -- no GPU is touched by an empty index."""
    import ebsd_vae_b200 as E

    cfg = E.FaissLatentVectorDatabaseConfig()
    assert cfg.npz_path == "faiss_index.npz" and cfg.dimension == 16
    db = E.FaissLatentVectorDatabase(E.FaissLatentVectorDatabaseConfig(npz_path=str(tmp_path / "store" / "idx.npz")))
    assert db.get_count() == 0 and db.npz_path == tmp_path / "store" / "idx.npz" and db.config.mode == "faiss"
    sims, idx = db.query_similar(np.ones(16))
    assert sims.size == 0 and idx.size == 0
    res = db.find_best_orientation(np.ones((1, 16)))
    assert res.success is False and np.isnan(res.best_orientation).all() and res.mean_orientation is None
    assert res.candidate_orientations.size == 0 and res.similar_indices is None and res.query_vector.shape == (16,)
    with pytest.raises(FileNotFoundError, match="NPZ file missing."):
        db.load()
