"""The oracle chain (normalise -> exact top-k -> consensus) against the reference's own Chroma path, end to end.

tests/golden/chroma_path.npz was written by oracle/make_golden_chroma_path.py: the unmodified
ChromaLatentVectorDatabase (add_vectors, query_similar, find_best_orientation(s_batch),
/root/reference/latice/index/chroma_db.py:144-410) over an exact cosine stand-in for the chromadb collection.
"""
import os

import numpy as np

from oracle import consensus_ref, topk_ref as T


def test_oracle_chain_reproduces_the_reference_chroma_path(golden_dir):
    g = np.load(os.path.join(golden_dir, "chroma_path.npz"))
    rows = T.normalize_rows(g["latents"].astype(np.float32))           # hnswlib stores float32 rows
    queries = T.normalize_rows(g["queries"].astype(np.float32))
    dots, idx = T.topk(rows, queries, 10)
    np.testing.assert_array_equal(idx, g["idx"])                        # same rows, same order (ids vec_{row})
    np.testing.assert_allclose(1.0 - dots.astype(np.float64), g["dist"], rtol=0, atol=3e-7)   # Chroma's 1 - cos
    thr, mrm, mit = g["params"]
    assert 0 < g["success"].sum() < len(g["success"])                   # the fixture holds both outcomes
    for i in range(len(idx)):
        cand = g["orientations"][idx[i]]
        # the metadata strings the reference stores (chroma_db.py:183-190)
        assert [",".join(map(str, o)) for o in cand.tolist()] == list(g["orientation_str"][i])
        o = consensus_ref.find_best_orientation(cand, float(thr), int(mrm), int(mit))
        assert o.success == bool(g["success"][i])
        sim = np.zeros(10, bool)
        sim[o.similar_indices] = True
        np.testing.assert_array_equal(sim, g["similar"][i])
        np.testing.assert_array_equal(cand[0], g["best"][i])            # best_orientation = nearest candidate (Chroma)
        if o.success:
            np.testing.assert_allclose(o.mean_orientation, g["mean"][i], atol=1e-9)


def test_oracle_chain_reproduces_the_reference_indexer_end_to_end(golden_dir, tmp_path):
    """tests/golden/indexer_path.npz (oracle/make_golden_indexer_path.py): the unmodified DiffractionPatternIndexer on
    the CPU -- transform, model, Chroma path.  The oracle restatements of every stage, chained, give its dictionary
    and its answers (the oracle encoder is torch fp32 like the reference: latents to 1e-5)."""
    import torch

    from oracle import encoder_ref, transform_ref

    g = np.load(os.path.join(golden_dir, "indexer_path.npz"))
    x = (g["k_u8"].astype(np.float64) + 0.5) / 255.0
    u8 = np.stack([transform_ref.transform_u8(p) for p in x[:16]])
    np.testing.assert_array_equal(u8, g["k_u8"][:16, 2:130, 4:132])           # centre crop of 132 x 136, same gray levels
    sd = encoder_ref.make_state_dict(42)
    crop = torch.from_numpy(np.ascontiguousarray(g["k_u8"][:, 2:130, 4:132]))
    mu, _ = encoder_ref.encode(sd, encoder_ref.u8_to_input(crop))
    mu = mu.numpy()
    rel = np.linalg.norm(mu - g["dict_latents"], axis=1) / np.linalg.norm(g["dict_latents"], axis=1)
    assert rel.max() < 1e-4
    (tmp_path / "angles.txt").write_text(str(g["angle_text"]))
    angles = transform_ref.parse_rotation_angles(str(tmp_path / "angles.txt"))
    np.testing.assert_array_equal(angles, g["dict_angles"])
    rows = T.normalize_rows(g["dict_latents"])
    dots, idx = T.topk(rows, T.normalize_rows(mu[:12]), 10)
    np.testing.assert_array_equal(angles[idx], g["batch_candidates"])
    np.testing.assert_allclose(1.0 - dots, g["batch_distances"], atol=2e-5)
    thr, mrm, mit = g["params"]
    for i in range(12):
        o = consensus_ref.find_best_orientation(angles[idx[i]], float(thr), int(mrm), int(mit))
        assert o.success == bool(g["batch_success"][i])
        if o.success:
            np.testing.assert_allclose(o.mean_orientation, g["batch_mean"][i], atol=1e-6)
