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
