"""Golden vectors for BASELINE configs[0], end to end through the reference: tests/golden/indexer_path.npz.

Test infrastructure; runs only in the build container (needs /root/reference).  The UNMODIFIED
``DiffractionPatternIndexer`` (latice/index/dp_indexer.py:59-297) with the unmodified ``VariationalAutoEncoderRawData``
(seed-42 random init = oracle/encoder_ref.make_state_dict(42); the real vae-best.pt is not in the checkout) on the CPU:
``build_dictionary()`` from a pattern ``.npy`` + an angle file through ``DPDataModule`` / ``DPdataset`` / the default
transform and the model, then ``index_pattern`` and ``index_patterns_batch`` on dictionary patterns and on noisy copies.
The dictionary is the unmodified ``ChromaLatentVectorDatabase`` over the exact cosine stand-in for the chromadb
collection of oracle/make_golden_chroma_path.py (the wheel is not installable; pytorch_lightning is a stub class).

The fixture stores the uint8 gray levels k of the patterns; the reference was fed the float64 array (k + 0.5) / 255,
which its transform quantises back to exactly k.

    python oracle/make_golden_indexer_path.py        # rewrites tests/golden/indexer_path.npz
"""
from __future__ import annotations

import os
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")

N, H, W, K = 64, 132, 136, 10
PARAMS = dict(top_n=K, orientation_threshold=0.6, min_required_matches=3, max_iterations=3)


def main() -> None:
    from oracle import encoder_ref, make_golden_chroma_path as chroma_stub  # noqa: F401  (installs nothing by itself)
    import types

    # the chromadb stand-in of make_golden_chroma_path.py
    class _Client:
        def __init__(self, *a, **k):
            self.collections = {}

        def get_collection(self, name):
            if name not in self.collections:
                raise ValueError(f"Collection {name} does not exist.")
            return self.collections[name]

        def create_collection(self, name, metadata=None):
            self.collections[name] = chroma_stub.ExactCosineCollection(name, metadata or {})
            return self.collections[name]

    errs = types.ModuleType("chromadb.errors")
    errs.InvalidCollectionException = type("InvalidCollectionException", (Exception,), {})
    chroma = types.ModuleType("chromadb")
    chroma.Client, chroma.PersistentClient, chroma.errors = _Client, _Client, errs
    sys.modules["chromadb"], sys.modules["chromadb.errors"] = chroma, errs
    from oracle import refload

    ref = refload.load()
    rng = np.random.default_rng(11)
    k_u8 = encoder_ref.synthetic_patterns(N, seed=7, size=144).numpy()[:, :H, :W]
    angles = np.round(rng.uniform(0, 1, size=(N, 3)) * np.array([360.0, 90.0, 360.0]), 4)
    # neighbours in the file are neighbours in orientation AND in pattern space: pattern 2i+1 is a noisy copy of 2i
    k_u8[1::2] = np.clip(k_u8[0::2].astype(np.int16) + rng.integers(-3, 4, size=k_u8[0::2].shape), 0, 255).astype(np.uint8)
    angles[1::2] = angles[0::2] + np.round(rng.normal(size=(N // 2, 3)) * 0.5, 4)
    with tempfile.TemporaryDirectory() as tmp:
        ppath, apath = os.path.join(tmp, "sample_pattern.npy"), os.path.join(tmp, "anglefile.txt")
        np.save(ppath, (k_u8.astype(np.float64) + 0.5) / 255.0)
        lines = ["eu\n", f"{N}\n"] + [" ".join(repr(float(v)) for v in a) + "\n" for a in angles]
        with open(apath, "w") as fh:
            fh.writelines(lines)
        torch.manual_seed(42)
        model = ref.model.VariationalAutoEncoderRawData()
        sd = encoder_ref.make_state_dict(42)
        for key, val in sd.items():
            assert torch.equal(model.state_dict()[key], val), key
        db = ref.chroma_db.ChromaLatentVectorDatabase(ref.chroma_db.LatentVectorDatabaseConfig(persist_directory=None))
        cfg = ref.dp_indexer.IndexerConfig(pattern_path=ppath, angles_path=apath, batch_size=16, device="cpu", top_n=K,
                                           orientation_threshold=PARAMS["orientation_threshold"])
        indexer = ref.dp_indexer.DiffractionPatternIndexer(model, db=db, config=cfg)
        indexer.build_dictionary()
        col = db.collection
        assert col.count() == N
        dict_latents = col.rows.copy()                                        # what the reference stored (float32)
        dict_angles = np.array([[m["phi1"], m["Phi"], m["phi2"]] for m in col.metas])
        queries = (k_u8[:12].astype(np.float64) + 0.5) / 255.0               # ndarray input: goes through the transform
        one = indexer.index_pattern(queries[4])                               # reference defaults: 18 matches of 10 -> fails
        enc_one = indexer.encode_pattern(queries[4])
        batch = indexer.index_patterns_batch(queries, **PARAMS)
        out = dict(
            k_u8=k_u8, angle_text=np.array("".join(lines)), dict_latents=dict_latents, dict_angles=dict_angles,
            one_success=np.array(one.success), one_candidates=one.candidate_orientations, one_distances=one.distances,
            one_latent=enc_one,
            batch_success=np.array([r.success for r in batch]),
            batch_best=np.array([r.best_orientation for r in batch]),
            batch_mean=np.array([r.mean_orientation if r.success else [np.nan] * 3 for r in batch]),
            batch_candidates=np.array([r.candidate_orientations for r in batch]),
            batch_distances=np.array([r.distances for r in batch]),
            params=np.array([PARAMS["orientation_threshold"], PARAMS["min_required_matches"], PARAMS["max_iterations"]]),
        )
    np.savez_compressed(os.path.join(GOLDEN, "indexer_path.npz"), **out)
    print("indexer_path.npz: batch success", int(out["batch_success"].sum()), "of", len(batch), "; index_pattern success",
          bool(one.success), "; size", os.path.getsize(os.path.join(GOLDEN, "indexer_path.npz")))


if __name__ == "__main__":
    main()
