"""Generate tests/golden/* by running the UNMODIFIED reference (build container only).

Test infrastructure.  Usage:  python -m oracle.make_golden
Needs /root/reference (see oracle/refload.py).  The outputs are small and committed; the GPU box
never runs this script.

Fixtures written:
  encoder_seed42.npz  -- 4 seeded uint8 patterns and the reference's mu / logvar for them, from
                         ``torch.manual_seed(42); VariationalAutoEncoderRawData()`` on torch CPU fp32
                         (latice/model.py), plus per-tensor float64 checksums of the hot weights.
  transform.npz       -- reference ``create_default_transform((128,128))`` outputs (as uint8) for seeded
                         inputs of several sizes/dtypes (latice/data_module.py:17-33).
  consensus.npz       -- ``ChromaLatentVectorDatabase.find_best_orientation`` (latice/index/chroma_db.py:261-342)
                         and the FAISS twin (latice/index/faiss_db.py:258-372) on seeded candidate sets,
                         incl. the reference's own known-answer case (tests/index/test_chroma_db.py:306-382).
  consensus_short.npz -- the Chroma class on candidate lists SHORTER than ``max_iterations``: the reference indexes
                         ``orientations[iteration]`` unguarded (chroma_db.py:302-303) but leaves the loop on success
                         (:324-326), so it raises IndexError only when every available reference failed.
  l2_normalize.npz    -- ``FaissLatentVectorDatabase._l2_normalize`` (latice/index/faiss_db.py:109-113, pure numpy) on
                         seeded float32 rows incl. a zero row: pins the normalisation leg of the search oracle.
  anglefile_sample.txt + angles_sample.npy -- a regenerated copy of the 625-row sample angle file and what the
                         reference parser (latice/data_module.py:87-116) returns for the original.
"""
from __future__ import annotations

import logging
import os
import sys
import types
import warnings
from unittest.mock import MagicMock, patch

import numpy as np
import torch

from oracle import encoder_ref, refload

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

TRANSFORM_CASES = (  # (seed, height, width, dtype, scale)
    (11, 128, 128, "float64", 1.0),
    (12, 160, 200, "float64", 1.0),
    (13, 129, 131, "float64", 1.0),
    (14, 100, 90, "float64", 1.0),
    (15, 127, 128, "float32", 1.0),
    (16, 300, 128, "float64", 1.0),
    (17, 133, 130, "float64", 1.5),  # values above 1.0 wrap modulo 256 after the uint8 cast
    (18, 140, 150, "uint8", 1.0),
)


def transform_input(seed: int, h: int, w: int, dtype: str, scale: float) -> np.ndarray:
    rng = np.random.default_rng(seed)
    if dtype == "uint8":
        return rng.integers(0, 256, size=(h, w), dtype=np.uint8)
    return (rng.random((h, w)) * scale).astype(dtype)


def consensus_cases(n_random: int = 240):
    """Seeded candidate sets: (eulers[k,3], threshold, min_required, max_iterations)."""
    from scipy.spatial.transform import Rotation as R

    rng = np.random.default_rng(20241018)
    cases = []
    known = np.array(
        [[30.0, 45.0, 60.0], [32.0, 44.0, 61.0], [31.0, 46.0, 59.0], [29.0, 45.0, 58.0], [28.0, 43.0, 62.0],
         [90.0, 90.0, 90.0]]
    )
    for thr, mrm, mit in ((0.3, 3, 2), (0.01, 5, 2), (3.0, 5, 3), (3.0, 18, 3), (0.3, 3, 3)):
        cases.append((known, thr, mrm, mit))
    # sample angle file style candidates: (0, i, 0) incl. gimbal rows and 360-degree aliases
    for start in (0, 85, 175, 355):
        cand = np.array([[0.0, float(start + j), 0.0] for j in range(10)])
        cases.append((cand, 0.2, 3, 3))
        cases.append((cand, 3.0, 5, 3))
    sym = refload.load().constants.QUAT_SYM
    for i in range(n_random):
        k = int(rng.choice([3, 6, 10, 20]))
        centre = R.random(random_state=int(rng.integers(1 << 31)))
        spread = float(rng.choice([0.01, 0.05, 0.2]))
        rots = []
        for j in range(k):
            r = R.from_rotvec(rng.normal(size=3) * spread) * centre
            kind = rng.random()
            if kind < 0.35:  # scrambled by a cubic operator
                r = r * sym[int(rng.integers(24))]
            elif kind < 0.5:  # outlier
                r = R.random(random_state=int(rng.integers(1 << 31)))
            rots.append(r.as_euler("zxz", degrees=True))
        thr = float(rng.choice([0.05, 0.3, 1.0, 3.0]))
        mrm = int(rng.choice([2, 3, 5, 8]))
        mit = int(rng.choice([1, 2, 3]))
        cases.append((np.array(rots), thr, mrm, min(mit, k)))
    return cases


def run_chroma(ref, cand, thr, mrm, mit):
    meta = [[{"phi1": float(o[0]), "Phi": float(o[1]), "phi2": float(o[2])} for o in cand]]
    res = {"metadatas": meta, "distances": [list(np.linspace(0.1, 0.9, len(cand)))]}
    cls = ref.chroma_db.ChromaLatentVectorDatabase
    with patch("chromadb.PersistentClient", return_value=MagicMock()), patch.object(cls, "query_similar",
                                                                                     return_value=res):
        db = cls()
        return db.find_best_orientation(np.ones(16), top_n=len(cand), orientation_threshold=thr,
                                        min_required_matches=mrm, max_iterations=mit)


def run_faiss(faiss_db, cand, thr_deg, mrm, mit):
    cls = faiss_db.FaissLatentVectorDatabase
    db = cls.__new__(cls)
    db._orientations = list(cand)
    ret = (np.linspace(0.9, 0.1, len(cand)), np.arange(len(cand)))
    with patch.object(cls, "query_similar", return_value=ret):
        return db.find_best_orientation(np.ones(16, dtype=np.float32), top_n=len(cand),
                                        orientation_threshold=thr_deg, min_required_matches=mrm, max_iterations=mit)


def main() -> None:
    logging.disable(logging.CRITICAL)
    warnings.filterwarnings("ignore")
    ref = refload.load()
    os.makedirs(GOLDEN, exist_ok=True)

    # ---- encoder
    torch.manual_seed(42)
    model = ref.model.VariationalAutoEncoderRawData().eval()
    sd = model.state_dict()
    pats = encoder_ref.synthetic_patterns(4, seed=1234)
    x = encoder_ref.u8_to_input(pats)
    with torch.no_grad():
        _, _, mu, _ = model(x)
        logvar = model.logvar(model.encoder(x).flatten(1, -1))
    checks = np.array([sd[k].double().sum().item() for k in encoder_ref.HOT_KEYS])
    abs_checks = np.array([sd[k].double().abs().sum().item() for k in encoder_ref.HOT_KEYS])
    np.savez_compressed(os.path.join(GOLDEN, "encoder_seed42.npz"), patterns=pats.numpy(), mu=mu.numpy(),
                        logvar=logvar.numpy(), weight_sums=checks, weight_abs_sums=abs_checks,
                        keys=np.array(encoder_ref.HOT_KEYS))

    # ---- transform
    tf = ref.data_module.create_default_transform((128, 128))
    outs = []
    for case in TRANSFORM_CASES:
        t = tf(transform_input(*case))
        assert t.shape == (1, 128, 128) and t.dtype == torch.float32
        k = torch.round(t[0] * 255).to(torch.uint8)
        assert torch.equal(k.float() / 255, t[0])
        outs.append(k.numpy())
    np.savez_compressed(os.path.join(GOLDEN, "transform.npz"), outputs=np.stack(outs),
                        cases=np.array([list(map(str, c)) for c in TRANSFORM_CASES]))

    # ---- angle file
    src = os.path.join(refload.REFERENCE_ROOT, "data", "anglefile_sample.txt")
    parsed = ref.data_module.DPdataset._parse_rotation_angles(None, src).to_numpy()
    np.save(os.path.join(GOLDEN, "angles_sample.npy"), parsed)
    with open(os.path.join(GOLDEN, "anglefile_sample.txt"), "w") as fh:
        fh.write("eu\n625\n")
        for i in range(625):
            fh.write(f"0 {i} 0\n")
    regenerated = ref.data_module.DPdataset._parse_rotation_angles(None, os.path.join(GOLDEN, "anglefile_sample.txt"))
    assert np.array_equal(regenerated.to_numpy(), parsed)

    # ---- consensus
    if "faiss" not in sys.modules:
        sys.modules["faiss"] = types.ModuleType("faiss")
    import latice.index.faiss_db as faiss_db

    cases = consensus_cases()
    kmax = max(len(c[0]) for c in cases)
    n = len(cases)
    cand = np.full((n, kmax, 3), np.nan)
    ks = np.zeros(n, dtype=np.int64)
    params = np.zeros((n, 3))
    out = {m: dict(success=np.zeros(n, bool), mean=np.full((n, 3), np.nan), best=np.full((n, 3), np.nan),
                   similar=np.zeros((n, kmax), bool)) for m in ("chroma", "faiss")}
    for i, (c, thr, mrm, mit) in enumerate(cases):
        k = len(c)
        cand[i, :k] = c
        ks[i] = k
        params[i] = (thr, mrm, mit)
        rc = run_chroma(ref, c, thr, mrm, mit)
        rf = run_faiss(faiss_db, c, np.degrees(thr), mrm, mit)  # same physical threshold, in degrees
        for m, r in (("chroma", rc), ("faiss", rf)):
            o = out[m]
            o["success"][i] = r.success
            o["best"][i] = r.best_orientation
            if r.mean_orientation is not None:
                o["mean"][i] = r.mean_orientation
            if r.similar_indices is not None:
                o["similar"][i, np.asarray(r.similar_indices, dtype=np.int64)] = True
    np.savez_compressed(
        os.path.join(GOLDEN, "consensus.npz"), cand=cand, k=ks, params=params,
        **{f"{m}_{key}": v for m, o in out.items() for key, v in o.items()},
    )
    # ---- consensus on lists shorter than max_iterations (lazy IndexError of the Chroma class)
    rng = np.random.default_rng(77)
    short = []
    base = np.array([[30.0, 45.0, 60.0], [30.5, 45.2, 60.1], [120.0, 10.0, 200.0]])
    for k_s in (1, 2, 3):
        for thr_s in (0.05, 3.0):
            for mrm_s in (1, 2, 3):
                for mit_s in (2, 3, 5):
                    short.append((base[:k_s] + rng.normal(scale=0.01, size=(k_s, 3)), thr_s, mrm_s, mit_s))
    ns = len(short)
    s_cand = np.full((ns, 3, 3), np.nan)
    s_k = np.zeros(ns, dtype=np.int64)
    s_params = np.zeros((ns, 3))
    s_raised = np.zeros(ns, bool)
    s_success = np.zeros(ns, bool)
    s_mean = np.full((ns, 3), np.nan)
    s_similar = np.zeros((ns, 3), bool)
    for i, (c, thr_s, mrm_s, mit_s) in enumerate(short):
        s_cand[i, : len(c)] = c
        s_k[i] = len(c)
        s_params[i] = (thr_s, mrm_s, mit_s)
        try:
            r = run_chroma(ref, c, thr_s, mrm_s, mit_s)
        except IndexError:
            s_raised[i] = True
            continue
        s_success[i] = r.success
        if r.mean_orientation is not None:
            s_mean[i] = r.mean_orientation
        if r.similar_indices is not None:
            s_similar[i, np.asarray(r.similar_indices, dtype=np.int64)] = True
    assert s_raised.any() and (~s_raised).any()
    np.savez_compressed(os.path.join(GOLDEN, "consensus_short.npz"), cand=s_cand, k=s_k, params=s_params,
                        raised=s_raised, success=s_success, mean=s_mean, similar=s_similar)

    # ---- the FAISS class's row normalisation (pure numpy, so it runs here although faiss itself is a stub)
    rng = np.random.default_rng(2024)
    rows = (rng.normal(size=(4096, 16)) * rng.uniform(1e-3, 1e3, size=(4096, 1))).astype(np.float32)
    rows[17] = 0.0
    fdb = faiss_db.FaissLatentVectorDatabase.__new__(faiss_db.FaissLatentVectorDatabase)
    normed = fdb._l2_normalize(rows.copy())
    assert normed.dtype == np.float32
    np.savez_compressed(os.path.join(GOLDEN, "l2_normalize.npz"), rows=rows, normalized=normed)

    print("golden fixtures written to", GOLDEN, "cases:", n,
          "chroma success:", int(out["chroma"]["success"].sum()), "faiss success:", int(out["faiss"]["success"].sum()))


if __name__ == "__main__":
    main()
