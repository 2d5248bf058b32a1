"""Oracle: quaternion consensus over the top-k candidate orientations (float64 numpy).

Test infrastructure (see oracle/__init__.py).  A quaternion-only restatement -- no scipy
objects, no intermediate Euler round trips -- of

* ``ChromaLatentVectorDatabase.find_best_orientation``  (latice/index/chroma_db.py:261-342)
  mode "chroma": raw misorientation in RADIANS is compared with ``orientation_threshold``
  (chroma_db.py:307-310), no symmetry before thresholding, ``best_orientation`` stays candidate 0,
  ``similar_indices`` are those of the last iteration executed;
* ``FaissLatentVectorDatabase.find_best_orientation``   (latice/index/faiss_db.py:258-372)
  mode "faiss": misorientation converted to DEGREES first (faiss_db.py:308-313), iterations
  clamped to the number of candidates (faiss_db.py:302), ``best_orientation`` becomes the mean on
  success (faiss_db.py:338-342);
* ``_find_symmetry_equivalent_orientation``              (chroma_db.py:344-375, faiss_db.py:374-393);
* the 24 cubic operators ``CUBIC_SYMMETRY``              (latice/utils/constants.py:13-38), which
  scipy parses scalar-LAST, i.e. each listed row is (x, y, z, w).

scipy conventions restated here (scipy.spatial.transform.Rotation):
``from_euler("zxz", [a, b, c], degrees=True)`` is extrinsic, R = Rz(c) Rx(b) Rz(a);
``p * q`` is the Hamilton product p (x) q; ``inv`` is the conjugate;
``magnitude`` = 2 atan2(|xyz|, |w|); ``mean`` = eigenvector of the largest eigenvalue of
sum q q^T; ``as_euler("zxz")`` per ``euler_zxz_from_quat`` below.

Pinned by tests/golden/consensus.npz (outputs of the unmodified reference code on seeded cases,
incl. the reference's own known-answer test tests/index/test_chroma_db.py:306-382).
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np

_S2 = 1.0 / math.sqrt(2.0)

# Rows as written in latice/utils/constants.py:13-38, interpreted (x, y, z, w) like scipy does.
CUBIC_XYZW = np.array(
    [
        [1, 0, 0, 0],
        [0, 1, 0, 0],
        [0, 0, 1, 0],
        [0, 0, 0, 1],
        [0.5, 0.5, 0.5, 0.5],
        [0.5, -0.5, -0.5, -0.5],
        [0.5, 0.5, -0.5, 0.5],
        [0.5, -0.5, 0.5, -0.5],
        [0.5, -0.5, 0.5, 0.5],
        [0.5, 0.5, -0.5, -0.5],
        [0.5, -0.5, -0.5, 0.5],
        [0.5, 0.5, 0.5, -0.5],
        [_S2, _S2, 0, 0],
        [_S2, 0, _S2, 0],
        [_S2, 0, 0, _S2],
        [_S2, -_S2, 0, 0],
        [_S2, 0, -_S2, 0],
        [_S2, 0, 0, -_S2],
        [0, _S2, _S2, 0],
        [0, -_S2, _S2, 0],
        [0, 0, _S2, _S2],
        [0, 0, -_S2, _S2],
        [0, _S2, 0, _S2],
        [0, -_S2, 0, _S2],
    ],
    dtype=np.float64,
)
# scipy's from_quat normalises its input; the table rows are unit already up to rounding.
CUBIC_XYZW = CUBIC_XYZW / np.linalg.norm(CUBIC_XYZW, axis=1, keepdims=True)


def quat_from_euler_zxz_deg(eulers_deg: np.ndarray) -> np.ndarray:
    """[...,3] (phi1, Phi, phi2) degrees -> [...,4] (x, y, z, w); extrinsic zxz."""
    e = np.deg2rad(np.asarray(eulers_deg, dtype=np.float64))
    a, b, c = e[..., 0], e[..., 1], e[..., 2]
    hb = 0.5 * b
    hp = 0.5 * (a + c)
    hm = 0.5 * (a - c)
    sb, cb = np.sin(hb), np.cos(hb)
    return np.stack([sb * np.cos(hm), -sb * np.sin(hm), cb * np.sin(hp), cb * np.cos(hp)], axis=-1)


def quat_mul(p: np.ndarray, q: np.ndarray) -> np.ndarray:
    px, py, pz, pw = p[..., 0], p[..., 1], p[..., 2], p[..., 3]
    qx, qy, qz, qw = q[..., 0], q[..., 1], q[..., 2], q[..., 3]
    return np.stack(
        [
            pw * qx + px * qw + py * qz - pz * qy,
            pw * qy - px * qz + py * qw + pz * qx,
            pw * qz + px * qy - py * qx + pz * qw,
            pw * qw - px * qx - py * qy - pz * qz,
        ],
        axis=-1,
    )


def quat_conj(q: np.ndarray) -> np.ndarray:
    return q * np.array([-1.0, -1.0, -1.0, 1.0])


def quat_angle(q: np.ndarray) -> np.ndarray:
    return 2.0 * np.arctan2(np.sqrt(q[..., 0] ** 2 + q[..., 1] ** 2 + q[..., 2] ** 2), np.abs(q[..., 3]))


def euler_zxz_from_quat(q: np.ndarray) -> np.ndarray:
    """(x,y,z,w) -> extrinsic zxz Euler angles in degrees, as scipy's ``as_euler("zxz", degrees=True)``."""
    x, y, z, w = (float(v) for v in q)
    half_sum = math.atan2(z, w)
    half_diff = math.atan2(-y, x)
    big = 2.0 * math.atan2(math.hypot(x, y), math.hypot(w, z))
    eps = 1e-7
    if abs(big) <= eps:  # gimbal lock, Phi = 0: third angle set to zero
        first, third = 2.0 * half_sum, 0.0
    elif abs(big - math.pi) <= eps:  # gimbal lock, Phi = pi
        first, third = 2.0 * half_diff, 0.0
    else:
        first, third = half_sum + half_diff, half_sum - half_diff
    out = []
    for ang in (first, big, third):
        if ang < -math.pi:
            ang += 2.0 * math.pi
        elif ang > math.pi:
            ang -= 2.0 * math.pi
        out.append(math.degrees(ang))
    return np.array(out, dtype=np.float64)


def chordal_mean(quats: np.ndarray) -> np.ndarray:
    m = quats.T @ quats
    vals, vecs = np.linalg.eigh(m)
    return vecs[:, -1]


@dataclass
class ConsensusResult:
    success: bool
    similar_indices: np.ndarray | None
    mean_quat: np.ndarray | None
    mean_orientation: np.ndarray | None
    best_orientation: np.ndarray
    ref_iteration: int


def symmetry_reduce(ref: np.ndarray, cand: np.ndarray, mode: str) -> np.ndarray:
    """Return the symmetry-equivalent of ``cand`` closest to ``ref`` as a quaternion."""
    if mode == "chroma":
        # all_j = cand^-1 * S_j ; j* = argmin angle(ref * all_j) ; result = all_j*^-1   (chroma_db.py:365-373)
        allq = quat_mul(quat_conj(cand)[None, :], CUBIC_XYZW)
        j = int(np.argmin(quat_angle(quat_mul(ref[None, :], allq))))
        return quat_conj(allq[j])
    # all_j = S_j * cand ; j* = argmin angle(ref^-1 * all_j) ; result = all_j*       (faiss_db.py:388-393)
    allq = quat_mul(CUBIC_XYZW, cand[None, :])
    j = int(np.argmin(quat_angle(quat_mul(quat_conj(ref)[None, :], allq))))
    return allq[j]


def find_best_orientation(
    cand_eulers_deg: np.ndarray,
    orientation_threshold: float = 1.0,
    min_required_matches: int = 18,
    max_iterations: int = 3,
    mode: str = "chroma",
) -> ConsensusResult:
    cand_eulers_deg = np.asarray(cand_eulers_deg, dtype=np.float64).reshape(-1, 3)
    k = len(cand_eulers_deg)
    quats = quat_from_euler_zxz_deg(cand_eulers_deg)
    if mode == "chroma":
        iterations = max_iterations
        if k < max_iterations:
            # the reference indexes orientations[iteration] unguarded (chroma_db.py:302-303)
            raise IndexError("top_n smaller than max_iterations")
    elif mode == "faiss":
        iterations = min(max_iterations, k)
    else:
        raise ValueError(mode)

    similar = None
    ref_it = -1
    for it in range(iterations):
        ref = quats[it]
        ref_it = it
        if mode == "chroma":
            ang = quat_angle(quat_mul(ref[None, :], quat_conj(quats)))
        else:
            ang = np.degrees(quat_angle(quat_mul(quat_conj(ref)[None, :], quats)))
        similar = np.where(ang < orientation_threshold)[0]
        if len(similar) >= min_required_matches:
            reduced = np.stack([symmetry_reduce(ref, quats[i], mode) for i in similar]) if len(similar) else None
            if reduced is None:
                # faiss twin: an empty similar set that still passes (min_required_matches <= 0)
                return ConsensusResult(True, similar, None, None, cand_eulers_deg[0], it)
            mq = chordal_mean(reduced)
            mean_e = euler_zxz_from_quat(mq)
            best = mean_e if mode == "faiss" else cand_eulers_deg[0]
            return ConsensusResult(True, similar, mq, mean_e, best, it)
    return ConsensusResult(False, similar, None, None, cand_eulers_deg[0], ref_it)
