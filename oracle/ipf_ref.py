"""Oracle: IPF colour key of an orientation (float64 numpy, plain loops).

Test infrastructure (see oracle/__init__.py).  Restates

* ``get_color_key``                         (latice/utils/utils.py:206-240): ``R.from_euler("zxz", angles, degrees=True)
  .as_matrix()``, pole = ROW 0 / 1 / 2 of the matrix for ``ipf_x`` / ``ipf_y`` / ``ipf_z``;
* ``ColorKeyGenerator.generate_ipf_color``  (latice/utils/colorkey.py:64-130): normalise the pole, apply the 24 cubic
  operators (``QUAT_SYM.as_matrix() @ v``) and append the negated vectors (48 candidates), take the FIRST candidate
  (flipped to z >= 0, ``USE_INVERSION``) whose (eta, chi) = (atan2(y, x), acos(z)) lies in the unit triangle
  0 <= eta <= 45 deg, 0 <= chi <= acos(1/sqrt 3); if none does, the angles of the LAST candidate are used; then
  rgb = (1 - c, (1 - e) c, e c) with c = chi/chi_max, e = eta/45 deg, square root, scale so that max = 255, round.

scipy conventions restated: extrinsic zxz quaternion (x, y, z, w) as in oracle/consensus_ref.py; ``as_matrix`` of a
unit quaternion.  Pinned by tests/golden/ipf.npz (outputs of the unmodified reference functions).
"""
from __future__ import annotations

import math

import numpy as np

from .consensus_ref import CUBIC_XYZW, quat_from_euler_zxz_deg


def quat_to_matrix(q) -> np.ndarray:
    """scipy Rotation.as_matrix for a unit quaternion (x, y, z, w)."""
    x, y, z, w = (float(v) for v in q)
    x2, y2, z2, w2 = x * x, y * y, z * z, w * w
    xy, zw, xz, yw, yz, xw = x * y, z * w, x * z, y * w, y * z, x * w
    return np.array([
        [x2 - y2 - z2 + w2, 2 * (xy - zw), 2 * (xz + yw)],
        [2 * (xy + zw), -x2 + y2 - z2 + w2, 2 * (yz - xw)],
        [2 * (xz - yw), 2 * (yz + xw), -x2 - y2 + z2 + w2],
    ])


CUBIC_MATRICES = np.stack([quat_to_matrix(q) for q in CUBIC_XYZW])
_ETA_MAX = 45.0 * (math.pi / 180)
_CHI_MAX = math.acos(1 / math.sqrt(3))


def ipf_color_of_pole(pole) -> list[int]:
    v = np.asarray(pole, dtype=np.float64)
    v = v / math.sqrt(float(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]))
    cands = [CUBIC_MATRICES[j] @ v for j in range(24)]
    cands = cands + [-c for c in cands]
    chi = eta = 0.0
    for c in cands:
        if c[2] < 0:
            c = -c
        chi = math.acos(min(1.0, max(-1.0, float(c[2]))))
        eta = math.atan2(float(c[1]), float(c[0]))
        if not (eta < 0 or eta > _ETA_MAX or chi < 0 or chi > _CHI_MAX):
            break
    chi_max = _CHI_MAX * (180 / math.pi)
    eta_deg, chi_deg = eta * (180 / math.pi), chi * (180 / math.pi)
    rgb = [1 - chi_deg / chi_max, 0.0, abs(eta_deg - 0) / (45 - 0)]
    rgb[1] = 1 - rgb[2]
    rgb[1] *= chi_deg / chi_max
    rgb[2] *= chi_deg / chi_max
    rgb = [math.sqrt(val) for val in rgb]
    m = max(rgb)
    return [int(round(255 * val / m)) for val in rgb]


def get_color_key(rot_angle, mode: str = "ipf_z") -> np.ndarray:
    """[n,3] ZXZ Euler angles in degrees -> uint8-valued int array [n,3]."""
    rot_angle = np.asarray(rot_angle, dtype=np.float64)
    if rot_angle.ndim < 2:
        rot_angle = rot_angle[None]
    row = {"ipf_x": 0, "ipf_y": 1, "ipf_z": 2}[mode]
    out = np.zeros((len(rot_angle), 3), dtype=np.int64)
    for i, e in enumerate(rot_angle):
        out[i] = ipf_color_of_pole(quat_to_matrix(quat_from_euler_zxz_deg(e[None])[0])[row])
    return out
