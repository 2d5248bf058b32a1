"""Oracle: the reference's input transform and angle-file parser, restated with plain loops.

Test infrastructure (see oracle/__init__.py).

``create_default_transform`` (latice/data_module.py:17-33) is
ToPILImage -> Grayscale -> CenterCrop(image_size) -> ToTensor.  For a 2-D float ndarray that is:

1. ``(x * 255).astype(uint8)`` -- truncation toward zero (torchvision to_pil_image); integer
   uint8 input is taken as is.
2. Grayscale on a mode "L" image: identity.
3. CenterCrop: when the image is smaller than the target along an edge, zero-pad
   ``(target - size) // 2`` before and ``(target - size + 1) // 2`` after; then crop starting at
   ``int(round((size - target) / 2.0))`` (Python round: half to even).
4. ToTensor: uint8 -> float32 / 255, shape [1, H, W].

This restatement returns the uint8 image (step 3's output); step 4 is ``k / 255``.
Pinned by tests/golden/transform.npz.  Deliberately written as per-pixel loops so that it
shares no vectorised code with the product's host transform.
"""
from __future__ import annotations

import numpy as np


def quantise(pattern: np.ndarray) -> np.ndarray:
    if pattern.ndim != 2:
        raise ValueError("pattern must be 2-D")
    if np.issubdtype(pattern.dtype, np.floating):
        return (pattern * 255).astype(np.uint8)
    if pattern.dtype == np.uint8:
        return pattern
    raise TypeError(f"unsupported pattern dtype {pattern.dtype}")


def _axis_plan(size: int, target: int) -> tuple[int, int]:
    """Return (pad_before, crop_start) along one axis."""
    pad_before = (target - size) // 2 if target > size else 0
    pad_after = (target - size + 1) // 2 if target > size else 0
    padded = size + pad_before + pad_after
    crop_start = int(round((padded - target) / 2.0)) if padded != target else 0
    return pad_before, crop_start


def centre_crop_u8(img: np.ndarray, image_size: tuple[int, int]) -> np.ndarray:
    th, tw = image_size
    h, w = img.shape
    pad_top, top = _axis_plan(h, th)
    pad_left, left = _axis_plan(w, tw)
    out = np.zeros((th, tw), dtype=np.uint8)
    for oy in range(th):
        sy = oy + top - pad_top
        if sy < 0 or sy >= h:
            continue
        for ox in range(tw):
            sx = ox + left - pad_left
            if 0 <= sx < w:
                out[oy, ox] = img[sy, sx]
    return out


def transform_u8(pattern: np.ndarray, image_size: tuple[int, int] = (128, 128)) -> np.ndarray:
    return centre_crop_u8(quantise(pattern), image_size)


def parse_rotation_angles(path) -> np.ndarray:
    """Restates DPdataset._parse_rotation_angles (latice/data_module.py:87-116): skip two header
    lines, split on single spaces dropping empties, three float64 columns (z1, x, z2) in degrees."""
    rows = []
    with open(path) as fh:
        for lineno, line in enumerate(fh):
            if lineno < 2:
                continue
            parts = [tok for tok in line.strip().split(" ") if tok]
            rows.append(parts)
    for r in rows:
        if len(r) != 3:
            raise ValueError(f"Failed to parse rotation angles file: expected 3 columns, got {len(r)}")
    return np.array([[float(t) for t in r] for r in rows], dtype=np.float64).reshape(-1, 3)
