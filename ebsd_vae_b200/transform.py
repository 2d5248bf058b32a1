"""Host-side input contract of the indexing path: 8-bit quantise + centre-crop, angle-file parsing.

Mirrors ``create_default_transform`` (latice/data_module.py:17-33 = ToPILImage -> Grayscale ->
CenterCrop -> ToTensor) and ``DPdataset._parse_rotation_angles`` (latice/data_module.py:87-116),
vectorised over the batch.  The encoder consumes the uint8 image directly: ToTensor's ``k / 255`` is
applied inside the first convolution kernel, so 16 KiB per pattern cross PCIe instead of 64 KiB.
"""
from __future__ import annotations

from pathlib import Path

import numpy as np


def _axis_window(size: int, target: int) -> tuple[int, int, int]:
    """(source start, destination start, length) of the copy along one axis (torchvision center_crop)."""
    if size >= target:
        start = int(round((size - target) / 2.0))  # Python round: half to even, like the reference
        return start, 0, target
    return 0, (target - size) // 2, size


def quantise_u8(patterns: np.ndarray) -> np.ndarray:
    """``(x * 255).astype(uint8)`` for float input (torchvision to_pil_image), identity for uint8."""
    if np.issubdtype(patterns.dtype, np.floating):
        return (patterns * 255).astype(np.uint8)
    if patterns.dtype == np.uint8:
        return patterns
    raise TypeError(f"Input type {patterns.dtype} is not supported")  # ToPILImage rejects other dtypes


def transform_batch_u8(patterns: np.ndarray, image_size: tuple[int, int] = (128, 128)) -> np.ndarray:
    """[B,H,W] or [H,W] patterns -> contiguous uint8 [B,th,tw]; value k stands for k/255."""
    if patterns.ndim == 2:
        patterns = patterns[None]
    if patterns.ndim != 3:
        raise ValueError(f"pic should be 2/3 dimensional. Got {patterns.ndim} dimensions.")
    th, tw = image_size
    _, h, w = patterns.shape
    sy, dy, ly = _axis_window(h, th)
    sx, dx, lx = _axis_window(w, tw)
    window = patterns[:, sy : sy + ly, sx : sx + lx]
    if (ly, lx) == (th, tw):
        return np.ascontiguousarray(quantise_u8(window))
    out = np.zeros((patterns.shape[0], th, tw), dtype=np.uint8)
    out[:, dy : dy + ly, dx : dx + lx] = quantise_u8(window)
    return out


def transform_batch_device(frames, image_size: tuple[int, int] = (128, 128), via_float64: bool = False):
    """Device version of :func:`transform_batch_u8`: CUDA tensor [B,H,W] (uint8 / float32 / float64) -> uint8
    [B,128,128] on the same device, bit-identical to the host function (``ebsd_quantize_crop``).

    ``via_float64=True`` is the dictionary path: ``DPdataset.__getitem__`` casts every frame to float64 before the
    transform (latice/data_module.py:132), so integer files are scaled by 255 and wrap modulo 256; the cast happens
    per pixel inside the kernel and uint8 / int16 / uint16 / int32 / int64 / float32 / float64 frames are accepted."""
    import torch

    from . import _native

    if tuple(image_size) != (128, 128):
        raise ValueError("the B200 encoder is specialised for image_size=(128, 128)")
    if frames.dim() != 3:
        raise ValueError(f"pic should be 2/3 dimensional. Got {frames.dim()} dimensions.")
    codes = {torch.uint8: 0, torch.float32: 1, torch.float64: 2}
    if via_float64:
        codes = {k: v | 16 for k, v in {**codes, torch.int16: 3, torch.uint16: 4, torch.int32: 5, torch.int64: 6}.items()}
    if frames.dtype not in codes:
        raise TypeError(f"Input type {frames.dtype} is not supported")
    frames = frames.contiguous()
    b, h, w = frames.shape
    sy, dy, ly = _axis_window(h, 128)
    sx, dx, lx = _axis_window(w, 128)
    out = torch.empty((b, 128, 128), dtype=torch.uint8, device=frames.device)
    with torch.cuda.device(frames.device):
        _native.check(
            _native.load().ebsd_quantize_crop(frames.data_ptr(), codes[frames.dtype], b, h, w, sy, dy, ly, sx, dx, lx,
                                              out.data_ptr(), torch.cuda.current_stream(frames.device).cuda_stream),
            "ebsd_quantize_crop",
        )
    return out


def parse_rotation_angles(path: str | Path) -> np.ndarray:
    """Angle file -> float64 [N,3] (z1, x, z2 = phi1, Phi, phi2 in degrees).

    Two header lines are skipped; fields are separated by single spaces (empty fields dropped), as in
    latice/data_module.py:100-110.  Errors follow the reference: FileNotFoundError is re-raised,
    anything else becomes ``ValueError("Failed to parse rotation angles file: ...")``.
    """
    path = Path(path)
    try:
        fast = _parse_rotation_angles_native(path)
        if fast is not None:
            return fast
        with open(path) as fh:
            lines = fh.readlines()[2:]
        rows = [[tok for tok in line.strip().split(" ") if tok] for line in lines]
        # pd.DataFrame(rows, columns=[z1, x, z2]) (data_module.py:105-110): the LONGEST row must have three fields;
        # shorter or empty rows (a trailing blank line) are padded with NaN
        longest = max((len(r) for r in rows), default=3)
        if longest != 3:
            raise ValueError(f"3 columns passed, passed data had {longest} columns")
        if all(len(r) == 3 for r in rows):   # the regular case: one C-level conversion
            return np.array(rows, dtype=np.float64).reshape(-1, 3)
        out = np.full((len(rows), 3), np.nan, dtype=np.float64)
        for i, r in enumerate(rows):
            for j, tok in enumerate(r):
                out[i, j] = float(tok)
        return out
    except FileNotFoundError:
        raise
    except Exception as exc:  # noqa: BLE001 - mirrors the reference's catch-all
        raise ValueError(f"Failed to parse rotation angles file: {exc}") from exc


def _parse_rotation_angles_native(path: Path) -> np.ndarray | None:
    """The regular case (three plain decimal numbers per line) in C, without the GIL (``ebsd_parse_angle_text``);
    None when the file is anything else -- the Python restatement above then reproduces the reference's behaviour."""
    from . import _native

    with open(path, "rb") as fh:
        raw = fh.read()
    pos = 0
    for _ in range(2):   # the two header lines (readlines()[2:])
        nl = raw.find(b"\n", pos)
        if nl < 0:
            return None
        pos = nl + 1
    if b"\r" in raw[:pos].replace(b"\r\n", b""):   # a lone CR would be a line break under universal newlines
        return None
    body = raw[pos:]
    cap = body.count(b"\n") + 1
    out = np.empty((cap, 3), dtype=np.float64)
    n = _native.load().ebsd_parse_angle_text(body, len(body), out.ctypes.data, cap)
    if n < 0:
        return None
    return out[:n]


def load_patterns(path: str | Path) -> np.ndarray:
    """``np.load`` with the reference's checks (latice/data_module.py:69-78)."""
    try:
        data = np.load(Path(path), mmap_mode="r")
    except Exception as exc:  # noqa: BLE001
        raise ValueError("Only .npy data files are supported.") from exc
    if data.ndim != 3:
        raise ValueError("The input dataset should be 3D.")
    return data
