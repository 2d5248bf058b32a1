"""IPF colour key for orientation maps, on the GPU.

``get_color_key`` keeps the signature and return types of the reference's ``latice.utils.utils.get_color_key``
(latice/utils/utils.py:206-240); the per-orientation Python loop over ``ColorKeyGenerator.generate_ipf_color``
(latice/utils/colorkey.py:64-130) is one launch of ``ebsd_ipf_color`` (csrc/consensus.cu), one thread per orientation.
"""
from __future__ import annotations

import numpy as np
import torch
from numpy.typing import NDArray

from . import _native

_AXIS = {"ipf_x": 0, "ipf_y": 1, "ipf_z": 2}


def ipf_colors_device(eulers: torch.Tensor, mode: str = "ipf_z") -> torch.Tensor:
    """CUDA tensor [n,3] of ZXZ Euler angles (degrees) -> CUDA uint8 tensor [n,3] (R, G, B)."""
    if mode not in _AXIS:
        raise ValueError(f"mode must be one of {sorted(_AXIS)}, got {mode!r}")
    if eulers.device.type != "cuda":
        raise RuntimeError("ipf_colors_device needs a CUDA tensor; ebsd_vae_b200 has no CPU path")
    e = eulers.to(torch.float64).reshape(-1, 3).contiguous()
    rgb = torch.empty((e.shape[0], 3), dtype=torch.uint8, device=e.device)
    with torch.cuda.device(e.device):
        _native.check(
            _native.load().ebsd_ipf_color(e.data_ptr(), e.shape[0], _AXIS[mode], rgb.data_ptr(),
                                          torch.cuda.current_stream(e.device).cuda_stream),
            "ebsd_ipf_color",
        )
    return rgb


def get_color_key(rot_angle: NDArray, mode: str = "ipf_z", hex_string: bool = False) -> NDArray | list[str]:
    """Generate colour keys for rotation angles (reference signature).  ``rot_angle``: (3,) or (n,3) degrees."""
    rot_angle = np.asarray(rot_angle, dtype=np.float64)
    rot_angle = rot_angle[np.newaxis, :] if rot_angle.ndim < 2 else rot_angle
    if not torch.cuda.is_available():
        raise RuntimeError("CUDA is not available; ebsd_vae_b200 has no CPU fallback")
    rgb = ipf_colors_device(torch.from_numpy(np.ascontiguousarray(rot_angle)).cuda(), mode).cpu().numpy().astype(np.int64)
    if not hex_string:
        return rgb
    return ["#{:02x}{:02x}{:02x}".format(*c) for c in rgb]
