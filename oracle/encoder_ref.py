"""Oracle: VAE encoder + mu/logvar heads, restated with torch functional ops on CPU.

Test infrastructure (see oracle/__init__.py).  Follows the reference
``VariationalAutoEncoderRawData`` (latice/model.py:83-150):

* encoder block = Conv2d(3x3, stride 1, pad 1) -> InstanceNorm2d(affine=False, eps=1e-5,
  biased variance over H*W) -> LeakyReLU(0.02)               (latice/model.py:93-98)
* ten blocks, channel plan 1->32->32->64->64->128->128->128->128->128->128, MaxPool2d(2,2)
  after blocks 1,3,5,7,9 (0-based)                           (latice/model.py:109-125)
* mu = Linear(2048,16)(enc.flatten(1)), logvar likewise      (latice/model.py:57-58,127-129)
* everything after logvar (rsample, linear2, decoder; latice/model.py:59-64) is discarded by
  the indexer (latice/index/dp_indexer.py:136,183,284) and is not restated.

Pinned against the unmodified reference by tests/golden/encoder_seed42.npz
(made by oracle/make_golden.py).
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

# (state_dict index inside ``encoder``, Cin, Cout, pooled afterwards)
ENCODER_PLAN = (
    (0, 1, 32, False),
    (1, 32, 32, True),
    (3, 32, 64, False),
    (4, 64, 64, True),
    (6, 64, 128, False),
    (7, 128, 128, True),
    (9, 128, 128, False),
    (10, 128, 128, True),
    (12, 128, 128, False),
    (13, 128, 128, True),
)
LATENT_DIM = 16
FLAT_DIM = 128 * 4 * 4
LEAKY_SLOPE = 0.02
IN_EPS = 1e-5

HOT_KEYS = tuple(
    [f"encoder.{i}.0.{p}" for i, _, _, _ in ENCODER_PLAN for p in ("weight", "bias")]
    + [f"{h}.0.{p}" for h in ("mu", "logvar") for p in ("weight", "bias")]
)


def make_state_dict(seed: int = 42) -> dict[str, torch.Tensor]:
    """Random-init weights in ``vae-best.pt`` format (hot keys only).

    Builds the parameterised layers in the order the reference constructor does
    (encoder convs, then mu, then logvar; latice/model.py:109-129) after
    ``torch.manual_seed(seed)``, so the tensors equal those of
    ``torch.manual_seed(seed); VariationalAutoEncoderRawData().state_dict()``.
    (The reference's ``self.apply(self.weights_init)`` runs before any sub-module exists,
    latice/model.py:16, so torch's default Kaiming-uniform init is what remains.)
    """
    torch.manual_seed(seed)
    sd: dict[str, torch.Tensor] = {}
    for idx, cin, cout, _ in ENCODER_PLAN:
        conv = nn.Conv2d(cin, cout, 3, stride=1, padding=1)
        sd[f"encoder.{idx}.0.weight"] = conv.weight.detach().clone()
        sd[f"encoder.{idx}.0.bias"] = conv.bias.detach().clone()
    for head in ("mu", "logvar"):
        lin = nn.Linear(FLAT_DIM, LATENT_DIM)
        sd[f"{head}.0.weight"] = lin.weight.detach().clone()
        sd[f"{head}.0.bias"] = lin.bias.detach().clone()
    return sd


def encoder_features(sd: dict[str, torch.Tensor], x: torch.Tensor, n_blocks: int = 10) -> torch.Tensor:
    """Run the first ``n_blocks`` encoder blocks (pools included). x: [B,1,128,128], values k/255."""
    dtype = x.dtype
    for bi, (idx, _, _, pooled) in enumerate(ENCODER_PLAN[:n_blocks]):
        w = sd[f"encoder.{idx}.0.weight"].to(dtype)
        b = sd[f"encoder.{idx}.0.bias"].to(dtype)
        x = F.conv2d(x, w, b, stride=1, padding=1)
        x = F.instance_norm(x, eps=IN_EPS)
        x = F.leaky_relu(x, LEAKY_SLOPE)
        if pooled:
            x = F.max_pool2d(x, 2, 2)
    return x


@torch.no_grad()
def encode(sd: dict[str, torch.Tensor], x: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
    """Return (mu, logvar), each [B,16], in the dtype of ``x`` (float32 or float64)."""
    feat = encoder_features(sd, x).flatten(1)
    dtype = x.dtype
    mu = F.linear(feat, sd["mu.0.weight"].to(dtype), sd["mu.0.bias"].to(dtype))
    logvar = F.linear(feat, sd["logvar.0.weight"].to(dtype), sd["logvar.0.bias"].to(dtype))
    return mu, logvar


def u8_to_input(patterns_u8: torch.Tensor, dtype=torch.float32) -> torch.Tensor:
    """uint8 [B,128,128] -> [B,1,128,128] with values k/255 (ToTensor, latice/data_module.py:31)."""
    return (patterns_u8.to(torch.float32) / 255.0).to(dtype).unsqueeze(1)


def synthetic_patterns(n: int, seed: int = 1234, size: int = 128) -> torch.Tensor:
    """Seeded, low-pass-filtered uint8 patterns [n,size,size] (Kikuchi-band-free stand-ins)."""
    g = torch.Generator().manual_seed(seed)
    coarse = torch.rand((n, 1, size // 8 + 2, size // 8 + 2), generator=g)
    img = F.interpolate(coarse, scale_factor=8, mode="bilinear", align_corners=False)
    off = (img.shape[-1] - size) // 2
    img = img[:, 0, off : off + size, off : off + size]
    img = img + 0.08 * torch.rand((n, size, size), generator=g)
    lo = img.amin(dim=(1, 2), keepdim=True)
    hi = img.amax(dim=(1, 2), keepdim=True)
    img = (img - lo) / (hi - lo)
    return (img * 255.0).to(torch.uint8)
