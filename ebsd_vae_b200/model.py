"""Encoder weights container and the CUDA encoder engine.

``VariationalAutoEncoderRawData`` keeps the reference's constructor signature and state_dict layout
(latice/model.py:83-150) for the part of the network the indexer uses -- the ten encoder convolutions and the
mu / logvar heads -- so ``model.load_state_dict(torch.load("vae-best.pt"))`` works as in the reference README
(README.md:90-93).  The decoder, ``linear2`` and the reparameterisation (latice/model.py:59-64, 131-150) are not
part of the indexing path (the indexer discards them, latice/index/dp_indexer.py:136,183,284); their keys are
accepted and ignored when loading.

``EncoderEngine`` owns the native ``ebsd_encoder`` handle (packed weights on the device) and runs the
hand-written sm_100a kernels.  There is no PyTorch implementation of the forward pass in this package.
"""
from __future__ import annotations

import ctypes
from collections import OrderedDict
from typing import Mapping

import numpy as np
import torch
import torch.nn as nn

from . import _native

# (index inside ``encoder``, Cin, Cout) -- MaxPool2d modules sit at indices 2, 5, 8, 11, 14 (latice/model.py:109-125)
CONV_PLAN = ((0, 1, 32), (1, 32, 32), (3, 32, 64), (4, 64, 64), (6, 64, 128), (7, 128, 128), (9, 128, 128),
             (10, 128, 128), (12, 128, 128), (13, 128, 128))
LATENT_DIM = 16
FLAT_DIM = 2048


def _hot_keys() -> list[str]:
    keys = []
    for idx, _, _ in CONV_PLAN:
        keys += [f"encoder.{idx}.0.weight", f"encoder.{idx}.0.bias"]
    for head in ("mu", "logvar"):
        keys += [f"{head}.0.weight", f"{head}.0.bias"]
    return keys


HOT_KEYS = tuple(_hot_keys())


def extract_hot_state_dict(state: Mapping[str, torch.Tensor]) -> "OrderedDict[str, torch.Tensor]":
    """Pick the encoder + head tensors out of a ``vae-best.pt`` style mapping.

    Accepts a plain state_dict, a Lightning checkpoint (``{"state_dict": ...}``) and key prefixes such as
    ``model.``.  Raises KeyError naming the first missing tensor.
    """
    if "state_dict" in state and isinstance(state["state_dict"], Mapping):
        state = state["state_dict"]
    prefix = None
    for cand in ("", "model.", "module.", "vae.", "model.model."):
        if cand + HOT_KEYS[0] in state:
            prefix = cand
            break
    if prefix is None:
        raise KeyError(f"'{HOT_KEYS[0]}' not found in state dict (keys start with {list(state)[:3]})")
    out = OrderedDict()
    for key in HOT_KEYS:
        if prefix + key not in state:
            raise KeyError(f"missing weight '{prefix + key}'")
        out[key] = state[prefix + key]
    expected = {f"encoder.{i}.0.weight": (co, ci, 3, 3) for i, ci, co in CONV_PLAN}
    expected.update({"mu.0.weight": (LATENT_DIM, FLAT_DIM), "logvar.0.weight": (LATENT_DIM, FLAT_DIM)})
    for key, shape in expected.items():
        if tuple(out[key].shape) != shape:
            raise ValueError(f"weight '{key}' has shape {tuple(out[key].shape)}, expected {shape}")
    return out


def load_vae_weights(path, map_location="cpu") -> "OrderedDict[str, torch.Tensor]":
    """``torch.load`` a ``vae-best.pt`` style file and return the tensors the indexing path needs."""
    state = torch.load(path, map_location=map_location, weights_only=True)
    return extract_hot_state_dict(state)


class EncoderEngine:
    """Native encoder: uint8/float32 patterns on the GPU -> (mu, logvar) on the GPU."""

    def __init__(self, state: Mapping[str, torch.Tensor], device: torch.device | str = "cuda") -> None:
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError("ebsd_vae_b200 runs on CUDA devices only (sm_100a); there is no CPU encoder")
        if device.index is None:
            device = torch.device("cuda", torch.cuda.current_device())
        self.device = device
        self._lib = _native.load()
        hot = extract_hot_state_dict(state)
        with torch.cuda.device(device):
            tensors = {k: v.detach().to(device=device, dtype=torch.float32).contiguous() for k, v in hot.items()}
            w = _native.EbsdWeights()
            for i, (idx, _, _) in enumerate(CONV_PLAN):
                w.conv_w[i] = tensors[f"encoder.{idx}.0.weight"].data_ptr()
                w.conv_b[i] = tensors[f"encoder.{idx}.0.bias"].data_ptr()
            w.mu_w = tensors["mu.0.weight"].data_ptr()
            w.mu_b = tensors["mu.0.bias"].data_ptr()
            w.logvar_w = tensors["logvar.0.weight"].data_ptr()
            w.logvar_b = tensors["logvar.0.bias"].data_ptr()
            handle = ctypes.c_void_p()
            stream = torch.cuda.current_stream(device).cuda_stream
            _native.check(
                self._lib.ebsd_encoder_create(ctypes.byref(handle), ctypes.byref(w), device.index, stream),
                "ebsd_encoder_create",
            )
        self._handle = handle
        self._workspace: torch.Tensor | None = None

    def close(self) -> None:
        if getattr(self, "_handle", None):
            self._lib.ebsd_encoder_destroy(self._handle)
            self._handle = None

    def __del__(self):  # pragma: no cover - best effort
        try:
            self.close()
        except Exception:
            pass

    def _get_workspace(self, batch: int) -> torch.Tensor:
        need = int(self._lib.ebsd_encoder_workspace_bytes(self._handle, batch))
        if self._workspace is None or self._workspace.numel() < need:
            self._workspace = torch.empty(need, dtype=torch.uint8, device=self.device)
        return self._workspace

    def encode(self, patterns: torch.Tensor, want_logvar: bool = False):
        """patterns: CUDA tensor [B,128,128] (uint8 = k/255 encoded, or float32 used as is). Returns mu[, logvar]."""
        if patterns.device != self.device:
            raise ValueError(f"patterns live on {patterns.device}, the encoder on {self.device}")
        if patterns.dim() == 4 and patterns.shape[1] == 1:
            patterns = patterns[:, 0]
        if patterns.dim() != 3 or tuple(patterns.shape[1:]) != (128, 128):
            raise ValueError(f"expected patterns of shape [B,128,128], got {tuple(patterns.shape)}")
        if patterns.dtype == torch.uint8:
            dtype = _native.PATTERN_U8
        elif patterns.dtype == torch.float32:
            dtype = _native.PATTERN_F32
        else:
            raise TypeError(f"patterns must be uint8 or float32, got {patterns.dtype}")
        patterns = patterns.contiguous()
        b = patterns.shape[0]
        mu = torch.empty((b, LATENT_DIM), dtype=torch.float32, device=self.device)
        logvar = torch.empty_like(mu) if want_logvar else None
        if b:
            with torch.cuda.device(self.device):
                ws = self._get_workspace(b)
                stream = torch.cuda.current_stream(self.device).cuda_stream
                _native.check(
                    self._lib.ebsd_encoder_forward(
                        self._handle, patterns.data_ptr(), dtype, b, mu.data_ptr(),
                        logvar.data_ptr() if logvar is not None else None, ws.data_ptr(), ws.numel(), stream),
                    "ebsd_encoder_forward",
                )
        return (mu, logvar) if want_logvar else mu


def _block(cin: int, cout: int) -> nn.Sequential:
    return nn.Sequential(nn.Conv2d(cin, cout, 3, stride=1, padding=1), nn.InstanceNorm2d(cout), nn.LeakyReLU(0.02))


class VariationalAutoEncoderRawData(nn.Module):
    """Weights container with the reference's layout; ``forward`` runs the native encoder.

    ``forward(x)`` returns ``(None, None, mu, std)``: the indexing path only consumes ``mu``
    (latice/index/dp_indexer.py:136), and the sampled ``z`` / reconstruction are not computed.
    """

    def __init__(self, inplanes: int = 32, latent_dim: int = 16):
        super().__init__()
        if inplanes != 32 or latent_dim != LATENT_DIM:
            raise ValueError("the B200 kernels are specialised for inplanes=32, latent_dim=16 (the reference defaults)")
        layers: list[nn.Module] = []
        for n, (idx, cin, cout) in enumerate(CONV_PLAN):
            layers.append(_block(cin, cout))
            if n % 2 == 1:
                layers.append(nn.MaxPool2d(2, 2))
        self.encoder = nn.Sequential(*layers)
        self.mu = nn.Sequential(nn.Linear(FLAT_DIM, latent_dim))
        self.logvar = nn.Sequential(nn.Linear(FLAT_DIM, latent_dim))
        self._engine: EncoderEngine | None = None

    def load_state_dict(self, state_dict, strict: bool = True, assign: bool = False):
        """Loads ``vae-best.pt``; decoder / linear2 keys of the full reference model are ignored."""
        hot = extract_hot_state_dict(state_dict)
        self._engine = None
        return super().load_state_dict(hot, strict=True, assign=assign)

    def _apply(self, fn, *args, **kwargs):
        self._engine = None  # device or dtype may have changed
        return super()._apply(fn, *args, **kwargs)

    def engine(self) -> EncoderEngine:
        device = next(self.parameters()).device
        if self._engine is None or self._engine.device != device:
            self._engine = EncoderEngine(self.state_dict(), device)
        return self._engine

    @torch.no_grad()
    def forward(self, x: torch.Tensor):
        if x.dim() == 4:
            x = x[:, 0]
        mu, logvar = self.engine().encode(x.contiguous(), want_logvar=True)
        return None, None, mu, torch.exp(logvar / 2)


def numpy_state_dict(state: Mapping[str, torch.Tensor]) -> dict[str, np.ndarray]:
    return {k: v.detach().cpu().numpy() for k, v in extract_hot_state_dict(state).items()}
