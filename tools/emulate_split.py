"""CPU emulation of the encoder's tensor-core arithmetic (csrc/encoder_fused.cuh, "Arithmetic"): how far the latents
move from float64 when every convolution of blocks 1..9 is computed as

    f16x1 : one fp16 pass                                   a_hi*w_hi
    f16x3 : the exact three-term fp16 split                 a_hi*w_hi + a_hi*w_lo + a_lo*w_hi          (round 1)
    f8corr: one fp16 pass + one e4m3 pass of corrections    a_hi*w_hi + e4m3(a)*e4m3(w_lo) + e4m3(a_lo)*e4m3(w)   (this build)

with the scales the kernels use (residuals x 4096, weights x the power of two that brings max|w| into (64, 128]).
Accumulation is float64 here (TMEM accumulates in fp32; the f16x3 row shows that this does not matter at 1e-6).

    python tools/emulate_split.py [n_patterns] [weight_seed]
"""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import encoder_ref as R  # noqa: E402

RES_SCALE = 4096.0


def f16(x):
    return x.to(torch.float16).to(torch.float64)


def e4m3(x):
    return x.clamp(-448, 448).to(torch.float32).to(torch.float8_e4m3fn).to(torch.float64)


def conv(a, w):
    return F.conv2d(a, w, None, 1, 1)


def layer(a, w, mode):
    a = a.to(torch.float32).to(torch.float64)   # activations are fp32 values in the kernels
    w = w.to(torch.float64)
    if mode == "exact":
        return conv(a, w)
    a_hi, w_hi = f16(a), f16(w)
    a_lo, w_lo = a - a_hi, w - w_hi
    if mode == "f16x1":
        return conv(a_hi, w_hi)
    if mode == "f16x3":
        return conv(a_hi, w_hi) + conv(a_hi, f16(w_lo)) + conv(f16(a_lo), w_hi)
    if mode == "f8corr":
        s = 2.0 ** np.floor(np.log2(128.0 / w.abs().max().item()))
        corr = conv(e4m3(a), e4m3(w_lo * (RES_SCALE * s))) + conv(e4m3(a_lo * RES_SCALE), e4m3(w * s))
        return conv(a_hi, w_hi) + corr / (RES_SCALE * s)
    raise ValueError(mode)


@torch.no_grad()
def encode(sd, x, mode):
    x = x.to(torch.float64)
    for bi, (idx, _, _, pooled) in enumerate(R.ENCODER_PLAN):
        y = layer(x, sd[f"encoder.{idx}.0.weight"], "exact" if bi == 0 else mode)   # conv0 runs in fp32 on CUDA cores
        y = F.leaky_relu(F.instance_norm(y, eps=R.IN_EPS), R.LEAKY_SLOPE)
        x = F.max_pool2d(y, 2, 2) if pooled else y
    return F.linear(x.flatten(1), sd["mu.0.weight"].double(), sd["mu.0.bias"].double())


def rel(a, b):
    return torch.linalg.norm(a - b, dim=1) / torch.linalg.norm(b, dim=1)


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 42
    sd = R.make_state_dict(seed)
    x = R.u8_to_input(R.synthetic_patterns(n, seed=5))
    ref = encode(sd, x, "exact")
    t32, _ = R.encode(sd, x)
    print(f"{n} patterns, weight seed {seed}; relative L2 error of mu against float64")
    print(f"  torch fp32           max {rel(t32.double(), ref).max().item():.3e}")
    for mode in ("f16x3", "f8corr", "f16x1"):
        r = rel(encode(sd, x, mode), ref)
        print(f"  {mode:20s} max {r.max().item():.3e}  median {r.median().item():.3e}")
