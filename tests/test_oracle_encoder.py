"""The encoder oracle (oracle/encoder_ref.py) against the reference's own outputs (tests/golden/encoder_seed42.npz)."""
import os

import numpy as np
import torch

from oracle import encoder_ref as E


def _golden(golden_dir):
    return np.load(os.path.join(golden_dir, "encoder_seed42.npz"))


def test_weight_generator_matches_reference_state_dict(golden_dir):
    g = _golden(golden_dir)
    sd = E.make_state_dict(42)
    assert list(g["keys"]) == list(E.HOT_KEYS)
    sums = np.array([sd[k].double().sum().item() for k in E.HOT_KEYS])
    abs_sums = np.array([sd[k].double().abs().sum().item() for k in E.HOT_KEYS])
    np.testing.assert_array_equal(sums, g["weight_sums"])
    np.testing.assert_array_equal(abs_sums, g["weight_abs_sums"])


def test_synthetic_patterns_are_reproducible(golden_dir):
    g = _golden(golden_dir)
    np.testing.assert_array_equal(E.synthetic_patterns(4, seed=1234).numpy(), g["patterns"])


def test_encoder_oracle_matches_reference_mu_logvar(golden_dir):
    g = _golden(golden_dir)
    sd = E.make_state_dict(42)
    mu, logvar = E.encode(sd, E.u8_to_input(torch.from_numpy(g["patterns"])))
    # same ops in the same order on the same torch build: expect (near) bit equality
    np.testing.assert_allclose(mu.numpy(), g["mu"], rtol=0, atol=1e-6)
    np.testing.assert_allclose(logvar.numpy(), g["logvar"], rtol=0, atol=1e-6)


def test_fp32_oracle_close_to_fp64(golden_dir):
    g = _golden(golden_dir)
    sd = E.make_state_dict(42)
    p = torch.from_numpy(g["patterns"])
    mu32, _ = E.encode(sd, E.u8_to_input(p))
    mu64, _ = E.encode(sd, E.u8_to_input(p, torch.float64))
    rel = (mu32.double() - mu64).norm(dim=1) / mu64.norm(dim=1)
    assert rel.max().item() < 5e-5


def test_conv_bias_is_cancelled_by_instance_norm():
    """SURVEY finding 2: InstanceNorm(affine=False) removes the per-channel conv bias."""
    sd = E.make_state_dict(7)
    nob = {k: (torch.zeros_like(v) if k.startswith("encoder") and k.endswith("bias") else v) for k, v in sd.items()}
    x = E.u8_to_input(E.synthetic_patterns(2, seed=3), torch.float64)
    a, _ = E.encode(sd, x)
    b, _ = E.encode(nob, x)
    assert (a - b).abs().max().item() < 1e-9
