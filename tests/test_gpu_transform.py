"""GPU parity of the device-side input transform (csrc/transform.cu, ebsd_quantize_crop) -- bit-exact against the
oracle restatement of create_default_transform (oracle/transform_ref.py, latice/data_module.py:17-33), on the golden
fixture produced by the unmodified reference and on seeded inputs incl. out-of-range values, odd sizes and padding."""
import numpy as np
import pytest
import torch

from oracle import transform_ref

pytestmark = pytest.mark.gpu


def _device(frames: np.ndarray) -> np.ndarray:
    from ebsd_vae_b200.transform import transform_batch_device
    out = transform_batch_device(torch.from_numpy(np.ascontiguousarray(frames)).cuda())
    torch.cuda.synchronize()
    return out.cpu().numpy()


@pytest.mark.parametrize("shape", [(3, 128, 128), (2, 150, 131), (2, 129, 128), (2, 100, 90), (1, 127, 200), (4, 64, 64)])
@pytest.mark.parametrize("dtype", [np.float64, np.float32, np.uint8])
def test_device_transform_matches_oracle(shape, dtype):
    rng = np.random.default_rng(hash((shape, np.dtype(dtype).name)) % (2**32))
    if dtype == np.uint8:
        frames = rng.integers(0, 256, size=shape, dtype=np.uint8)
    else:
        frames = rng.uniform(-0.2, 1.3, size=shape).astype(dtype)      # includes values that wrap modulo 256
        frames.flat[:: 97] = (rng.uniform(0, 255, size=frames.flat[:: 97].shape)).astype(dtype)   # 0..255-valued data
        frames.flat[5] = np.nan
        frames.flat[6] = np.inf
        frames.flat[7] = -1e12
        frames.flat[8] = 8421504.0      # 2^31 / 255 .. boundary of the 32-bit truncation
        frames.flat[9] = 8421505.0
    with np.errstate(all="ignore"):
        want = np.stack([transform_ref.transform_u8(f) for f in frames])
    got = _device(frames)
    assert got.dtype == np.uint8 and got.shape == (shape[0], 128, 128)
    np.testing.assert_array_equal(got, want)


def test_device_transform_matches_reference_golden(golden_dir):
    """The fixture holds the outputs of the unmodified reference transform for oracle.make_golden.TRANSFORM_CASES."""
    import os

    from oracle.make_golden import TRANSFORM_CASES, transform_input
    g = np.load(os.path.join(golden_dir, "transform.npz"))
    for case, want in zip(TRANSFORM_CASES, g["outputs"]):
        frame = transform_input(*case)
        if frame.dtype not in (np.float64, np.float32, np.uint8):
            continue
        got = _device(frame[None])[0]
        np.testing.assert_array_equal(got, want, err_msg=str(case))


def test_encode_batch_uses_device_transform_and_matches_host_path():
    """encode_patterns_batch(float ndarray) = encoder(host-transformed uint8): same latents bit for bit."""
    import ebsd_vae_b200 as E
    from ebsd_vae_b200.transform import transform_batch_u8
    from oracle import encoder_ref as R

    model = E.VariationalAutoEncoderRawData()
    model.load_state_dict(R.make_state_dict(42))
    indexer = E.DiffractionPatternIndexer(model, config=E.IndexerConfig(device="cuda"))
    rng = np.random.default_rng(5)
    pats = rng.uniform(0, 1, size=(6, 140, 133))
    a = indexer.encode_patterns_batch(pats)
    b = indexer.engine.encode(torch.from_numpy(transform_batch_u8(pats)).cuda()).cpu().numpy()
    np.testing.assert_array_equal(a, b)
