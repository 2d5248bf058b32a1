"""The transform / angle-file oracle against the reference's outputs (tests/golden/transform.npz, angles_sample.npy)."""
import os

import numpy as np

from oracle import transform_ref as T
from oracle.make_golden import TRANSFORM_CASES, transform_input


def test_transform_matches_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "transform.npz"))
    assert len(g["outputs"]) == len(TRANSFORM_CASES)
    for case, want in zip(TRANSFORM_CASES, g["outputs"]):
        got = T.transform_u8(transform_input(*case), (128, 128))
        np.testing.assert_array_equal(got, want, err_msg=str(case))


def test_angle_file_parser(golden_dir):
    want = np.load(os.path.join(golden_dir, "angles_sample.npy"))
    got = T.parse_rotation_angles(os.path.join(golden_dir, "anglefile_sample.txt"))
    assert got.dtype == np.float64 and got.shape == (625, 3)
    np.testing.assert_array_equal(got, want)
