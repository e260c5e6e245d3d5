"""The numpy oracle against the reference's own fixture and reference-generated goldens."""
import glob
import os

import numpy as np
import pytest

from oracle import scattering2d as o2


def test_oracle_reference_fixture(golden_dir):
    # the reference's tests/scattering2d/test_data_2d.npz, asserted there with allclose
    d = np.load(os.path.join(golden_dir, "ref_fixture_2d.npz"))
    y = o2.scattering2d(d["x"], int(d["J"]), pre_pad=bool(d["pre_pad"]))
    assert y.shape == d["Sx"].shape
    assert np.allclose(y, d["Sx"])
    y1 = o2.scattering2d(d["x"], int(d["J"]), max_order=1)
    assert np.allclose(y1, d["Sx"][..., :y1.shape[-3], :, :])


SMALL = ["c1_J2_32", "J3_64", "J2_33x47", "J1_31", "J4_32_L4", "J5_32", "J3_64_o1", "J2_prepad48", "J2_24x40"]


@pytest.mark.parametrize("name", SMALL)
def test_oracle_vs_reference_golden(golden_dir, name):
    d = np.load(os.path.join(golden_dir, f"golden_2d_{name}.npz"))
    y = o2.scattering2d(d["x"].astype(np.float64), int(d["J"]), int(d["L"]), int(d["max_order"]),
                        bool(d["pre_pad"]))
    assert y.shape == d["Sx64"].shape
    assert np.abs(y - d["Sx64"]).max() <= 1e-6 * np.abs(d["Sx64"]).max()
    assert y.shape[-3] == o2.n_channels(int(d["J"]), int(d["L"]), int(d["max_order"]))


def test_oracle_headline_shape(golden_dir):
    d = np.load(os.path.join(golden_dir, "golden_2d_c2_J3_256.npz"))
    y = o2.scattering2d(d["x"][:1].astype(np.float64), 3, 8)
    assert y.shape == (1, 217, 32, 32)
    assert np.abs(y - d["Sx64"][:1]).max() <= 1e-6 * np.abs(d["Sx64"]).max()


def test_oracle_channel_order():
    # [S0] + [S1 by n1] + [S2 by (n1, n2), j2 > j1]  (core/scattering2d.py:80-86)
    rng = np.random.RandomState(0)
    y, paths = o2.scattering2d(rng.randn(1, 16, 16), 2, 4, return_paths=True)
    assert paths[0] == () and paths[1:9] == [(n,) for n in range(8)]
    assert paths[9:] == [(n1, n2) for n1 in range(4) for n2 in range(4, 8)]
    assert y.shape == (1, 1 + 8 + 16, 4, 4)
