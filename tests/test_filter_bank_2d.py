"""Product filter bank against reference-generated checksums and the oracle restatement."""
import os

import numpy as np
import pytest

from kymatio_b200.filter_bank2d import filter_bank_2d, padded_size_2d
from oracle import scattering2d as o2


def _flat(fb):
    out = list(fb["phi"]["levels"])
    for p in fb["psi"]:
        out.extend(p["levels"])
    return out


@pytest.mark.parametrize("cfg", [(40, 40, 2, 8), (48, 64, 2, 6), (64, 64, 4, 8), (272, 272, 3, 8)])
def test_filters_match_reference_checksums(golden_dir, cfg):
    Mp, Np, J, L = cfg
    g = np.load(os.path.join(golden_dir, "golden_filters_2d.npz"))
    key = f"{Mp}x{Np}_J{J}_L{L}"
    levels = _flat(filter_bank_2d(Mp, Np, J, L))
    assert len(levels) == len(g[key + "_sum"])
    for i, lev in enumerate(levels):
        assert lev.dtype == np.float32
        scale = max(1.0, g[key + "_l2"][i])
        assert abs(lev.sum(dtype=np.float64) - g[key + "_sum"][i]) <= 2e-5 * scale
        assert abs(np.sqrt((lev.astype(np.float64) ** 2).sum()) - g[key + "_l2"][i]) <= 1e-6 * scale
        samp = lev.ravel()[:: max(1, lev.size // 16)][:16]
        assert np.abs(samp - g[key + "_samples"][i]).max() <= 1e-6
    if Mp <= 64:
        fb = filter_bank_2d(Mp, Np, J, L)
        assert np.abs(fb["phi"]["levels"][0] - g[key + "_phi0"]).max() <= 1e-6
        assert np.abs(fb["psi"][-1]["levels"][0] - g[key + "_psi_last_l0"]).max() <= 1e-6
        assert np.abs(fb["psi"][3]["levels"][0] - g[key + "_psi3_l0"]).max() <= 1e-6


def test_filters_match_oracle_layout():
    fb, ob = filter_bank_2d(40, 48, 3, 4), o2.filter_bank(40, 48, 3, 4)
    assert len(fb["psi"]) == len(ob["psi"]) == 12
    for p, q in zip(fb["psi"], ob["psi"]):
        assert (p["j"], p["theta"]) == (q["j"], q["theta"])
        assert len(p["levels"]) == len(q["levels"]) == min(p["j"] + 1, 2)
        for a, b in zip(p["levels"], q["levels"]):
            assert a.shape == b.shape and np.abs(a - b).max() <= 1e-6
    for a, b in zip(fb["phi"]["levels"], ob["phi"]["levels"]):
        assert a.shape == b.shape and np.abs(a - b).max() <= 1e-6


def test_padded_size():
    assert padded_size_2d(256, 256, 3) == (272, 272) == o2.padded_size(256, 256, 3)
    assert padded_size_2d(32, 32, 2) == (40, 40)
    assert padded_size_2d(224, 224, 4) == (256, 256)
    assert padded_size_2d(32, 32, 5) == (96, 96)
