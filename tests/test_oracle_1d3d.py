"""The 1-D / 3-D numpy oracles against reference-generated goldens and the reference's own fixtures.  The filter
banks are constructor-time code outside the hot path: they come from the unmodified reference (baseline/_ref)."""
import os

import numpy as np
import pytest

from conftest import import_reference
from oracle import scattering1d as o1
from oracle import scattering3d as o3


@pytest.fixture(scope="module")
def ref():
    if not import_reference():
        pytest.skip("reference not installed under baseline/_ref")
    from kymatio.scattering1d.frontend.numpy_frontend import ScatteringNumPy1D
    from kymatio.scattering3d.frontend.numpy_frontend import HarmonicScatteringNumPy3D
    return ScatteringNumPy1D, HarmonicScatteringNumPy3D


def _kw1d(d):
    kw = dict(J=int(d["J"]), shape=int(d["shape"]) if "shape" in d.files else d["x"].shape[-1])
    Q = tuple(int(q) for q in np.atleast_1d(d["Q"]))
    kw["Q"] = Q if len(Q) > 1 else Q[0]
    if "max_order" in d.files:
        kw["max_order"] = int(d["max_order"])
    return kw


def _run1d(S, x):
    return o1.scattering1d(x, S.phi_f, S.psi1_f, S.psi2_f, S.log2_stride, S.pad_left, S.pad_right, S.ind_start,
                           S.ind_end, S.max_order)


@pytest.mark.parametrize("name", ["J5_Q4_2048", "J6_Q16_512", "J4_Q2_1000_o1"])
def test_oracle1d_vs_reference_golden(ref, golden_dir, name):
    d = np.load(os.path.join(golden_dir, f"golden_1d_{name}.npz"))
    S = ref[0](**_kw1d(d))
    y = _run1d(S, d["x"])
    assert y.shape == d["Sx64"].shape
    assert np.abs(y - d["Sx64"]).max() <= 2e-6 * np.abs(d["Sx64"]).max()       # x was stored as float32
    keys = o1.meta_order(S.psi1_f, S.psi2_f, S.max_order)
    assert [str(k) for k in keys] == list(d["key"])


def test_oracle1d_reference_fixture(ref, golden_dir):
    # tests/scattering1d/test_torch_scattering1d.py:82-112 (the reference's own golden file)
    d = np.load(os.path.join(golden_dir, "ref_fixture_1d.npz"))
    S = ref[0](int(d["J"]), d["x"].shape[-1], int(d["Q"]))
    y = _run1d(S, d["x"])
    assert np.allclose(y, d["Sx"], atol=1e-6 * np.abs(d["Sx"]).max() + 1e-7)


def test_engine1d_schedule_matches_oracle_order(ref):
    from kymatio_b200.engine1d import schedule
    S = ref[0](J=6, shape=4096, Q=(8, 2))
    sch = schedule(S._N_padded, S.log2_stride, S.phi_f, S.psi1_f, S.psi2_f)
    keys = o1.meta_order(S.psi1_f, S.psi2_f, 2)
    assert sch["K"] == len(keys) and sch["M"] == S._N_padded >> S.log2_stride
    got = {}
    for kind, n1, n2, ch in sch["order"]:
        got[ch] = () if kind == "S0" else (n1,) if kind == "S1" else (n1, n2)
    assert [got[c] for c in range(sch["K"])] == keys
    # every path is launched exactly once, with the reference's subsampling exponents (core/scattering1d.py:60,89-90)
    seen = set()
    for g in sch["groups"]:
        for n1, ch in zip(g["n1"], g["chan"]):
            assert S.psi1_f[n1]["j"] == g["j1"] and g["k1"] == min(g["j1"], S.log2_stride) and keys[ch] == (n1,)
            seen.add(ch)
        for c in g["children"]:
            assert c["j2"] > g["j1"] and c["k2"] == max(min(c["j2"], S.log2_stride) - g["k1"], 0)
            for n1, ch in zip(g["n1"], c["chan"]):
                assert keys[ch] == (n1, c["n2"])
                seen.add(ch)
    assert seen == set(range(1, sch["K"]))


def test_engine1d_support_helpers():
    from kymatio_b200.engine1d import circular_support, lowpass_bins
    f = np.zeros(100); f[95:] = 1; f[:4] = 1
    assert circular_support(f, 1e-7) == (95, 9)
    f = np.zeros(100); f[10:20] = 1
    assert circular_support(f, 1e-7) == (10, 10)
    assert circular_support(np.ones(64), 1e-7)[1] == 64
    assert circular_support(np.zeros(8), 1e-7) == (0, 0)
    n = np.arange(1024)
    g = np.exp(-0.5 * (np.minimum(n, 1024 - n) / 10.0) ** 2)
    Fc = lowpass_bins(g, 1e-9)
    assert Fc % 16 == 0 and g[Fc] <= 1e-9 and g[Fc - 17] > 1e-9
    assert lowpass_bins(np.ones(64), 1e-9) == 33


@pytest.mark.parametrize("name", ["J2_L2_16", "J1_L3_12x16x20"])
def test_oracle3d_vs_reference_golden(ref, golden_dir, name):
    d = np.load(os.path.join(golden_dir, f"golden_3d_{name}.npz"))
    kw = dict(J=int(d["J"]), shape=tuple(int(v) for v in d["shape"]), L=int(d["L"]))
    if "integral_powers" in d.files:
        kw["integral_powers"] = tuple(float(v) for v in d["integral_powers"])
    S = ref[1](**kw)
    y = o3.scattering3d(d["x"], S.filters, S.L, S.J, S.integral_powers, S.max_order, S.rotation_covariant)
    assert y.shape == d["Sx64"].shape
    assert np.abs(y - d["Sx64"]).max() <= 2e-6 * np.abs(d["Sx64"]).max()


def test_oracle3d_reference_fixture(ref, golden_dir):
    # tests/scattering3d/test_torch_scattering3d.py (test_data_3d.npz, relative L1 error gate of the reference)
    d = np.load(os.path.join(golden_dir, "ref_fixture_3d.npz"))
    S = ref[1](int(d["J"]), d["x"].shape[-3:], L=int(d["L"]), sigma_0=1, integral_powers=tuple(d["integral_powers"]))
    y = o3.scattering3d(d["x"], S.filters, S.L, S.J, S.integral_powers, S.max_order, S.rotation_covariant)
    y = y.reshape(y.shape[0], -1)
    ref_s = d["Sx"]
    order_0 = None
    # the fixture stacks order 0 (integrals of |x|) in front of the scattering coefficients in some versions
    if ref_s.shape[1] != y.shape[1]:
        order_0 = ref_s.shape[1] - y.shape[1]
        ref_s = ref_s[:, order_0:]
    assert np.abs(y - ref_s).sum() / np.abs(ref_s).sum() < 1e-5


LIVE_1D = [dict(J=6, shape=4096, Q=(8, 2)), dict(J=5, shape=3000, Q=(6, 1), stride=8), dict(J=7, shape=8192, Q=12, T=32),
           dict(J=4, shape=777, Q=3, max_order=1), dict(J=5, shape=2048, Q=(4, 1), oversampling=1)]


@pytest.mark.parametrize("kw", LIVE_1D)
def test_oracle1d_vs_live_reference(ref, kw):
    """The restatement against the reference itself run here (numpy frontend, float64) on configurations the committed
    goldens do not cover: stride, T, oversampling, odd lengths, max_order=1."""
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        S = ref[0](**kw)
    x = np.random.RandomState(3).randn(2, kw["shape"])
    y = _run1d(S, x)
    want = S(x)
    assert y.shape == want.shape
    assert np.abs(y - want).max() <= 1e-10 * np.abs(want).max()


@pytest.mark.parametrize("kw", [dict(J=2, shape=(16, 16, 16), L=2, rotation_covariant=False),
                                dict(J=2, shape=(12, 10, 14), L=1, max_order=1, integral_powers=(0.5, 1.0, 2.0, 3.0))])
def test_oracle3d_vs_live_reference(ref, kw):
    S = ref[1](**kw)
    x = np.random.RandomState(4).randn(2, *kw["shape"])
    y = o3.scattering3d(x, S.filters, S.L, S.J, S.integral_powers, S.max_order, S.rotation_covariant)
    want = S(x)
    assert y.shape == want.shape
    assert np.abs(y - want).max() <= 1e-5 * np.abs(want).max()       # the reference casts its integrals to float32


@pytest.mark.parametrize("key", ["8x10x12_J2_L3", "9x8x7_J1_L2"])
def test_filter_bank_3d_closed_form_matches_reference(golden_dir, key):
    """The closed form the device-side synthesis implements (oracle/filters3d.py) against banks generated by the reference
    (kymatio/scattering3d/filter_bank.py:5-166; even and odd axis lengths)."""
    from oracle import filters3d
    g = np.load(os.path.join(golden_dir, "golden_filters_3d.npz"))
    M, N, O, J, L, s0 = g[key + "_cfg"]
    M, N, O, J, L = int(M), int(N), int(O), int(J), int(L)
    bank = filters3d.solid_harmonic_filter_bank(M, N, O, J, L, float(s0))
    for l in range(L + 1):
        assert np.abs(bank[l] - g[key + f"_l{l}"]).max() <= 2e-7      # the reference rounds its grid and bank to float32
    assert np.abs(filters3d.gaussian_filter_bank(M, N, O, J, float(s0)) - g[key + "_gauss"]).max() <= 2e-7
