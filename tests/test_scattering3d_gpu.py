"""3-D solid-harmonic scattering through the unmodified kymatio torch frontend with
backend='torch_b200' (eager primitives), against reference-generated goldens and the reference's own
fixture (tests/scattering3d/test_torch_scattering3d.py:151-201: rel-L1 < 1e-6 there in fp64; here fp32)."""
import os

import numpy as np
import pytest
import torch

from conftest import import_reference

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def plugin():
    if not import_reference():
        pytest.skip("reference not installed under baseline/_ref")
    import kymatio_b200.kymatio_plugin as p
    p.install()
    return p


def _rel_l1(a, b):
    return float(np.abs(a - b).sum() / max(np.abs(a).sum(), np.abs(b).sum()))


@pytest.mark.parametrize("shape", [(8, 8, 8), (12, 16, 20), (32, 32, 32), (128, 128, 128)])
def test_fft3d_against_torch(plugin, shape):
    be = plugin.backend3d
    B = 1 if shape[0] >= 128 else 2
    z = torch.randn(B, *shape, 2, device="cuda")
    ref = torch.view_as_real(torch.fft.fftn(torch.view_as_complex(z), dim=[-1, -2, -3]))
    tol = 3e-6 * float(ref.abs().max()) * np.log2(np.prod(shape))
    assert torch.allclose(be._fft(z, False), ref, atol=tol)
    refi = torch.view_as_real(torch.fft.ifftn(torch.view_as_complex(z), dim=[-1, -2, -3]))
    assert torch.allclose(be.ifft(z), refi, atol=3e-6 * float(refi.abs().max()) * np.log2(np.prod(shape)) + 1e-9)


def test_modulus_rotation_and_integrals(plugin):
    be = plugin.backend3d
    x = torch.randn(2, 6, 5, 4, 2, device="cuda")
    m1 = be.modulus_rotation(x)
    assert torch.allclose(m1, torch.sqrt((x ** 2).sum(-1, keepdim=True)))
    y = torch.randn(2, 6, 5, 4, 2, device="cuda")
    m2 = be.modulus_rotation(y, m1)
    assert torch.allclose(m2, torch.sqrt(m1 ** 2 + (y ** 2).sum(-1, keepdim=True)), atol=1e-6)
    u = torch.rand(3, 7, 6, 5, 1, device="cuda") + 0.1
    got = be.compute_integrals(u, (0.5, 1.0, 2.0))
    ref = torch.stack([(u ** q).reshape(3, -1).sum(1) for q in (0.5, 1.0, 2.0)], 1)
    assert got.shape == (3, 3) and torch.allclose(got, ref, rtol=1e-5)


@pytest.mark.parametrize("name", ["J2_L2_16", "J2_L2_32", "J1_L3_12x16x20"])
def test_scattering3d_golden(plugin, golden_dir, name):
    from kymatio.torch import HarmonicScattering3D
    d = np.load(os.path.join(golden_dir, f"golden_3d_{name}.npz"))
    kw = dict(J=int(d["J"]), shape=tuple(int(v) for v in d["shape"]), L=int(d["L"]))
    if "integral_powers" in d.files:
        kw["integral_powers"] = tuple(float(v) for v in d["integral_powers"])
    S = HarmonicScattering3D(backend="torch_b200", **kw).cuda()
    y = S(torch.from_numpy(d["x"]).cuda())
    assert tuple(y.shape) == d["Sx64"].shape
    assert _rel_l1(y.cpu().numpy().astype(np.float64), d["Sx64"]) < 1e-4


def test_scattering3d_reference_fixture(plugin, golden_dir):
    from kymatio.torch import HarmonicScattering3D
    d = np.load(os.path.join(golden_dir, "ref_fixture_3d.npz"))
    x = torch.from_numpy(d["x"]).cuda()
    J, L, powers = int(d["J"]), int(d["L"]), tuple(float(v) for v in d["integral_powers"])
    S = HarmonicScattering3D(J, x.shape[1:], L=L, sigma_0=1, integral_powers=powers, max_order=2,
                             backend="torch_b200").cuda()
    order_0 = plugin.backend3d.compute_integrals(x, powers)
    y = S(x)
    out = torch.cat([order_0.reshape(x.shape[0], -1), y.reshape(x.shape[0], -1)], 1).cpu().numpy()
    assert out.shape == d["Sx"].shape
    assert _rel_l1(out.astype(np.float64), d["Sx"].astype(np.float64)) < 1e-4


FUSED_CASES = [
    dict(J=2, shape=(16, 16, 16), L=2),
    dict(J=1, shape=(32, 16, 16), L=3, integral_powers=(1.0, 2.0)),
    dict(J=2, shape=(32, 32, 32), L=1, max_order=1),
    dict(J=2, shape=(16, 16, 16), L=2, rotation_covariant=False),
    dict(J=2, shape=(64, 64, 64), L=1, integral_powers=(0.5, 1.0, 2.0, 3.0)),
    dict(J=1, shape=(96, 64, 48), L=1),                      # non-power-of-two instance (radix-3 factors along M and O)
    dict(J=2, shape=(128, 128, 128), L=2),                   # BASELINE configs[3] (C4) at its own size
]


@pytest.mark.parametrize("kw", FUSED_CASES)
def test_fused3d_vs_reference_torch_float64(plugin, kw):
    """The fused band kernels (col_prod / plane / col_fwd) against the reference's own torch backend run in float64
    on the same device, element by element (integrals are sums of positive terms: relative error is meaningful)."""
    from kymatio.torch import HarmonicScattering3D
    from kymatio_b200 import _lib
    torch.manual_seed(3)
    x = torch.randn(2, *kw["shape"], device="cuda")
    Sb = HarmonicScattering3D(backend="torch_b200", **kw).cuda()
    Sr = HarmonicScattering3D(backend="torch", **kw).cuda().double()
    _lib.timing_enable(True)
    y = Sb(x)
    labels = {r["label"].split(":")[0] for r in _lib.timing_report()}
    _lib.timing_enable(False)
    assert {"3d_col_prod", "3d_plane_leaf"} <= labels, labels          # the fused kernels ran, not the eager primitives
    ref = Sr(x.double()).double()
    assert y.shape == ref.shape and y.dtype == torch.float32
    err = ((y.double() - ref).abs() / ref.abs().clamp_min(1e-30)).max().item()
    assert err < 1e-4, err


def test_fused3d_matches_eager_primitives(plugin):
    from kymatio.torch import HarmonicScattering3D
    x = torch.randn(3, 32, 32, 32, device="cuda")
    S = HarmonicScattering3D(J=2, shape=(32, 32, 32), L=2, backend="torch_b200").cuda()
    y = S(x)
    plugin.install(fused=False)
    try:
        ye = S(x)
    finally:
        plugin.install(fused=True)
    assert ((y - ye).abs() / ye.abs().clamp_min(1e-30)).max().item() < 1e-4


@pytest.mark.parametrize("kw", [dict(J=1, shape=(16, 16, 16), L=1), dict(J=2, shape=(8, 12, 10), L=2, integral_powers=(1.0, 2.0)),
                                dict(J=1, shape=(16, 16, 16), L=2, rotation_covariant=False, max_order=1)])
def test_3d_gradients_match_reference_autograd(plugin, kw):
    """Gradients through backend='torch_b200' HarmonicScattering3D (kernels + hand-written adjoints of cdgmm3d, fftn,
    modulus_rotation and compute_integrals) against the reference torch backend's autograd in float64."""
    from kymatio.torch import HarmonicScattering3D
    torch.manual_seed(1)
    Sb = HarmonicScattering3D(backend="torch_b200", **kw).cuda()
    Sr = HarmonicScattering3D(backend="torch", **kw).cuda().double()
    x = torch.randn(2, *kw["shape"], device="cuda")
    xb = x.clone().requires_grad_(True)
    yb = Sb(xb)
    w = torch.rand_like(yb) / yb.detach().abs().clamp_min(1e-12)          # equalise the very different scales of the powers
    (yb * w).sum().backward()
    xr = x.double().requires_grad_(True)
    (Sr(xr).double() * w.double()).sum().backward()
    assert xb.grad is not None and torch.isfinite(xb.grad).all()
    err = (xb.grad.double() - xr.grad).abs().max() / xr.grad.abs().max()
    assert err < 1e-3, float(err)


@pytest.mark.parametrize("shape", [(16, 16, 16), (8, 32, 32), (64, 16, 16), (128, 128, 128), (96, 64, 48), (192, 128, 96)])
def test_fused3d_rfft_against_torch(plugin, shape):
    """U0_hat of the fused path (half-plane forward transforms + radix-2 along O + transform along M)."""
    from kymatio_b200.engine3d import Engine3D
    eng = Engine3D(*shape, torch.device("cuda:0"))
    x = torch.randn(2, *shape, device="cuda")
    got = eng.rfft(x)
    ref = torch.view_as_real(torch.fft.fftn(x, dim=(-3, -2, -1)))
    assert (got - ref).abs().max() <= 3e-6 * ref.abs().max() * np.log2(np.prod(shape))


def test_fused3d_batch_shapes(plugin):
    from kymatio.torch import HarmonicScattering3D
    S = HarmonicScattering3D(J=1, shape=(16, 16, 16), L=1, backend="torch_b200").cuda()
    x = torch.randn(2, 2, 16, 16, 16, device="cuda")
    y = S(x)
    flat = S(x.reshape(4, 16, 16, 16))
    assert y.shape[:2] == (2, 2)
    assert ((y.reshape(flat.shape) - flat).abs() / flat.abs().clamp_min(1e-30)).max().item() < 1e-5   # float64 atomics
    one = S(x[0, 0])
    assert ((one - flat[0]).abs() / flat[0].abs().clamp_min(1e-30)).max().item() < 1e-5


def test_c4_size_golden(plugin, golden_dir):
    """BASELINE configs[3] at its own size against the committed reference-generated golden (input regenerated from
    the seed, tests/golden/make_golden.py:golden_3d_c4)."""
    from kymatio.torch import HarmonicScattering3D
    d = np.load(os.path.join(golden_dir, "golden_3d_c4_J2_L2_128.npz"))
    x = np.random.RandomState(int(d["seed"])).randn(1, 128, 128, 128).astype(np.float32)
    assert np.array_equal(x.ravel()[:8], d["x_first8"])
    S = HarmonicScattering3D(J=2, shape=(128, 128, 128), L=2, backend="torch_b200").cuda()
    y = S(torch.from_numpy(x).cuda()).cpu().numpy().astype(np.float64)
    assert y.shape == d["Sx64"].shape
    assert (np.abs(y - d["Sx64"]) / np.abs(d["Sx64"])).max() < 1e-4
