"""The drop-in boundary: unmodified kymatio torch frontend + backend='torch_b200'.

Primitive tests mirror tests/scattering2d/test_torch_backend_2d.py of the reference; the integration
tests mirror tests/scattering2d/test_torch_scattering2d.py:46-77."""
import os

import numpy as np
import pytest
import torch

from conftest import import_reference
from parity import assert_parity

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def plugin():
    if not import_reference():
        pytest.skip("reference not installed under baseline/_ref")
    import kymatio_b200.kymatio_plugin as p
    p.install()
    yield p
    p.uninstall()


def test_string_and_object_backend_routes(plugin, golden_dir):
    from kymatio.torch import Scattering2D
    d = np.load(os.path.join(golden_dir, "golden_2d_c1_J2_32.npz"))
    x = torch.from_numpy(d["x"]).cuda()
    for be in ("torch_b200", plugin.backend2d):
        S = Scattering2D(2, (32, 32), backend=be).cuda()
        assert S.backend.name == "torch_b200"
        y = S(x)
        assert tuple(y.shape) == d["Sx64"].shape and y.is_contiguous()
        assert_parity(y.cpu().numpy(), d["Sx64"], what=str(be))


def test_reference_fixture_through_kymatio_frontend(plugin, golden_dir):
    from kymatio.torch import Scattering2D
    d = np.load(os.path.join(golden_dir, "ref_fixture_2d.npz"))
    x = torch.from_numpy(d["x"]).cuda()
    S = Scattering2D(int(d["J"]), x.shape[-2:], pre_pad=bool(d["pre_pad"]), backend="torch_b200").cuda()
    assert_parity(S(x).cpu().numpy(), d["Sx"], what="fixture")
    S1 = Scattering2D(int(d["J"]), x.shape[-2:], max_order=1, backend="torch_b200").cuda()
    y1 = S1(x)
    assert_parity(y1.cpu().numpy(), d["Sx"][..., :y1.shape[-3], :, :], what="fixture o1")


def test_same_as_reference_torch_backend_on_gpu(plugin):
    from kymatio.torch import Scattering2D
    x = torch.randn(3, 48, 40, device="cuda")
    for kw in (dict(), dict(out_type="list"), dict(max_order=1), dict(L=4)):
        Sr = Scattering2D(2, (48, 40), backend="torch", **kw).cuda()
        Sb = Scattering2D(2, (48, 40), backend="torch_b200", **kw).cuda()
        yr, yb = Sr(x), Sb(x)
        if kw.get("out_type") == "list":
            assert [(a["j"], a["n"], a["theta"]) for a in yr] == [(a["j"], a["n"], a["theta"]) for a in yb]
            yr = torch.stack([a["coef"] for a in yr], 1)
            yb = torch.stack([a["coef"] for a in yb], 1)
        assert_parity(yb.cpu().numpy(), yr.cpu().numpy(), what=str(kw))


def test_pre_pad_and_state_dict(plugin):
    from kymatio.torch import Scattering2D
    Sr = Scattering2D(2, (32, 32), pre_pad=True, backend="torch").cuda()
    Sb = Scattering2D(2, (32, 32), pre_pad=True, backend="torch_b200").cuda()
    assert list(Sr.state_dict()) == list(Sb.state_dict())
    x = torch.randn(2, 40, 40, device="cuda")
    assert_parity(Sb(x).cpu().numpy(), Sr(x).cpu().numpy(), what="pre_pad")


def test_unfused_protocol_drives_reference_core(plugin):
    """The unchanged per-primitive core on top of the torch_b200 primitives."""
    from kymatio.torch import Scattering2D
    plugin.install(fused=False)
    try:
        x = torch.randn(2, 32, 32, device="cuda")
        yb = Scattering2D(2, (32, 32), backend="torch_b200").cuda()(x)
        yr = Scattering2D(2, (32, 32), backend="torch").cuda()(x)
        assert_parity(yb.cpu().numpy(), yr.cpu().numpy(), what="unfused")
    finally:
        plugin.install(fused=True)


# ---------------------------------------------------------------------------- primitives
def test_pad_unpad(plugin):
    be = plugin.backend2d
    pad = be.Pad((2, 2, 2, 2), (4, 4))
    x = torch.randn(1, 4, 4, device="cuda")
    z = pad(x)
    assert z.shape == (1, 8, 8, 1)
    assert torch.allclose(z[0, 2, 2], x[0, 0, 0]) and torch.allclose(z[0, 1, 0], x[0, 1, 2])
    assert torch.allclose(z[0, 1, 1], x[0, 1, 1]) and torch.allclose(z[0, 1, 2], x[0, 1, 0])
    assert torch.allclose(z[0, 1, 3], x[0, 1, 1])
    ref = torch.nn.ReflectionPad2d(2)(x[None])[0]
    assert torch.equal(z[..., 0], ref)
    y = be.unpad(torch.randn(4, 4, 1, device="cuda"))
    assert y.shape == (2, 2)


def test_modulus(plugin):
    be = plugin.backend2d
    x = torch.rand(100, 10, 4, 2, device="cuda")
    y = be.modulus(x)
    assert torch.allclose(y[..., 0], torch.sqrt(torch.sum(x * x, 3)))
    with pytest.raises(TypeError) as e:
        be.modulus(x[..., 0].contiguous())
    assert "should be complex" in e.value.args[0]
    with pytest.raises(RuntimeError) as e:
        be.modulus(x[::2, ::2])
    assert "contiguous" in e.value.args[0]
    with pytest.raises(TypeError) as e:
        be.modulus(x.cpu())
    assert "Use the torch backend" in e.value.args[0]


def test_subsample_fourier(plugin):
    be = plugin.backend2d
    x = torch.rand(10, 1, 128, 128, 2, device="cuda")
    ref = x.view(10, 1, 16, 8, 16, 8, 2).mean(4).mean(2)
    z = be.subsample_fourier(x, k=16)
    assert z.shape == (10, 1, 8, 8, 2) and torch.allclose(ref, z, atol=1e-6)
    with pytest.raises(TypeError) as e:
        be.subsample_fourier(x[..., 0].clone(), k=16)
    assert "should be complex" in e.value.args[0]
    with pytest.raises(RuntimeError) as e:
        be.subsample_fourier(x[::2, ::2], k=16)
    assert "must be contiguous" in e.value.args[0]


def test_cdgmm(plugin):
    be = plugin.backend2d
    A = torch.randn(3, 2, 6, 5, 2, device="cuda")
    Br = torch.randn(6, 5, 1, device="cuda")
    Bc = torch.randn(6, 5, 2, device="cuda")
    assert torch.allclose(be.cdgmm(A, Br), A * Br, atol=1e-7, rtol=1e-6)
    ref = torch.view_as_real(torch.view_as_complex(A) * torch.view_as_complex(Bc))
    assert torch.allclose(be.cdgmm(A, Bc), ref, atol=1e-6, rtol=1e-6)
    with pytest.raises(RuntimeError) as e:
        be.cdgmm(A, torch.randn(4, 5, 1, device="cuda"))
    assert "not compatible" in e.value.args[0]
    with pytest.raises(TypeError) as e:
        be.cdgmm(A[..., :1].contiguous(), Br)
    assert "should be complex" in e.value.args[0]
    with pytest.raises(TypeError) as e:
        be.cdgmm(A.double(), Br)
    assert "must be of the same dtype" in e.value.args[0]
    with pytest.raises(TypeError) as e:
        be.cdgmm(A, Br.cpu())
    assert "must be on CPU" in e.value.args[0]


@pytest.mark.parametrize("shape", [(4, 4), (40, 40), (34, 20), (272, 272), (9, 15)])
def test_fft_against_torch(plugin, shape):
    be = plugin.backend2d
    x = torch.randn(3, *shape, 1, device="cuda")
    X = be.rfft(x)
    ref = torch.view_as_real(torch.fft.fft2(x[..., 0].to(torch.complex64)))
    assert torch.allclose(X, ref, atol=1e-4 * float(ref.abs().max()))
    z = torch.randn(3, *shape, 2, device="cuda")
    zi = be.ifft(z)
    refi = torch.view_as_real(torch.fft.ifft2(torch.view_as_complex(z)))
    assert torch.allclose(zi, refi, atol=1e-5 * float(refi.abs().max()) + 1e-7)
    assert torch.allclose(be.irfft(z)[..., 0], refi[..., 0], atol=1e-5 * float(refi.abs().max()) + 1e-7)
    assert torch.allclose(be.ifft(be.rfft(x))[..., 0], x[..., 0], atol=1e-5)
    with pytest.raises(TypeError):
        be.rfft(z)
    with pytest.raises(TypeError):
        be.ifft(x)
    with pytest.raises(RuntimeError):
        be.ifft(z[::2])


def test_2d_eager_primitives_refuse_gradients(plugin):
    """The 2-D eager primitives are forward-only (gradients come from the fused path): asking for a gradient
    through them raises instead of returning a silently detached tensor."""
    x = torch.randn(2, 8, 8, 2, device="cuda", requires_grad=True)
    with pytest.raises(RuntimeError, match="does not propagate gradients"):
        plugin.backend2d.modulus(x)
