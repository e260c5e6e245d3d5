"""Gradients of the 2D scattering w.r.t. its input (reference: tests/scattering2d/
test_torch_scattering2d.py:238-248 gradcheck in float64; kymatio/backend/torch_backend.py:64-96)."""
import numpy as np
import pytest
import torch

from conftest import import_reference
from parity import assert_parity

pytestmark = pytest.mark.gpu


def test_eager_graph_matches_fused_forward():
    from kymatio_b200 import Scattering2D
    from kymatio_b200.ops2d import eager_scattering2d
    for (J, shape, L, mo) in [(2, (32, 32), 8, 2), (3, (40, 56), 4, 2), (2, (24, 24), 8, 1)]:
        S = Scattering2D(J, shape, L=L, max_order=mo).cuda()
        x = torch.randn(3, *shape, device="cuda")
        y = S(x)
        eng = S._engine(x.dtype, x.device)
        phi, psi = S.load_filters()
        t, l = (eng.Mp - shape[0]) // 2, (eng.Np - shape[1]) // 2
        pads = (t, eng.Mp - shape[0] - t, l, eng.Np - shape[1] - l)
        ye = eager_scattering2d(x, J, L, mo, pads, phi, psi)
        assert_parity(ye.cpu().numpy(), y.cpu().numpy(), tol=1e-5, what=str((J, shape)))


def test_gradcheck_float64():
    from kymatio_b200 import Scattering2D
    S = Scattering2D(2, (8, 8)).cuda().double()
    x = torch.rand(2, 1, 8, 8, device="cuda", dtype=torch.float64, requires_grad=True)
    assert torch.autograd.gradcheck(S, x, nondet_tol=1e-5, fast_mode=True)


def test_gradient_matches_reference_torch_backend():
    if not import_reference():
        pytest.skip("reference not installed under baseline/_ref")
    from kymatio.torch import Scattering2D as RefScattering2D
    from kymatio_b200 import Scattering2D
    torch.manual_seed(0)
    for (J, shape) in [(2, (32, 32)), (3, (48, 40))]:
        x = torch.randn(2, *shape, device="cuda")
        w = torch.randn(2, 1 + 8 * J + 64 * J * (J - 1) // 2, shape[0] // 2 ** J, shape[1] // 2 ** J, device="cuda")
        grads = []
        for cls in (RefScattering2D, Scattering2D):
            S = cls(J, shape).cuda()
            xi = x.clone().requires_grad_(True)
            (S(xi) * w).sum().backward()
            grads.append(xi.grad)
        assert_parity(grads[1].cpu().numpy()[:, None], grads[0].cpu().numpy()[:, None], tol=1e-4, what=str((J, shape)))


def test_zero_input_gradient_is_finite():
    from kymatio_b200 import Scattering2D
    S = Scattering2D(2, (16, 16)).cuda()
    x = torch.zeros(1, 16, 16, device="cuda", requires_grad=True)
    S(x).sum().backward()
    assert torch.isfinite(x.grad).all()


def test_fused_order2_block_matches_per_op_graph():
    """Order2 (fused tile forward + backward) vs the same block on the per-op graph: values and gradients."""
    from kymatio_b200 import Scattering2D
    from kymatio_b200.ops2d import eager_scattering2d
    for (J, shape, L, dt, tol) in [(3, (64, 64), 8, torch.float32, 2e-5), (2, (24, 40), 4, torch.float64, 1e-9),
                                   (3, (256, 256), 8, torch.float32, 2e-5)]:
        S = Scattering2D(J, shape, L=L).cuda()
        if dt == torch.float64:
            S = S.double()
        B = 2
        x = torch.randn(B, *shape, device="cuda", dtype=dt)
        S(x)                                              # binds the filters
        eng = S._engine(dt, x.device)
        # float64 keeps the exact Fourier low-pass (float32-born filters are rank-1 only to 2.5e-7): no fused block
        assert eng.order2_channels(0) == (L * (J - 1) * L if dt == torch.float32 else 0)
        phi, psi = S.load_filters()
        t, l = (eng.Mp - shape[0]) // 2, (eng.Np - shape[1]) // 2
        pads = (t, eng.Mp - shape[0] - t, l, eng.Np - shape[1] - l)
        w = torch.randn(B, eng.K, eng.out_h, eng.out_w, device="cuda", dtype=dt)
        outs, grads = [], []
        for e in (None, eng):
            xi = x.clone().requires_grad_(True)
            y = eager_scattering2d(xi, J, L, 2, pads, phi, psi, eng=e)
            (y * w).sum().backward()
            outs.append(y.detach()); grads.append(xi.grad)
        assert_parity(outs[1].cpu().numpy(), outs[0].cpu().numpy(), tol=tol, what="order2 fwd")
        assert_parity(grads[1].cpu().numpy()[:, None], grads[0].cpu().numpy()[:, None], tol=10 * tol, what="order2 bwd")


@pytest.mark.parametrize("J,shape,B", [(4, (224, 224), 2), (3, (256, 256), 1)])
def test_gradient_at_config_shapes_through_kymatio_plugin(J, shape, B):
    """BASELINE configs[4] (C5: J=4, 224x224) and the headline shape: gradient of the UNMODIFIED
    kymatio.torch.Scattering2D(backend='torch_b200') against the reference torch backend's autograd (contract:
    tests/scattering2d/test_torch_scattering2d.py:238-248 + kymatio/backend/torch_backend.py:64-96).  This exercises the
    static 128/64/32 (resp. 136/68) backward tiles and their atomics scatter against something other than the repo."""
    if not import_reference():
        pytest.skip("reference not installed under baseline/_ref")
    import kymatio_b200.kymatio_plugin as plugin
    plugin.install()
    try:
        from kymatio.torch import Scattering2D as KScattering2D
        torch.manual_seed(4)
        x = torch.randn(B, *shape, device="cuda")
        Sb = KScattering2D(J, shape, backend="torch_b200").cuda()
        Sr = KScattering2D(J, shape, backend="torch").cuda()
        xb = x.clone().requires_grad_(True)
        yb = Sb(xb)
        w = torch.randn_like(yb)
        (yb * w).sum().backward()
        xr = x.clone().requires_grad_(True)
        yr = Sr(xr)
        (yr * w).sum().backward()
        assert_parity(yb.detach().cpu().numpy(), yr.detach().cpu().numpy(), what="fwd " + str((J, shape)))
        assert_parity(xb.grad.cpu().numpy()[:, None], xr.grad.cpu().numpy()[:, None], tol=1e-4, what="grad " + str((J, shape)))
    finally:
        plugin.uninstall()


@pytest.mark.parametrize("J,shape,B", [(4, (224, 224), 3), (3, (256, 256), 2), (3, (40, 56), 5)])
def test_kept_spectra_equal_recomputed(J, shape, B, monkeypatch):
    """Training-mode forward (scat_plan2d_forward_save): same output as the inference forward, the kept first-order spectra
    are the ones the first-order blocks would recompute, and the gradient does not depend on which of the two the backward
    uses (atomics in the scatter: equal to float32 summation order)."""
    from kymatio_b200 import Scattering2D
    torch.manual_seed(3)
    S = Scattering2D(J, shape).cuda()
    x = torch.randn(B, *shape, device="cuda")
    eng = S._engine(x.dtype, x.device)
    with torch.no_grad():
        y0 = S(x)
    y1, saved = eng.forward_saving(x)
    assert torch.equal(y0, y1)
    assert saved is not None         # also at sizes without compiled instances (runtime-size tile kernels)
    U0 = None
    for j1, kept in enumerate(saved):
        if kept is None:
            continue
        if U0 is None:
            from kymatio_b200.ops2d import PadReflect, Fft2, _to_complex
            t, l = (eng.Mp - shape[0]) // 2, (eng.Np - shape[1]) // 2
            U0 = Fft2.apply(_to_complex(PadReflect.apply(x, (t, eng.Mp - shape[0] - t, l, eng.Np - shape[1] - l))), False)
        _, u1 = eng.order1_forward(j1, U0, B, True)
        assert tuple(kept.shape) == tuple(u1.shape)
        assert (kept - u1).abs().max() <= 2e-5 * u1.abs().max()
    w = torch.randn_like(y0)

    def grad():
        xi = x.clone().requires_grad_(True)
        (S(xi) * w).sum().backward()
        return xi.grad

    g_keep = grad()
    monkeypatch.setenv("SCAT_B200_SAVE_U1", "0")
    g_recompute = grad()
    assert (g_keep - g_recompute).abs().max() <= 1e-5 * g_recompute.abs().max()


@pytest.mark.parametrize("kw,shape", [(dict(J=2, max_order=1), (32, 32)), (dict(J=2, pre_pad=True), (32, 32)),
                                      (dict(J=3, L=4), (64, 64)), (dict(J=4, max_order=1), (224, 224)),
                                      (dict(J=2, out_type="list"), (28, 28))])
def test_gradient_variants_match_reference_torch_backend(kw, shape):
    """max_order=1, pre_pad, other L, list output: the fused blocks (kept spectra, static and runtime-size tiles) against
    the reference torch backend's autograd, through the unmodified kymatio frontend."""
    if not import_reference():
        pytest.skip("reference not installed under baseline/_ref")
    import kymatio_b200.kymatio_plugin as plugin
    plugin.install()
    from kymatio.torch import Scattering2D
    torch.manual_seed(1)
    in_shape = tuple((n // 2 ** kw["J"] + 2) * 2 ** kw["J"] for n in shape) if kw.get("pre_pad") else shape   # already padded
    x = torch.randn(2, *in_shape, device="cuda")
    grads = []
    for backend in ("torch", "torch_b200"):
        S = Scattering2D(shape=shape, backend=backend, **kw).cuda()
        xi = x.clone().requires_grad_(True)
        y = S(xi)
        if kw.get("out_type") == "list":
            y = torch.stack([c["coef"] for c in y], dim=1)
        torch.manual_seed(2)
        w = torch.randn_like(y)
        (y * w).sum().backward()
        grads.append(xi.grad)
    plugin.uninstall()
    assert_parity(grads[1].cpu().numpy()[:, None], grads[0].cpu().numpy()[:, None], tol=1e-4, what=str(kw))
