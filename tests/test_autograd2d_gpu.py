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
