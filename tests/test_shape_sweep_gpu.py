"""Shape / parameter sweep of the fused 2D forward against the reference's own torch backend on the same GPU
(reference: tests/scattering2d/test_torch_scattering2d.py:180-222 size agnosticism; here much wider).
Exercises the generic runtime-size tile and streaming instances, odd prime factors of the padded size
(3, 5, 7, 11, 13, 17, 19, 23, ...), non-square images, pre-padding, L != 8, max_order = 1, large images."""
import numpy as np
import pytest
import torch

from conftest import import_reference
from parity import assert_parity

pytestmark = pytest.mark.gpu

CASES = [
    # J, shape, L, max_order, pre_pad, batch
    (1, (2, 2), 8, 2, False, 3), (1, (5, 9), 3, 2, False, 2), (2, (4, 4), 8, 2, False, 2),
    (2, (37, 21), 8, 2, False, 2), (3, (71, 103), 4, 2, False, 1), (3, (8, 8), 8, 2, False, 2),
    (2, (100, 60), 8, 1, False, 2), (4, (48, 80), 2, 2, False, 1), (4, (130, 130), 8, 2, False, 1),
    (3, (150, 170), 8, 2, False, 1), (2, (180, 90), 8, 2, False, 1), (5, (64, 96), 8, 2, False, 1),
    (3, (128, 128), 8, 2, False, 2), (2, (64, 64), 8, 2, False, 2), (3, (56, 56), 8, 2, True, 2),
    (2, (28, 44), 5, 2, True, 2), (1, (61, 61), 8, 2, False, 1), (3, (512, 512), 8, 2, False, 1),
    (2, (300, 500), 8, 2, False, 1), (4, (224, 224), 8, 2, False, 2), (3, (224, 224), 8, 2, False, 2),
    (3, (1000, 40), 8, 1, False, 1),
]


@pytest.fixture(scope="module")
def ref_cls():
    if not import_reference():
        pytest.skip("reference not installed under baseline/_ref")
    from kymatio.torch import Scattering2D
    return Scattering2D


@pytest.mark.parametrize("case", CASES, ids=[f"J{c[0]}_{c[1][0]}x{c[1][1]}_L{c[2]}_o{c[3]}{'_pp' if c[4] else ''}" for c in CASES])
def test_sweep_matches_reference_torch_backend(ref_cls, case):
    from kymatio_b200 import Scattering2D
    J, shape, L, mo, pre_pad, B = case
    Sr = ref_cls(J, shape, L=L, max_order=mo, pre_pad=pre_pad, backend="torch").cuda()
    Sb = Scattering2D(J, shape, L=L, max_order=mo, pre_pad=pre_pad).cuda()
    in_shape = (Sr._M_padded, Sr._N_padded) if pre_pad else shape
    torch.manual_seed(hash(case) % 1000)
    x = torch.randn(B, *in_shape, device="cuda")
    yr, yb = Sr(x), Sb(x)
    assert yb.shape == yr.shape
    # the reference's own fp32 path carries ~1e-6 error, so compare at 2e-5 (gate for the float64 oracle: 1e-4)
    assert_parity(yb.cpu().numpy(), yr.cpu().numpy(), tol=2e-5, what=str(case))


def test_large_batch_is_chunked_consistently():
    from kymatio_b200 import Scattering2D
    S = Scattering2D(2, (32, 32)).cuda()
    x = torch.randn(600, 32, 32, device="cuda")
    y = S(x)
    y2 = torch.cat([S(x[:100]), S(x[100:])])
    assert torch.equal(y, y2)
