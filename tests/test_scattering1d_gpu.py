"""1-D scattering through the unmodified kymatio torch frontend with backend='torch_b200' (eager
primitives: every arithmetic op is one of this library's kernels), against reference-generated goldens
and the reference's own fixture (tests/scattering1d/test_torch_scattering1d.py:82-112)."""
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
    return p


@pytest.mark.parametrize("n", [8, 64, 96, 512, 1000, 4096, 2 ** 17])
def test_fft1d_against_torch(plugin, n):
    be = plugin.backend1d
    z = torch.randn(3, 1, n, 2, device="cuda")
    ref = torch.view_as_real(torch.fft.fft(torch.view_as_complex(z)))
    out = be.cfft(z)
    assert torch.allclose(out, ref, atol=2e-6 * float(ref.abs().max()) * np.log2(n))
    refi = torch.view_as_real(torch.fft.ifft(torch.view_as_complex(z)))
    assert torch.allclose(be.ifft(z), refi, atol=2e-6 * float(refi.abs().max()) * np.log2(n) + 1e-8)
    x = torch.randn(2, 1, n, 1, device="cuda")
    assert torch.allclose(be.irfft(be.rfft(x)), x, atol=1e-5)


def test_pad_subsample_1d(plugin):
    be = plugin.backend1d
    x = torch.randn(2, 1, 16, device="cuda")
    assert torch.equal(be.pad(x, 5, 7)[..., 0], torch.nn.functional.pad(x, (5, 7), mode="reflect"))
    with pytest.raises(ValueError):
        be.pad(x, 16, 2)
    z = torch.randn(3, 2, 64, 2, device="cuda")
    ref = z.view(3, 2, 4, 16, 2).mean(dim=-3)
    assert torch.allclose(be.subsample_fourier(z, 4), ref, atol=1e-6)
    # F^-1(periodised) == decimated F^-1   (tests/scattering1d/test_torch_backend_1d.py:141-172)
    t = torch.randn(1, 1, 128, 1, device="cuda")
    lhs = be.irfft(be.subsample_fourier(be.rfft(t), 2))
    assert torch.allclose(lhs[..., 0], t[..., ::2, 0], atol=1e-5)


@pytest.mark.parametrize("name", ["J5_Q4_2048", "J6_Q16_512", "J4_Q2_1000_o1", "J8_Q8_65536"])
def test_scattering1d_golden(plugin, golden_dir, name):
    from kymatio.torch import Scattering1D
    d = np.load(os.path.join(golden_dir, f"golden_1d_{name}.npz"))
    kw = dict(J=int(d["J"]), shape=int(d["shape"]), Q=tuple(int(q) for q in np.atleast_1d(d["Q"])))
    if len(kw["Q"]) == 1:
        kw["Q"] = kw["Q"][0]
    if "max_order" in d.files:
        kw["max_order"] = int(d["max_order"])
    S = Scattering1D(backend="torch_b200", **kw).cuda()
    x = torch.from_numpy(d["x"]).cuda()
    y = S(x)
    assert tuple(y.shape) == d["Sx64"].shape
    assert_parity(y.cpu().numpy(), d["Sx64"], channel_axis=-2, what=name)


def test_scattering1d_reference_fixture(plugin, golden_dir):
    from kymatio.torch import Scattering1D
    d = np.load(os.path.join(golden_dir, "ref_fixture_1d.npz"))
    x = torch.from_numpy(d["x"]).cuda()
    S = Scattering1D(int(d["J"]), x.shape[-1], int(d["Q"]), backend="torch_b200").cuda()
    y = S(x)
    assert_parity(y.cpu().numpy(), d["Sx"], channel_axis=-2, what="fixture 1d")
