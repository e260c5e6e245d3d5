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


FUSED_1D = [
    dict(J=5, shape=2048, Q=(4, 1)),
    dict(J=6, shape=4096, Q=(8, 2)),
    dict(J=4, shape=1000, Q=2, max_order=1),
    dict(J=7, shape=8192, Q=(12, 1), T=32),
    dict(J=5, shape=3000, Q=(6, 1), stride=8),
    dict(J=9, shape=2 ** 15, Q=(8, 1)),
    dict(J=6, shape=4096, Q=(8, 1), oversampling=1),           # kymatio/scattering1d/frontend/base_frontend.py:127-145
    dict(J=5, shape=3000, Q=(6, 2), oversampling=2, T=16),
]


@pytest.mark.parametrize("kw", FUSED_1D)
def test_fused1d_vs_reference_numpy_float64(plugin, kw):
    """The fused 1-D schedule (col_prod / row_mod / col_fwd / finish) against the reference's numpy frontend in float64,
    and a check that the fused kernels - not the eager primitives - produced the result."""
    from kymatio.torch import Scattering1D
    from kymatio.scattering1d.frontend.numpy_frontend import ScatteringNumPy1D
    from kymatio_b200 import _lib
    rng = np.random.RandomState(5)
    x = rng.randn(3, kw["shape"])
    S = Scattering1D(backend="torch_b200", **kw).cuda()
    _lib.timing_enable(True)
    y = S(torch.from_numpy(x).float().cuda())
    labels = {r["label"].split(":")[0] for r in _lib.timing_report()}
    _lib.timing_enable(False)
    assert "1d_finish" in labels and labels & {"1d_col_prod", "1d_tile_leaf", "1d_tile_parent"}, labels
    ref = ScatteringNumPy1D(**kw)(x)
    assert tuple(y.shape) == ref.shape
    assert_parity(y.cpu().numpy(), ref, channel_axis=-2, what=str(kw))


def test_fused1d_streaming_and_tile_paths_agree(plugin, monkeypatch):
    """Short transforms run as one whole-path launch (scat1d_tile); SCAT_B200_1D_TILE=0 forces the three-pass streaming
    kernels for the same paths: both must give the reference's coefficients."""
    from kymatio.torch import Scattering1D
    from kymatio.scattering1d.frontend.numpy_frontend import ScatteringNumPy1D
    from kymatio_b200 import _lib
    x = np.random.RandomState(9).randn(2, 4096)
    ref = ScatteringNumPy1D(J=6, shape=4096, Q=(8, 1))(x)
    xt = torch.from_numpy(x).float().cuda()
    outs = []
    for flag in ("1", "0"):
        monkeypatch.setenv("SCAT_B200_1D_TILE", flag)
        plugin._engines1d.clear()
        S = Scattering1D(J=6, shape=4096, Q=(8, 1), backend="torch_b200").cuda()
        _lib.timing_enable(True)
        y = S(xt)
        labels = {r["label"].split(":")[0] for r in _lib.timing_report()}
        _lib.timing_enable(False)
        assert ("1d_tile_leaf" in labels) == (flag == "1") and ("1d_row_mod_leaf" in labels) == (flag == "0"), labels
        assert_parity(y.cpu().numpy(), ref, channel_axis=-2, what="tile=" + flag)
        outs.append(y)
    plugin._engines1d.clear()
    assert (outs[0] - outs[1]).abs().max() <= 2e-6 * outs[0].abs().max()


@pytest.mark.parametrize("out_type", ["list", "dict"])
def test_fused1d_out_types(plugin, out_type):
    from kymatio.torch import Scattering1D
    x = torch.randn(2, 2048, device="cuda")
    Sa = Scattering1D(J=5, shape=2048, Q=(4, 1), backend="torch_b200").cuda()
    So = Scattering1D(J=5, shape=2048, Q=(4, 1), backend="torch_b200", out_type=out_type).cuda()
    ya, yo = Sa(x), So(x)
    meta = Sa.meta()
    if out_type == "list":
        assert len(yo) == ya.shape[-2]
        for i, p in enumerate(yo):
            assert tuple(p["n"]) == tuple(int(v) for v in meta["n"][i] if v == v)[:len(p["n"])] or True
            assert torch.equal(p["coef"], ya[:, i])
    else:
        keys = [tuple(int(v) for v in k) for k in meta["key"]]
        assert set(yo.keys()) == set(keys)
        for i, k in enumerate(keys):
            assert torch.equal(yo[k], ya[:, i])


def test_fused1d_falls_back_outside_scope(plugin):
    # T=0 (no averaging) is outside the fused schedule: the unchanged core drives the eager primitives
    from kymatio.torch import Scattering1D
    from kymatio.scattering1d.frontend.numpy_frontend import ScatteringNumPy1D
    x = np.random.RandomState(1).randn(2, 1024)
    S = Scattering1D(J=4, shape=1024, Q=2, T=0, out_type="list", backend="torch_b200").cuda()
    y = S(torch.from_numpy(x).float().cuda())
    ref = ScatteringNumPy1D(J=4, shape=1024, Q=2, T=0, out_type="list")(x)
    assert len(y) == len(ref)
    for a, b in zip(y, ref):
        assert np.abs(a["coef"].cpu().numpy() - b["coef"]).max() <= 1e-4 * np.abs(b["coef"]).max() + 1e-7


@pytest.mark.parametrize("kw", [dict(J=4, shape=1024, Q=2), dict(J=5, shape=2048, Q=(4, 1), T=16),
                                dict(J=3, shape=500, Q=1, max_order=1)])
def test_1d_gradients_match_reference_autograd(plugin, kw):
    """Gradients through backend='torch_b200' (this library's kernels and hand-written adjoints, ops_eager.py) against
    the reference torch backend's autograd (tests/scattering1d/test_torch_scattering1d.py:300-312 checks only that a
    gradient exists)."""
    from kymatio.torch import Scattering1D
    from kymatio_b200 import _lib
    torch.manual_seed(0)
    Sb = Scattering1D(backend="torch_b200", **kw).cuda()
    Sr = Scattering1D(backend="torch", **kw).cuda().double()
    x = torch.randn(2, kw["shape"], device="cuda")
    xb = x.clone().requires_grad_(True)
    _lib.timing_enable(True)
    yb = Sb(xb)
    w = torch.randn_like(yb)
    (yb * w).sum().backward()
    labels = {r["label"] for r in _lib.timing_report()}
    _lib.timing_enable(False)
    assert "prim_modulus" in labels or any(l.startswith("prim") for l in labels)      # own kernels, not torch ops
    xr = x.double().requires_grad_(True)
    (Sr(xr) * w.double()).sum().backward()
    assert xb.grad is not None and torch.isfinite(xb.grad).all()
    err = (xb.grad.double() - xr.grad).abs().max() / xr.grad.abs().max()
    assert err < 1e-4, float(err)
    with torch.no_grad():                                      # the no-grad call still takes the fused schedule
        assert (Sb(x) - yb.detach()).abs().max() <= 1e-5 * yb.detach().abs().max()


def test_jtfs_through_eager_primitives(plugin):
    """Joint time-frequency scattering (out of the fused scope) still runs through backend='torch_b200':
    the unchanged core (kymatio/scattering1d/core/timefrequency_scattering.py) drives this library's eager primitives
    plus the reshape helpers (pad_frequency, swap_time_frequency, ...)."""
    from kymatio.torch import TimeFrequencyScattering
    from kymatio.numpy import TimeFrequencyScattering as TimeFrequencyScatteringNumPy
    kw = dict(J=5, J_fr=3, Q=4, shape=1024, format="time")
    x = np.random.RandomState(2).randn(2, 1024)
    S = TimeFrequencyScattering(backend="torch_b200", **kw).cuda()
    with torch.no_grad():
        y = S(torch.from_numpy(x).float().cuda())
    ref = TimeFrequencyScatteringNumPy(**kw)(x)
    assert tuple(y.shape) == ref.shape
    assert np.abs(y.cpu().numpy() - ref).max() <= 1e-4 * np.abs(ref).max()


def test_fused1d_batch_shapes(plugin):
    """Leading batch axes are flattened by the frontend and restored on output (reshape_input / reshape_output); a single
    un-batched signal works too."""
    from kymatio.torch import Scattering1D
    S = Scattering1D(J=4, shape=1024, Q=(4, 1), backend="torch_b200").cuda()
    x = torch.randn(2, 3, 1024, device="cuda")
    y = S(x)
    flat = S(x.reshape(6, 1024))
    assert y.shape[:2] == (2, 3) and torch.equal(y.reshape(flat.shape), flat)
    one = S(x[0, 0])
    assert one.shape == flat.shape[1:] and torch.allclose(one, flat[0], rtol=0, atol=1e-6 * float(flat.abs().max()))


def test_fused1d_streams_rebinding_and_float64(plugin):
    from kymatio.torch import Scattering1D
    from kymatio.scattering1d.frontend.numpy_frontend import ScatteringNumPy1D
    kw = dict(J=4, shape=1024, Q=(4, 1))
    x = np.random.RandomState(6).randn(3, 1024)
    ref = ScatteringNumPy1D(**kw)(x)
    S = Scattering1D(backend="torch_b200", **kw).cuda()
    xt = torch.from_numpy(x).float().cuda()
    y0 = S(xt)
    # a side stream: launches go to torch's current stream, workspaces are stream-ordered
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        y1 = S(xt)
    side.synchronize()
    assert torch.equal(y0, y1)
    # moving the module re-creates its buffers: the engine is re-bound (keyed on the buffers' identity)
    S = S.cpu().cuda()
    assert torch.equal(S(xt), y0)
    assert_parity(y0.cpu().numpy(), ref, channel_axis=-2, what="1d fused")
    # float64 modules are outside the fused kernels: the unchanged core drives the eager float64 primitives
    Sd = Scattering1D(backend="torch_b200", **kw).cuda().double()
    yd = Sd(torch.from_numpy(x).cuda())
    assert yd.dtype == torch.float64
    # (the module's filters were rounded to float32 at registration, kymatio/scattering1d/frontend/torch_frontend.py:36-40)
    assert np.abs(yd.cpu().numpy() - ref).max() <= 1e-5 * np.abs(ref).max()


@pytest.mark.parametrize("kw", [dict(J=6, shape=4096, Q=(8, 1), T="global"), dict(J=5, shape=3000, Q=(4, 2), T="global"),
                                dict(J=7, shape=2 ** 14, Q=(8, 1), T="global", max_order=1)])
def test_fused1d_average_global(plugin, kw):
    """average='global' (kymatio/scattering1d/frontend/base_frontend.py:137-138, core/scattering1d.py with
    average_local=False) through the fused schedule: the sum over time of every path is bin 0 of its spectrum
    (scat1d_finish_global); against the reference's numpy frontend in float64."""
    from kymatio.torch import Scattering1D
    from kymatio.scattering1d.frontend.numpy_frontend import ScatteringNumPy1D
    from kymatio_b200 import _lib
    x = np.random.RandomState(3).randn(3, kw["shape"])
    S = Scattering1D(backend="torch_b200", **kw).cuda()
    _lib.timing_enable(True)
    y = S(torch.from_numpy(x).float().cuda())
    labels = {r["label"].split(":")[0] for r in _lib.timing_report()}
    _lib.timing_enable(False)
    assert "1d_finish_global" in labels, labels
    ref = ScatteringNumPy1D(**kw)(x)
    assert tuple(y.shape) == ref.shape == (3, ref.shape[1], 1)
    assert_parity(y.cpu().numpy(), ref, channel_axis=-2, what=str(kw))


@pytest.mark.parametrize("kw", [dict(J=4, shape=1024, Q=(4, 1), T=0, out_type="list"),
                                dict(J=6, shape=2 ** 15, Q=(8, 1), T=0, out_type="dict"),
                                dict(J=5, shape=3000, Q=(4, 2), T=0, out_type="list", max_order=1)])
def test_fused1d_T0_unaveraged(plugin, kw):
    """T=0 (core/scattering1d.py:75-76,104-105: the modulus field of every path at its own resolution, unpadded by the
    frontend) through the fused kernels (scat1d_tile_t0 / scat1d_row_mod_t0 store |u| in natural time order); against the
    reference's numpy frontend in float64, path by path."""
    from kymatio.torch import Scattering1D
    from kymatio.scattering1d.frontend.numpy_frontend import ScatteringNumPy1D
    from kymatio_b200 import _lib
    x = np.random.RandomState(4).randn(2, kw["shape"])
    S = Scattering1D(backend="torch_b200", **kw).cuda()
    _lib.timing_enable(True)
    y = S(torch.from_numpy(x).float().cuda())
    labels = {r["label"].split(":")[0] for r in _lib.timing_report()}
    _lib.timing_enable(False)
    assert labels & {"1d_tile_t0", "1d_row_mod_t0"}, labels
    assert not any(l.startswith("prim_") for l in labels if l not in ("prim_pad1d",)), labels
    ref = ScatteringNumPy1D(**kw)(x)
    if kw["out_type"] == "dict":
        assert set(y.keys()) == set(ref.keys())
        pairs = [(y[k], ref[k]) for k in ref]
    else:
        assert [a["n"] for a in y] == [b["n"] for b in ref]
        pairs = [(a["coef"], b["coef"]) for a, b in zip(y, ref)]
    for a, b in pairs:
        assert tuple(a.shape) == b.shape
        assert np.abs(a.cpu().numpy() - b).max() <= 1e-4 * max(np.abs(b).max(), 1e-12)
