"""Differentiable versions of the eager 1-D / 3-D backend primitives: every ``torch.autograd.Function`` pairs a
forward kernel of this library with its hand-written adjoint.  torch records the tape (and applies
the scalar 1/N or N of the FFT adjoints).

Used by ``kymatio_plugin`` when a gradient is requested through ``backend='torch_b200'`` Scattering1D /
HarmonicScattering3D: the unchanged reference core then drives these ops (the fused forward-only schedules of
engine1d.py / engine3d.py serve the no-grad case).  Tensors use the torch backend's layout (real: trailing axis 1,
complex: trailing axis 2), contiguous, CUDA.
"""
import ctypes

import torch

from . import _lib
from .engine2d import _DTYPES
from .ops2d import Modulus as _Modulus2          # shape-agnostic |z| with the ModulusStable adjoint


def _st(t):
    return ctypes.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def _code(t):
    return _DTYPES[t.dtype]


def _call(fn, ref, *args):
    with torch.cuda.device(ref.device):
        _lib.check(fn(*args))


class FromReal(torch.autograd.Function):
    """(..., 1) real -> (..., 2) complex with zero imaginary part; adjoint = real part."""

    @staticmethod
    def forward(ctx, x):
        x = x.contiguous()
        out = torch.empty(x.shape[:-1] + (2,), dtype=x.dtype, device=x.device)
        _call(_lib.load().scat_complex_from_real, x, x.data_ptr(), out.data_ptr(), x.numel(), _code(x), _st(x))
        return out

    @staticmethod
    def backward(ctx, g):
        g = g.contiguous()
        out = torch.empty(g.shape[:-1] + (1,), dtype=g.dtype, device=g.device)
        _call(_lib.load().scat_real_part, g, g.data_ptr(), out.data_ptr(), out.numel(), _code(g), _st(g))
        return out


class RealPart(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z):
        z = z.contiguous()
        out = torch.empty(z.shape[:-1] + (1,), dtype=z.dtype, device=z.device)
        _call(_lib.load().scat_real_part, z, z.data_ptr(), out.data_ptr(), out.numel(), _code(z), _st(z))
        return out

    @staticmethod
    def backward(ctx, g):
        g = g.contiguous()
        out = torch.empty(g.shape[:-1] + (2,), dtype=g.dtype, device=g.device)
        _call(_lib.load().scat_complex_from_real, g, g.data_ptr(), out.data_ptr(), g.numel(), _code(g), _st(g))
        return out


class FftN(torch.autograd.Function):
    """Unnormalised forward / (1/N)-normalised inverse transform through ``exec_fn(x, inverse) -> tensor``;
    adjoint = the opposite transform (y = F x -> gx = N ifft(g); y = ifft x -> gx = fft(g) / N)."""

    @staticmethod
    def forward(ctx, x, inverse, exec_fn, n_total):
        ctx.inverse, ctx.exec_fn, ctx.n_total = inverse, exec_fn, n_total
        return exec_fn(x.contiguous(), inverse)

    @staticmethod
    def backward(ctx, g):
        g = g.contiguous()
        if ctx.inverse:
            return ctx.exec_fn(g, False) / ctx.n_total, None, None, None
        return ctx.exec_fn(g, True) * ctx.n_total, None, None, None


class Cdgmm(torch.autograd.Function):
    """A[..., n, 2] * B[n, 1 or 2] (kymatio/backend/torch_backend.py:148-219); adjoint w.r.t. A = multiply by conj(B)."""

    @staticmethod
    def forward(ctx, A, B):
        A = A.contiguous()
        n = B.numel() // B.shape[-1]
        ctx.n, ctx.cplx = n, B.shape[-1] == 2
        ctx.save_for_backward(B)
        out = torch.empty_like(A)
        _call(_lib.load().scat_cdgmm, A, A.data_ptr(), B.data_ptr(), out.data_ptr(), A.numel() // 2 // n, n,
              int(ctx.cplx), _code(A), _st(A))
        return out

    @staticmethod
    def backward(ctx, g):
        (B,) = ctx.saved_tensors
        g = g.contiguous()
        out = torch.empty_like(g)
        _call(_lib.load().scat_cdgmm, g, g.data_ptr(), B.data_ptr(), out.data_ptr(), g.numel() // 2 // ctx.n, ctx.n,
              2 if ctx.cplx else 0, _code(g), _st(g))
        return out, None


class SubsampleFourier1d(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, k):
        x = x.contiguous()
        N = x.shape[-2]
        ctx.k, ctx.shape = k, x.shape
        out = torch.empty(x.shape[:-2] + (N // k, 2), dtype=x.dtype, device=x.device)
        _call(_lib.load().scat_subsample_fourier1d, x, x.data_ptr(), out.data_ptr(), x.numel() // (2 * N), N, int(k),
              _code(x), _st(x))
        return out

    @staticmethod
    def backward(ctx, g):
        g = g.contiguous()
        N = ctx.shape[-2]
        gin = torch.empty(ctx.shape, dtype=g.dtype, device=g.device)
        _call(_lib.load().scat_subsample_fourier1d_bwd, g, g.data_ptr(), gin.data_ptr(), gin.numel() // (2 * N), N,
              int(ctx.k), _code(g), _st(g))
        return gin, None


class Pad1d(torch.autograd.Function):
    """Reflect padding along the last axis; adjoint = fold-add (the (B, 1, N) case of the 2-D kernel)."""

    @staticmethod
    def forward(ctx, x, pad_left, pad_right):
        x = x.contiguous()
        N = x.shape[-1]
        ctx.pads, ctx.shape = (int(pad_left), int(pad_right)), x.shape
        out = torch.empty(x.shape[:-1] + (N + pad_left + pad_right,), dtype=x.dtype, device=x.device)
        _call(_lib.load().scat_pad1d, x, x.data_ptr(), out.data_ptr(), x.numel() // N, N, int(pad_left), int(pad_right),
              _code(x), _st(x))
        return out

    @staticmethod
    def backward(ctx, g):
        g = g.contiguous()
        N = ctx.shape[-1]
        gx = torch.empty(ctx.shape, dtype=g.dtype, device=g.device)
        _call(_lib.load().scat_pad2d_bwd, g, g.data_ptr(), gx.data_ptr(), gx.numel() // N, 1, N, 0, 0, ctx.pads[0],
              ctx.pads[1], _code(g), _st(g))
        return gx, None, None


class ModulusRotation(torch.autograd.Function):
    """sqrt(prev^2 + |x|^2) (kymatio/scattering3d/backend/torch_backend.py:102-124); prev may be None."""

    @staticmethod
    def forward(ctx, x, prev):
        x = x.contiguous()
        prev_c = None if prev is None else prev.contiguous()
        out = torch.empty(x.shape[:-1] + (1,), dtype=x.dtype, device=x.device)
        _call(_lib.load().scat_modulus_rotation, x, x.data_ptr(), None if prev_c is None else prev_c.data_ptr(),
              out.data_ptr(), out.numel(), _code(x), _st(x))
        ctx.has_prev = prev_c is not None
        ctx.save_for_backward(x, out) if prev_c is None else ctx.save_for_backward(x, out, prev_c)
        return out

    @staticmethod
    def backward(ctx, g):
        saved = ctx.saved_tensors
        x, out = saved[0], saved[1]
        prev = saved[2] if ctx.has_prev else None
        g = g.contiguous()
        gx = torch.empty_like(x)
        gprev = torch.empty_like(out) if prev is not None else None
        _call(_lib.load().scat_modulus_rotation_bwd, x, x.data_ptr(), None if prev is None else prev.data_ptr(),
              out.data_ptr(), g.data_ptr(), gx.data_ptr(), None if gprev is None else gprev.data_ptr(), out.numel(),
              _code(x), _st(x))
        return gx, gprev


class ComputeIntegrals(torch.autograd.Function):
    """out[b][p] = sum_voxels x^q_p (kymatio/scattering3d/backend/torch_backend.py:127-151), result in torch's default
    dtype like the reference."""

    @staticmethod
    def forward(ctx, x, powers):
        x = x.contiguous()
        B = x.shape[0]
        pw = torch.tensor([float(q) for q in powers], dtype=torch.float32, device=x.device)
        acc = torch.zeros((B, len(powers)), dtype=torch.float64, device=x.device)
        _call(_lib.load().scat_compute_integrals, x, x.data_ptr(), acc.data_ptr(), B, x.numel() // max(B, 1), pw.data_ptr(),
              len(powers), _code(x), _st(x))
        ctx.save_for_backward(x, pw)
        return acc.to(torch.get_default_dtype())

    @staticmethod
    def backward(ctx, g):
        x, pw = ctx.saved_tensors
        g = g.to(x.dtype).contiguous()
        gx = torch.empty_like(x)
        B = x.shape[0]
        _call(_lib.load().scat_compute_integrals_bwd, x, x.data_ptr(), g.data_ptr(), gx.data_ptr(), B,
              x.numel() // max(B, 1), pw.data_ptr(), pw.numel(), _code(x), _st(x))
        return gx, None


def modulus(x):
    """(..., 2) -> (..., 1) with the reference's ModulusStable adjoint (kymatio/backend/torch_backend.py:64-96)."""
    return _Modulus2.apply(x)[..., None]
