"""Differentiable building blocks over the library's own kernels, and the angle-batched eager cascade
used to back-propagate through the fused forward (recomputed in ``backward``).

Each ``torch.autograd.Function`` pairs a forward kernel with its hand-written adjoint (SURVEY
Appendix B): FFT <-> the opposite FFT, filter multiply <-> filter multiply-accumulate over the filter
axis, periodisation <-> replication / k^2, modulus <-> ``x g / |x|`` (0 at 0, the reference's
``ModulusStable.backward``, kymatio/backend/torch_backend.py:85-96), reflect pad <-> fold-add.
Complex tensors use the torch backend's layout (trailing axis of size 2); everything is contiguous.
"""
import ctypes
import threading

import torch

from . import _lib
from .engine2d import _DTYPES


def _st(t):
    return ctypes.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def _code(t):
    return _DTYPES[t.dtype]


_tables = {}


def _fft_tables(n0, n1, ref):
    key = (n0, n1, ref.dtype, ref.device.index)
    buf = _tables.get(key)
    if buf is None:
        lib = _lib.load()
        nbytes = lib.scat_fft2d_const_bytes(n0, n1, _code(ref))
        if nbytes == 0:
            raise _lib.ScatB200Error(lib.scat_last_error().decode())
        with torch.cuda.device(ref.device):
            buf = torch.empty(nbytes, dtype=torch.uint8, device=ref.device)
            _lib.check(lib.scat_fft2d_init(buf.data_ptr(), n0, n1, _code(ref), _st(ref)))
        _tables[key] = buf
    return buf


def _fft2_raw(x, inverse):
    """x: (..., n0, n1, 2) contiguous -> same shape; unnormalised forward / (1/N) inverse."""
    n0, n1 = x.shape[-3], x.shape[-2]
    out = torch.empty_like(x)
    G = x.numel() // (n0 * n1 * 2)
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().scat_fft2d_exec(_fft_tables(n0, n1, x).data_ptr(), x.data_ptr(), out.data_ptr(), G,
                                               n0, n1, int(inverse), _code(x), _st(x)))
    return out


class Fft2(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, inverse):
        ctx.inverse = inverse
        return _fft2_raw(x.contiguous(), inverse)

    @staticmethod
    def backward(ctx, g):
        g = g.contiguous()
        N = g.shape[-3] * g.shape[-2]
        if int(ctx.inverse) == 2:            # y = F^H x        ->  gx = F g
            return _fft2_raw(g, False), None
        if ctx.inverse:                      # y = (1/N) F^H x  ->  gx = (1/N) F g
            return _fft2_raw(g, False) / N, None
        return _fft2_raw(g, 2), None         # y = F x        ->  gx = F^H g (unnormalised inverse, mode 2)


class Modulus(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        x = x.contiguous()
        out = torch.empty(x.shape[:-1], dtype=x.dtype, device=x.device)
        with torch.cuda.device(x.device):
            _lib.check(_lib.load().scat_modulus(x.data_ptr(), out.data_ptr(), out.numel(), _code(x), _st(x)))
        ctx.save_for_backward(x)
        return out

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        g = g.contiguous()
        gx = torch.empty_like(x)
        with torch.cuda.device(x.device):
            _lib.check(_lib.load().scat_modulus_bwd(x.data_ptr(), g.data_ptr(), gx.data_ptr(), g.numel(), _code(x),
                                                    _st(x)))
        return gx


class FilterBank(torch.autograd.Function):
    """A: (nb, n0, n1, 2), W: (nf, n0, n1) real -> (nb, nf, n0, n1, 2)."""

    @staticmethod
    def forward(ctx, A, W):
        A, W = A.contiguous(), W.contiguous()
        nb, nf = A.shape[0], W.shape[0]
        n = A.shape[1] * A.shape[2]
        out = torch.empty((nb, nf) + tuple(A.shape[1:]), dtype=A.dtype, device=A.device)
        with torch.cuda.device(A.device):
            _lib.check(_lib.load().scat_cdgmm_bcast(A.data_ptr(), W.data_ptr(), out.data_ptr(), nb, nf, n, 0,
                                                    _code(A), _st(A)))
        ctx.save_for_backward(W)
        return out

    @staticmethod
    def backward(ctx, g):
        (W,) = ctx.saved_tensors
        g = g.contiguous()
        nb, nf = g.shape[0], g.shape[1]
        n = g.shape[2] * g.shape[3]
        gA = torch.empty((nb,) + tuple(g.shape[2:]), dtype=g.dtype, device=g.device)
        with torch.cuda.device(g.device):
            _lib.check(_lib.load().scat_cdgmm_bcast(g.data_ptr(), W.data_ptr(), gA.data_ptr(), nb, nf, n, 1,
                                                    _code(g), _st(g)))
        return gA, None


class Periodize(torch.autograd.Function):
    """(G, n0, n1, 2) -> (G, n0/k, n1/k, 2): Fourier-domain subsampling by k."""

    @staticmethod
    def forward(ctx, x, k):
        x = x.contiguous()
        ctx.k, ctx.shape = k, x.shape
        if k == 1:
            return x
        G, n0, n1 = x.shape[0], x.shape[1], x.shape[2]
        out = torch.empty((G, n0 // k, n1 // k, 2), dtype=x.dtype, device=x.device)
        with torch.cuda.device(x.device):
            _lib.check(_lib.load().scat_subsample_fourier2d(x.data_ptr(), out.data_ptr(), G, n0, n1, k, _code(x),
                                                            _st(x)))
        return out

    @staticmethod
    def backward(ctx, g):
        if ctx.k == 1:
            return g, None
        g = g.contiguous()
        G, n0, n1, _ = ctx.shape
        gin = torch.empty(ctx.shape, dtype=g.dtype, device=g.device)
        with torch.cuda.device(g.device):
            _lib.check(_lib.load().scat_subsample_fourier2d_bwd(g.data_ptr(), gin.data_ptr(), G, n0, n1, ctx.k,
                                                                _code(g), _st(g)))
        return gin, None


class PadReflect(torch.autograd.Function):
    """(B, M, N) -> (B, M+t+b, N+l+r) reflect padding."""

    @staticmethod
    def forward(ctx, x, pads):
        x = x.contiguous()
        t, b, l, r = pads
        ctx.pads, ctx.shape = pads, x.shape
        B, M, N = x.shape
        out = torch.empty((B, M + t + b, N + l + r), dtype=x.dtype, device=x.device)
        with torch.cuda.device(x.device):
            _lib.check(_lib.load().scat_pad2d(x.data_ptr(), out.data_ptr(), B, M, N, t, b, l, r, _code(x), _st(x)))
        return out

    @staticmethod
    def backward(ctx, g):
        g = g.contiguous()
        t, b, l, r = ctx.pads
        B, M, N = ctx.shape
        gx = torch.empty(ctx.shape, dtype=g.dtype, device=g.device)
        with torch.cuda.device(g.device):
            _lib.check(_lib.load().scat_pad2d_bwd(g.data_ptr(), gx.data_ptr(), B, M, N, t, b, l, r, _code(g), _st(g)))
        return gx, None


class _RecomputeState(threading.local):
    """Set while Engine2D.backward re-runs the cascade only to rebuild the autograd graph: the VALUES of the leaf outputs
    are then never read (their gradients come from the caller), so operators whose backward does not need their own
    forward result skip the forward kernels.  ``saved``: {j1: U1 of that scale for the current chunk} kept by the forward
    (Engine2D.forward_saving) - the first-order blocks then hand it out instead of computing it again.
    Thread-local: autograd runs the backward of different devices (nn.DataParallel replicas) on different threads."""
    active = False
    saved = None


_Recompute = _RecomputeState()


class Order2(torch.autograd.Function):
    """All second-order paths below first-order scale j1, fused: U1 (B*L, n0, n1, 2) -> (B, C2, o0, o1).
    Forward and backward are one tile-kernel launch per (j1, j2) pair (csrc/tile2d.cuh)."""

    @staticmethod
    def forward(ctx, U1, eng, j1, batch):
        U1 = U1.contiguous()
        ctx.eng, ctx.j1, ctx.batch = eng, j1, batch
        ctx.save_for_backward(U1)
        if _Recompute.active:       # graph rebuild: the second-order outputs are not needed, only their gradients are
            return U1.new_empty((batch, eng.order2_channels(j1), eng.out_h, eng.out_w))
        return eng.order2_forward(j1, U1, batch)

    @staticmethod
    def backward(ctx, g):
        (U1,) = ctx.saved_tensors
        return ctx.eng.order2_backward(ctx.j1, U1, g.contiguous(), ctx.batch), None, None, None


class Order1Tile(torch.autograd.Function):
    """First-order block of a scale whose field fits one CTA, fused: U0 (B, Mp, Np, 2) -> S1 (B, L, oh, ow) and, for
    scales with children, U1 (B*L, n0, n1, 2).  Forward = the tile kernel of the forward engine; backward = k2d_tile_adj
    (Re F^H gU1) + the backward tile (recompute u, low-pass adjoint, modulus backward, forward transform, scatter into gU0)."""

    @staticmethod
    def forward(ctx, U0, eng, j1, batch, want_u1):
        U0 = U0.contiguous()
        ctx.eng, ctx.j1, ctx.batch, ctx.want_u1 = eng, j1, batch, want_u1
        ctx.save_for_backward(U0)
        kept = _Recompute.saved.get(j1) if (_Recompute.active and _Recompute.saved) else None
        if _Recompute.active and (not want_u1 or kept is not None):
            # graph rebuild: S1 is never read, U1 was kept by the forward (or is not needed) - no kernel runs
            s1 = U0.new_empty((batch, eng.geometry["L"], eng.out_h, eng.out_w))
            u1 = kept
        else:
            s1, u1 = eng.order1_forward(j1, U0, batch, want_u1)
        if not want_u1:
            u1 = U0.new_zeros((0,))
        ctx.mark_non_differentiable(*(() if want_u1 else (u1,)))
        return s1, u1

    @staticmethod
    def backward(ctx, gs1, gu1):
        (U0,) = ctx.saved_tensors
        gs1 = gs1.contiguous() if gs1 is not None else torch.zeros(
            (ctx.batch, ctx.eng.geometry["L"], ctx.eng.out_h, ctx.eng.out_w), dtype=U0.dtype, device=U0.device)
        gu1 = gu1.contiguous() if (ctx.want_u1 and gu1 is not None) else None
        return ctx.eng.order1_backward(ctx.j1, U0, gs1, gu1, ctx.batch), None, None, None, None


class Order1Stream(torch.autograd.Function):
    """First-order block at full resolution (streaming chain), fused: U0 -> S1 (B, L, oh, ow) and, for scales with
    children, U1 (B*L, n0, n1, 2).  Backward = mirrored chain: row passes of the recomputed product and of gU1, the
    horizontal half of the low-pass adjoint of gS1, one fused column kernel (vertical half + modulus backward), one row
    kernel that reduces over the angles (csrc/bwd2d.cuh)."""

    @staticmethod
    def forward(ctx, U0, eng, j1, batch, want_u1):
        U0 = U0.contiguous()
        ctx.eng, ctx.j1, ctx.batch, ctx.want_u1 = eng, j1, batch, want_u1
        ctx.save_for_backward(U0)
        ctx.set_materialize_grads(False)
        kept = _Recompute.saved.get(j1) if (_Recompute.active and _Recompute.saved) else None
        if _Recompute.active and (not want_u1 or kept is not None):
            # graph rebuild: S1 is never read, U1 was kept by the forward (or is not needed) - no kernel runs
            s1 = U0.new_empty((batch, eng.geometry["L"], eng.out_h, eng.out_w))
            u1 = kept if want_u1 else U0.new_zeros((0,))
        else:
            s1, u1 = eng.order1_forward(j1, U0, batch, True, want_s1=not _Recompute.active)
            if s1 is None:
                s1 = U0.new_empty((batch, eng.geometry["L"], eng.out_h, eng.out_w))
            if not want_u1:
                u1 = U0.new_zeros((0,))
        ctx.mark_non_differentiable(*(() if want_u1 else (u1,)))
        return s1, u1

    @staticmethod
    def backward(ctx, gs1, gu1):
        (U0,) = ctx.saved_tensors
        gs1 = gs1.contiguous() if gs1 is not None else None
        gu1 = gu1.contiguous() if (ctx.want_u1 and gu1 is not None) else None
        if gs1 is None and gu1 is None:
            return None, None, None, None, None
        return ctx.eng.order1_backward(ctx.j1, U0, gs1, gu1, ctx.batch), None, None, None, None


def _to_complex(x):
    from .ops_eager import FromReal          # kernel + adjoint (real part); no zero tensor, no stack copy
    return FromReal.apply(x.contiguous()[..., None])


def _low(U, phi_level, k):
    """unpad(Re ifft2(periodise_k(U * phi)))  - core/scattering2d.py:19-23.  U: (G, n0, n1, 2)."""
    Z = Periodize.apply(FilterBank.apply(U, phi_level[None])[:, 0], k)
    return Fft2.apply(Z, True)[..., 1:-1, 1:-1, 0]


class LowPassLeaf(torch.autograd.Function):
    """``_low`` as ONE graph node for an output that nothing else consumes (S0): the forward is skipped while the graph is
    only being rebuilt (``_Recompute``), the backward is the hand-written adjoint chain - zero-embed, forward transform / N,
    replicate / k^2 (adjoint of the periodisation), multiply by phi."""

    @staticmethod
    def forward(ctx, U, phi_level, k):
        U = U.contiguous()
        ctx.k, ctx.shape = k, U.shape
        ctx.save_for_backward(phi_level)
        G, n0, n1 = U.shape[0], U.shape[1], U.shape[2]
        if _Recompute.active:
            return U.new_empty((G, n0 // k - 2, n1 // k - 2))
        with torch.no_grad():
            return _low(U, phi_level, k).contiguous()

    @staticmethod
    def backward(ctx, g):
        (phi_level,) = ctx.saved_tensors
        G, n0, n1, _ = ctx.shape
        k, m0, m1 = ctx.k, n0 // ctx.k, n1 // ctx.k
        lib = _lib.load()
        gz = g.new_zeros((G, m0, m1, 2))
        gz[:, 1:-1, 1:-1, 0] = g
        gZ = _fft2_raw(gz, False)
        gZ.mul_(1.0 / (m0 * m1))
        gV = torch.empty(ctx.shape, dtype=g.dtype, device=g.device)
        gU = torch.empty(ctx.shape, dtype=g.dtype, device=g.device)
        with torch.cuda.device(g.device):
            if k > 1:
                _lib.check(lib.scat_subsample_fourier2d_bwd(gZ.data_ptr(), gV.data_ptr(), G, n0, n1, k, _code(g), _st(g)))
            else:
                gV = gZ
            W = phi_level.contiguous()
            _lib.check(lib.scat_cdgmm_bcast(gV.data_ptr(), W.data_ptr(), gU.data_ptr(), G, 1, n0 * n1, 1, _code(g), _st(g)))
        return gU, None, None


def _order2_per_op(U1, psi, phi, j1, J, L, B):
    """All second-order paths below first-order scale j1 on the per-primitive differentiable ops (core:55-83)."""
    per_j2 = []
    for j2 in range(j1 + 1, J):
        V2 = FilterBank.apply(U1, psi[j2][j1])
        V2 = Periodize.apply(V2.reshape((B * L * L,) + tuple(V2.shape[2:])), 2 ** (j2 - j1))
        A2 = Modulus.apply(Fft2.apply(V2, True))
        U2 = Fft2.apply(_to_complex(A2), False)
        s2 = _low(U2, phi[j2], 2 ** (J - j2))
        per_j2.append(s2.reshape((B, L, 1, L) + tuple(s2.shape[1:])))
    s2 = torch.cat(per_j2, dim=2)                           # (B, theta1, j2, theta2, o0, o1)
    return s2.reshape((B, -1) + tuple(s2.shape[4:]))


def eager_scattering2d(x, J, L, max_order, pads, phi_levels, psi_levels, eng=None):
    """Angle-batched restatement of kymatio/scattering2d/core/scattering2d.py:14-86 on differentiable ops.

    x: (B, M, N) CUDA; pads: (top, bottom, left, right) or None when pre-padded;
    phi_levels: J tensors (n0, n1[, 1]); psi_levels: flattened in registration order.
    Returns (B, K, M/2^J, N/2^J) in the reference's channel order.
    With ``eng`` (an Engine2D bound to the same filters) the first-order block of every scale (Order1Tile / Order1Stream)
    and the second-order block below it (Order2) run as fused kernels, forward and backward; what remains on the
    per-primitive graph acts on ONE field per image (pad, U0, S0).
    """
    B = x.shape[0]
    phi = [p.reshape(p.shape[0], p.shape[1]) for p in phi_levels]
    # psi[j][res] -> (L, n0, n1)
    psi, n = [], 0
    n_levels = [min(j + 1, max(J - 1, 1)) for j in range(J)]
    for j in range(J):
        per_theta = []
        for _ in range(L):
            per_theta.append([psi_levels[n + r].reshape(psi_levels[n + r].shape[0], psi_levels[n + r].shape[1])
                              for r in range(n_levels[j])])
            n += n_levels[j]
        psi.append([torch.stack([per_theta[t][r] for t in range(L)]) for r in range(n_levels[j])])

    xp = PadReflect.apply(x, tuple(pads)) if pads is not None else x
    U0 = Fft2.apply(_to_complex(xp), False)
    S0 = LowPassLeaf.apply(U0, phi[0], 2 ** J)[:, None]
    S1, S2 = [], []
    for j1 in range(J):
        has_children = max_order >= 2 and j1 < J - 1
        mode = eng.order1_mode(j1) if eng is not None else 0
        if mode:
            # fused first-order block (forward and backward): see Order1Tile / Order1Stream
            if mode == 1:
                s1, U1 = Order1Tile.apply(U0, eng, j1, B, has_children)
            else:
                s1, U1 = Order1Stream.apply(U0, eng, j1, B, has_children)
            S1.append(s1)
            if has_children:
                if eng.order2_channels(j1) > 0:
                    S2.append(Order2.apply(U1, eng, j1, B))
                else:
                    S2.append(_order2_per_op(U1, psi, phi, j1, J, L, B))
            continue
        # the 1/N of the inverse transform is folded into the (small) filter tensor: the transform itself then runs
        # unnormalised in both directions of the graph, without a scaling pass over the full-size fields
        n_j1 = (U0.shape[1] >> j1) * (U0.shape[2] >> j1)
        V = FilterBank.apply(U0, psi[j1][0] * (1.0 / n_j1))
        V = Periodize.apply(V.reshape((B * L,) + tuple(V.shape[2:])), 2 ** j1)
        A = Modulus.apply(Fft2.apply(V, 2))
        U1 = Fft2.apply(_to_complex(A), False)
        s1 = _low(U1, phi[j1], 2 ** (J - j1))
        S1.append(s1.reshape((B, L) + tuple(s1.shape[1:])))
        if max_order < 2 or j1 >= J - 1:
            continue
        if eng is not None and eng.order2_channels(j1) > 0:
            S2.append(Order2.apply(U1, eng, j1, B))         # fused second-order block (forward and backward)
            continue
        S2.append(_order2_per_op(U1, psi, phi, j1, J, L, B))
    return torch.cat([S0] + S1 + S2, dim=1)
