"""Autograd glue for the fused 2D engine.

The reference gets gradients by replaying torch autograd over ~2k recorded primitive ops,
with one custom backward for the modulus (kymatio/backend/torch_backend.py:64-96).  Here
the whole forward is one opaque call, so the backward is one call too (mirrored cascade,
SURVEY Appendix B).
"""
import torch


class _Scattering2DFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, eng):
        ctx.eng = eng
        ctx.save_for_backward(x)
        # the first-order spectra are kept for the backward when they fit (Engine2D.saved_u1_buffers)
        out, ctx.saved_u1 = eng.forward_saving(x)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        (x,) = ctx.saved_tensors
        eng = ctx.eng
        saved, ctx.saved_u1 = ctx.saved_u1, None
        return eng.backward(x, grad_out.contiguous(), saved), None


def scattering2d_apply(eng, x):
    if torch.is_grad_enabled() and x.requires_grad:
        return _Scattering2DFn.apply(x, eng)
    return eng.forward(x)
