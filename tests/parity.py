"""Parity metrics shared by the tests, smoke() and bench.py.

north_star tolerance: coefficients within max relative error 1e-4 in fp32 of the
reference's numpy backend.  Element-wise relative error is ill-defined on near-zero
coefficients (the reference's own fp32 vs fp64 runs differ by 5e-4 there), so the gate is
  (1) max|a-b| / max|b|            <= 1e-4   over the tensor
  (2) per-channel relative L2 norm  <= 1e-4
"""
import numpy as np

TOL = 1e-4


def parity_report(a, b, channel_axis=-3):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    max_rel = float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))
    ax = channel_axis % a.ndim
    red = tuple(i for i in range(a.ndim) if i != ax)
    num = np.sqrt(((a - b) ** 2).sum(axis=red))
    den = np.sqrt((b ** 2).sum(axis=red))
    ch = num / np.maximum(den, 1e-300)
    return {"max_rel": max_rel, "chan_l2_max": float(ch.max()), "chan_argmax": int(ch.argmax())}


def assert_parity(a, b, tol=TOL, channel_axis=-3, what=""):
    r = parity_report(a, b, channel_axis)
    assert r["max_rel"] <= tol and r["chan_l2_max"] <= tol, (what, r)
    return r
