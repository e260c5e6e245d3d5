"""Filter banks synthesised on the GPU (constructor path).

The reference builds its 2-D Morlet bank with Python loops over 25 periods per Gabor and an O(M N 4^res) quadruple loop
per level (kymatio/scattering2d/filter_bank.py:5-53, :56-91, :131-175) - seconds at 272x272 - and its 3-D solid-harmonic
bank with whole-volume numpy expressions (kymatio/scattering3d/filter_bank.py:5-166).  Here the same filters come from
a handful of kernel launches (csrc/filters.cuh) and stay on the device; containers and layouts are the reference's.
The per-filter scalars (rotation entries rounded to float32, the 3.1415 of the normalisation) follow the reference's
arithmetic; the per-pixel work is done in float64 and rounded once, so the result agrees with the reference's filters to
float32 rounding (tests/test_filters_gpu.py, against reference-generated fixtures).
"""
import math

import numpy as np
import torch

from . import _lib

__all__ = ["filter_bank_2d_gpu", "solid_harmonic_filter_bank_gpu", "gaussian_filter_bank_gpu"]


def _st(dev):
    import ctypes
    return ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def _gabor_params(sigma, theta, xi, slant, zero_mean):
    """Scalars of one filter, computed as the reference does (rotation matrices in float32, filter_bank.py:159-163,172)."""
    ct, st = np.cos(theta), np.sin(theta)
    rot = np.array([[ct, -st], [st, ct]], np.float32)
    rot_inv = np.array([[ct, st], [-st, ct]], np.float32)
    shape_mat = np.array([[1, 0], [0, slant * slant]])
    curv = rot.dot(shape_mat.dot(rot_inv)) / (2 * sigma * sigma)
    norm = 2 * 3.1415 * sigma * sigma / slant
    return [float(curv[0, 0]), float(curv[0, 1] + curv[1, 0]), float(curv[1, 1]), float(xi * ct), float(xi * st),
            1.0 / norm, 1.0 if zero_mean else 0.0, 0.0]


def filter_bank_2d_gpu(Mp, Np, J, L=8, device=None, as_numpy=False, max_bytes=1 << 30):
    """The reference's ``filter_bank(M, N, J, L)`` (kymatio/scattering2d/filter_bank.py:5-53) built on ``device``.

    Returns ``{'phi': {'j': J, 'levels': [...]}, 'psi': [{'j', 'theta', 'levels': [...]}, ...]}`` with float32 CUDA tensors
    (numpy arrays with ``as_numpy``) of shape ``(Mp/2^res, Np/2^res)``.
    """
    from .ops2d import _fft2_raw
    if not torch.cuda.is_available():
        raise RuntimeError("filter_bank_2d_gpu needs a CUDA device")
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    lib = _lib.load()
    Mp, Np, J, L = int(Mp), int(Np), int(J), int(L)
    specs = []           # (params, number of levels)
    for j in range(J):
        for theta in range(L):
            specs.append((_gabor_params(0.8 * 2 ** j, (int(L - L / 2 - 1) - theta) * np.pi / L, 3.0 / 4.0 * np.pi / 2 ** j,
                                        4.0 / L, True), min(j + 1, max(J - 1, 1))))
    specs.append((_gabor_params(0.8 * 2 ** (J - 1), 0, 0, 1.0, False), J))
    levels = []
    per_filter = Mp * Np * (16 + 16 + 8)           # spatial + spectrum + envelope
    step = max(1, min(len(specs), max_bytes // per_filter))
    with torch.cuda.device(dev):
        for f0 in range(0, len(specs), step):
            chunk = specs[f0:f0 + step]
            F = len(chunk)
            prm = torch.tensor([c[0] for c in chunk], dtype=torch.float64).to(dev)
            spatial = torch.empty((F, Mp, Np, 2), dtype=torch.float64, device=dev)
            env = torch.empty((F, Mp, Np), dtype=torch.float64, device=dev)
            sums = torch.empty((F, 3), dtype=torch.float64, device=dev)
            _lib.check(lib.scat_filters2d_spatial(prm.data_ptr(), F, Mp, Np, spatial.data_ptr(), env.data_ptr(),
                                                  sums.data_ptr(), _st(dev)))
            spec = _fft2_raw(spatial, False)
            for i, (_, nlev) in enumerate(chunk):
                per = []
                for res in range(nlev):
                    out = torch.empty((Mp >> res, Np >> res), dtype=torch.float32, device=dev)
                    _lib.check(lib.scat_filters2d_fold(spec[i].data_ptr(), out.data_ptr(), Mp, Np, res, _st(dev)))
                    per.append(out)
                levels.append(per)
    if as_numpy:
        levels = [[t.cpu().numpy() for t in per] for per in levels]
    psi = [{"j": n // L, "theta": n % L, "levels": levels[n]} for n in range(J * L)]
    return {"phi": {"j": J, "levels": levels[J * L]}, "psi": psi}


def _solid_harmonic_norm(l):
    """Real normalisation of the order-l wavelets in the Fourier domain (kymatio/scattering3d/filter_bank.py:153-163);
    the factor (-i)^l is applied by the kernel."""
    if l % 2 == 0:
        dfact = 1.0
        for i in range(l + 1, 0, -2):
            dfact *= i
        c = 1.0 / (2 * math.pi * math.sqrt(l + 0.5) * dfact)
    else:
        c = 1.0 / (2 ** (0.5 * (l + 3)) * math.sqrt(math.pi * (2 * l + 1)) * math.factorial((l + 1) // 2))
    return c * (2 * math.pi) ** 1.5


def solid_harmonic_filter_bank_gpu(M, N, O, J, L, sigma_0, device=None, as_numpy=False):
    """``solid_harmonic_filter_bank(M, N, O, J, L, sigma_0, fourier=True)`` on the device: a list over l = 0..L of
    ``(J+1, 2l+1, M, N, O)`` complex64 arrays (as complex tensors, or numpy arrays with ``as_numpy``)."""
    if not torch.cuda.is_available():
        raise RuntimeError("solid_harmonic_filter_bank_gpu needs a CUDA device")
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    lib = _lib.load()
    out = []
    with torch.cuda.device(dev):
        sig = torch.tensor([sigma_0 * 2 ** j for j in range(J + 1)], dtype=torch.float64).to(dev)
        for l in range(L + 1):
            buf = torch.empty((J + 1, 2 * l + 1, M, N, O, 2), dtype=torch.float32, device=dev)
            _lib.check(lib.scat_filters3d_solid_harmonic(buf.data_ptr(), sig.data_ptr(), J + 1, l, _solid_harmonic_norm(l),
                                                         int(M), int(N), int(O), _st(dev)))
            t = torch.view_as_complex(buf)
            out.append(t.cpu().numpy() if as_numpy else t)
    return out


def gaussian_filter_bank_gpu(M, N, O, J, sigma_0, device=None, as_numpy=False):
    """``gaussian_filter_bank(M, N, O, J, sigma_0, fourier=True)`` on the device: ``(J+1, M, N, O)`` complex64."""
    if not torch.cuda.is_available():
        raise RuntimeError("gaussian_filter_bank_gpu needs a CUDA device")
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    lib = _lib.load()
    with torch.cuda.device(dev):
        sig = torch.tensor([sigma_0 * 2 ** j for j in range(J + 1)], dtype=torch.float64).to(dev)
        buf = torch.empty((J + 1, M, N, O, 2), dtype=torch.float32, device=dev)
        _lib.check(lib.scat_filters3d_gaussian(buf.data_ptr(), sig.data_ptr(), J + 1, int(M), int(N), int(O), _st(dev)))
    t = torch.view_as_complex(buf)
    return t.cpu().numpy() if as_numpy else t
