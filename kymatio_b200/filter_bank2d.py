"""Constructor-time Morlet filter bank for the 2D frontend (numpy, CPU, one-off).

Produces the same Fourier-domain filters, at the same resolutions and in the same
container layout, as the reference's ``filter_bank`` (kymatio/scattering2d/
filter_bank.py:5-53): ``phi = {'j': J, 'levels': [res 0..J-1]}`` and ``psi`` a list of
``J*L`` dicts ``{'j', 'theta', 'levels': [res 0..min(j, J-2)]}``, every level a real
float32 ``(Mp/2^res, Np/2^res)`` array.  The arithmetic follows the reference's
precision path (5x5 periodised Gabor accumulated in complex64, normalisation
``2*3.1415*sigma^2/slant`` - filter_bank.py:159-173) so that the filters agree with the
reference's to float32 rounding; checked in tests/test_filter_bank_2d.py.
"""
import functools

import numpy as np
import scipy.fft

__all__ = ["filter_bank_2d", "padded_size_2d"]


def padded_size_2d(M, N, J):
    """kymatio/scattering2d/utils.py:19-22."""
    s = 2 ** J
    return ((M + s) // s + 1) * s, ((N + s) // s + 1) * s


def _periodised_gabor(M, N, sigma, theta, xi, slant):
    """Gaussian envelope times plane wave, summed over the 5x5 neighbouring periods."""
    ct, st = np.cos(theta), np.sin(theta)
    rot = np.array([[ct, -st], [st, ct]], np.float32)
    rot_inv = np.array([[ct, st], [-st, ct]], np.float32)
    shape_mat = np.array([[1, 0], [0, slant * slant]])
    curv = rot.dot(shape_mat.dot(rot_inv)) / (2 * sigma * sigma)
    cross = curv[0, 1] + curv[1, 0]
    out = np.zeros((M, N), np.complex64)
    rows = np.arange(M)[:, None]
    cols = np.arange(N)[None, :]
    for pr in range(-2, 3):
        u = rows + pr * M
        for pc in range(-2, 3):
            v = cols + pc * N
            envelope = -(curv[0, 0] * u * u + cross * u * v + curv[1, 1] * v * v)
            phase = u * xi * ct + v * xi * st
            out += np.exp(envelope + 1.j * phase)
    out /= (2 * 3.1415 * sigma * sigma / slant)
    return out


def _zero_mean_morlet(M, N, sigma, theta, xi, slant):
    carrier = _periodised_gabor(M, N, sigma, theta, xi, slant)
    envelope = _periodised_gabor(M, N, sigma, theta, 0, slant)
    beta = np.sum(carrier) / np.sum(envelope)
    return carrier - beta * envelope


def _fold_to_resolution(f_hat, res):
    """Band-limit to the lowest 1/2^res of the spectrum and fold the aliases
    (filter_bank.py:56-91) - vectorised as a sum over the 2^res x 2^res tiles."""
    M, N = f_hat.shape
    keep = np.ones((M, N), np.float32)
    r0, rl = int(M * 2 ** (-res - 1)), int(M * (1 - 2 ** (-res)))
    c0, cl = int(N * 2 ** (-res - 1)), int(N * (1 - 2 ** (-res)))
    keep[r0:r0 + rl, :] = 0
    keep[:, c0:c0 + cl] = 0
    masked = f_hat * keep
    k = 2 ** res
    m, n = M // k, N // k
    return masked.reshape(k, m, k, n).sum(axis=(0, 2), dtype=masked.dtype)


@functools.lru_cache(maxsize=16)
def _filter_bank_cached(Mp, Np, J, L):
    psi = []
    for j in range(J):
        for theta in range(L):
            w = _zero_mean_morlet(Mp, Np, 0.8 * 2 ** j, (int(L - L / 2 - 1) - theta) * np.pi / L,
                                  3.0 / 4.0 * np.pi / 2 ** j, 4.0 / L)
            w_hat = np.real(scipy.fft.fft2(w))
            n_levels = min(j + 1, max(J - 1, 1))
            psi.append({"j": j, "theta": theta,
                        "levels": [_fold_to_resolution(w_hat, r) for r in range(n_levels)]})
    g_hat = np.real(scipy.fft.fft2(_periodised_gabor(Mp, Np, 0.8 * 2 ** (J - 1), 0, 0, 1.0)))
    phi = {"j": J, "levels": [_fold_to_resolution(g_hat, r) for r in range(J)]}
    return phi, psi


def filter_bank_2d(Mp, Np, J, L=8):
    """Returns ``{'phi': ..., 'psi': ...}`` (fresh containers, shared read-only arrays)."""
    phi, psi = _filter_bank_cached(int(Mp), int(Np), int(J), int(L))
    return {"phi": {"j": phi["j"], "levels": list(phi["levels"])},
            "psi": [{"j": p["j"], "theta": p["theta"], "levels": list(p["levels"])} for p in psi]}
