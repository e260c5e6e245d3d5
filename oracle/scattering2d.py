"""numpy oracle for the 2D scattering hot path (TEST INFRASTRUCTURE - see
oracle/__init__.py).

Restates, in plain numpy:
  * padded sizes          kymatio/scattering2d/utils.py:19-22
  * pad amounts           kymatio/scattering2d/frontend/base_frontend.py:26-28
  * Morlet filter bank    kymatio/scattering2d/filter_bank.py:5-170
  * reflect pad           kymatio/scattering2d/backend/numpy_backend.py:20-25
  * subsample_fourier     kymatio/scattering2d/backend/numpy_backend.py:49-76
  * rfft / ifft / irfft   kymatio/scattering2d/backend/numpy_backend.py:87-99
  * cdgmm / modulus       kymatio/backend/numpy_backend.py:45-95
  * unpad                 kymatio/scattering2d/backend/numpy_backend.py:32-46
  * the cascade + order   kymatio/scattering2d/core/scattering2d.py:14-86

The FFTs themselves live in scipy.fft (unpinned third-party dependency of the
reference, requirements.txt:1-5); the oracle calls the same scipy entry points
the reference calls.  Pinned against the reference's golden fixture and
against reference outputs generated in the build container
(tests/golden/make_golden.py); see tests/test_oracle_2d.py.
"""
import numpy as np
import scipy.fft

__all__ = ["padded_size", "pad_amounts", "filter_bank", "scattering2d", "n_channels"]


def padded_size(M, N, J):
    # utils.py:21-22
    return ((M + 2 ** J) // 2 ** J + 1) * 2 ** J, ((N + 2 ** J) // 2 ** J + 1) * 2 ** J


def pad_amounts(M, N, J):
    # base_frontend.py:26-28  -> [top, bottom, left, right]
    Mp, Np = padded_size(M, N, J)
    return [(Mp - M) // 2, (Mp - M + 1) // 2, (Np - N) // 2, (Np - N + 1) // 2]


def n_channels(J, L, max_order=2):
    # doc/source/userguide.rst:118
    K = 1 + L * J
    if max_order == 2:
        K += L * L * J * (J - 1) // 2
    return K


# ----------------------------------------------------------------------------
# filter bank (filter_bank.py)
# ----------------------------------------------------------------------------
def _gabor(M, N, sigma, theta, xi, slant):
    # filter_bank.py:132-175: 5x5 periodised Gaussian-windowed plane wave,
    # accumulated in complex64, normalised by 2*3.1415*sigma^2/slant.
    gab = np.zeros((M, N), np.complex64)
    R = np.array([[np.cos(theta), -np.sin(theta)], [np.sin(theta), np.cos(theta)]], np.float32)
    R_inv = np.array([[np.cos(theta), np.sin(theta)], [-np.sin(theta), np.cos(theta)]], np.float32)
    D = np.array([[1, 0], [0, slant * slant]])
    curv = np.dot(R, np.dot(D, R_inv)) / (2 * sigma * sigma)
    for ex in (-2, -1, 0, 1, 2):
        for ey in (-2, -1, 0, 1, 2):
            xx = (np.arange(M) + ex * M)[:, None]
            yy = (np.arange(N) + ey * N)[None, :]
            arg = -(curv[0, 0] * xx * xx + (curv[0, 1] + curv[1, 0]) * xx * yy
                    + curv[1, 1] * yy * yy) \
                + 1.j * (xx * xi * np.cos(theta) + yy * xi * np.sin(theta))
            gab += np.exp(arg)
    gab /= (2 * 3.1415 * sigma * sigma / slant)
    return gab


def _morlet(M, N, sigma, theta, xi, slant):
    # filter_bank.py:94-129
    wv = _gabor(M, N, sigma, theta, xi, slant)
    wv_mod = _gabor(M, N, sigma, theta, 0, slant)
    K = np.sum(wv) / np.sum(wv_mod)
    return wv - K * wv_mod


def _periodize_filter(x, res):
    # filter_bank.py:56-91: zero the high-frequency band, then fold 2^res x
    # 2^res aliases onto the low-resolution grid.
    M, N = x.shape
    mask = np.ones(x.shape, np.float32)
    len_x = int(M * (1 - 2 ** (-res)))
    start_x = int(M * 2 ** (-res - 1))
    len_y = int(N * (1 - 2 ** (-res)))
    start_y = int(N * 2 ** (-res - 1))
    mask[start_x:start_x + len_x, :] = 0
    mask[:, start_y:start_y + len_y] = 0
    x = np.multiply(x, mask)
    k = 2 ** res
    m, n = M // k, N // k
    crop = np.zeros((m, n), x.dtype)
    for i in range(k):
        for j in range(k):
            crop += x[i * m:(i + 1) * m, j * n:(j + 1) * n]
    return crop


def filter_bank(Mp, Np, J, L=8):
    # filter_bank.py:5-53
    psi = []
    for j in range(J):
        for theta in range(L):
            sig = _morlet(Mp, Np, 0.8 * 2 ** j, (int(L - L / 2 - 1) - theta) * np.pi / L,
                          3.0 / 4.0 * np.pi / 2 ** j, 4.0 / L)
            sig_hat = np.real(scipy.fft.fft2(sig))
            levels = [_periodize_filter(sig_hat, res) for res in range(min(j + 1, max(J - 1, 1)))]
            psi.append({"j": j, "theta": theta, "levels": levels})
    phi_hat = np.real(scipy.fft.fft2(_gabor(Mp, Np, 0.8 * 2 ** (J - 1), 0, 0, 1.0)))
    phi = {"j": J, "levels": [_periodize_filter(phi_hat, res) for res in range(J)]}
    return {"phi": phi, "psi": psi}


# ----------------------------------------------------------------------------
# backend primitives (numpy_backend.py)
# ----------------------------------------------------------------------------
def _subsample_fourier(x, k):
    # scattering2d/backend/numpy_backend.py:72-76
    y = x.reshape(-1, k, x.shape[1] // k, k, x.shape[2] // k)
    return y.mean(axis=(1, 3))


def _low(U_hat, phi_level, k):
    # core/scattering2d.py:19-23 (and :43-47, :71-75)
    S = _subsample_fourier(U_hat * phi_level, k)
    return scipy.fft.ifft2(S).real[..., 1:-1, 1:-1]


def scattering2d(x, J, L=8, max_order=2, pre_pad=False, filters=None, return_paths=False):
    """Scattering coefficients of ``x`` (..., M, N) -> (..., K, M/2^J, N/2^J).

    Computes in the precision of ``x`` for complex data (float32 filters, as
    the reference registers them) - pass float64 ``x`` for the fp64 oracle.
    """
    x = np.asarray(x)
    batch_shape = x.shape[:-2]
    x = x.reshape((-1,) + x.shape[-2:])
    if pre_pad:
        Mp, Np = x.shape[-2:]
        U_r = x
    else:
        M, N = x.shape[-2:]
        Mp, Np = padded_size(M, N, J)
        t, b, l, r = pad_amounts(M, N, J)
        U_r = np.pad(x, ((0, 0), (t, b), (l, r)), mode="reflect")
    if filters is None:
        filters = filter_bank(Mp, Np, J, L)
    phi, psi = filters["phi"], filters["psi"]

    U0 = scipy.fft.fft2(U_r)
    S0, S1, S2 = [], [], []
    S0.append(((), _low(U0, phi["levels"][0], 2 ** J)))
    for n1, p1 in enumerate(psi):
        j1 = p1["j"]
        U1 = U0 * p1["levels"][0]
        if j1 > 0:
            U1 = _subsample_fourier(U1, 2 ** j1)
        U1 = scipy.fft.fft2(np.abs(scipy.fft.ifft2(U1)))
        S1.append(((n1,), _low(U1, phi["levels"][j1], 2 ** (J - j1))))
        if max_order < 2:
            continue
        for n2, p2 in enumerate(psi):
            j2 = p2["j"]
            if j2 <= j1:
                continue
            U2 = _subsample_fourier(U1 * p2["levels"][j1], 2 ** (j2 - j1))
            U2 = scipy.fft.fft2(np.abs(scipy.fft.ifft2(U2)))
            S2.append(((n1, n2), _low(U2, phi["levels"][j2], 2 ** (J - j2))))
    paths = S0 + S1 + S2
    out = np.stack([c for _, c in paths], axis=-3)
    out = out.reshape(batch_shape + out.shape[-3:])
    if return_paths:
        return out, [n for n, _ in paths]
    return out
