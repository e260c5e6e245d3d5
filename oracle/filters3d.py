"""numpy oracle for the 3-D solid-harmonic / Gaussian filter banks (TEST INFRASTRUCTURE - see oracle/__init__.py).

Closed-form restatement of kymatio/scattering3d/filter_bank.py:5-166 in float64 - the arithmetic the device kernels
(kymatio_b200/csrc/filters.cuh: kf_solid_harmonic3d, kf_gaussian3d) implement:

    grid      w_a(i) = (-ceil(n/2) + ((i + n//2) mod n)) * 2 pi / n      (np.mgrid[-n//2 : -n//2 + n] after ifftshift:
                                                                          filter_bank.py:80-84,135-139; odd n keeps the
                                                                          reference's off-by-one origin)
    angles    cos(polar) = -z/r, sin(polar) = |xy|/r, azimuth = atan2(y, x)   (scattering3d/utils.py get_3d_angles,
                                                                                (z, y, x) = grid axes 0, 1, 2)
    wavelet   c_l (2 pi)^{3/2} (-i)^l (r sigma)^l exp(-r^2 sigma^2 / 2) Y_l^m(polar, azimuth)    (:141-165)
    Y_l^m     sqrt((2l+1)/(4 pi) (l-m)!/(l+m)!) P_l^m(cos polar) e^{i m azimuth}, Condon-Shortley phase, by the
              standard upward recursion; Y_l^{-m} = (-1)^m conj(Y_l^m)     (scipy.special.sph_harm's convention)

Pinned against reference-generated banks (tests/golden/golden_filters_3d.npz) in tests/test_oracle_1d3d.py.
"""
import math

import numpy as np

__all__ = ["solid_harmonic_filter_bank", "gaussian_filter_bank"]


def _axis(n):
    i = np.arange(n)
    return (-((n + 1) // 2) + ((i + n // 2) % n)) * (2 * np.pi / n)


def _norm(l):
    if l % 2 == 0:
        dfact = 1.0
        for i in range(l + 1, 0, -2):
            dfact *= i
        c = 1.0 / (2 * math.pi * math.sqrt(l + 0.5) * dfact)
    else:
        c = 1.0 / (2 ** (0.5 * (l + 3)) * math.sqrt(math.pi * (2 * l + 1)) * math.factorial((l + 1) // 2))
    return c * (2 * math.pi) ** 1.5


def gaussian_filter_bank(M, N, O, J, sigma_0):
    z, y, x = np.meshgrid(_axis(M), _axis(N), _axis(O), indexing="ij")
    r2 = x * x + y * y + z * z
    return np.stack([np.exp(-0.5 * r2 * (sigma_0 * 2 ** j) ** 2) for j in range(J + 1)]).astype(np.complex128)


def solid_harmonic_filter_bank(M, N, O, J, L, sigma_0):
    z, y, x = np.meshgrid(_axis(M), _axis(N), _axis(O), indexing="ij")
    r2 = x * x + y * y + z * z
    r, rxy = np.sqrt(r2), np.sqrt(x * x + y * y)
    safe = np.where(r > 0, r, 1.0)
    ct, st = np.where(r > 0, -z / safe, 0.0), np.where(r > 0, rxy / safe, 1.0)
    az = np.arctan2(y, x)
    bank = []
    for l in range(L + 1):
        out = np.zeros((J + 1, 2 * l + 1, M, N, O), np.complex128)
        for j in range(J + 1):
            sigma = sigma_0 * 2 ** j
            gauss = np.exp(-0.5 * r2 * sigma * sigma)
            if l == 0:
                out[j, 0] = gauss
                continue
            radial = (r * sigma) ** l * gauss * _norm(l) * (-1j) ** l
            for m in range(l + 1):
                pmm = np.ones_like(r)
                for i in range(1, m + 1):
                    pmm = pmm * (-(2 * i - 1) * st)
                plm = pmm
                if l > m:
                    p0, p1 = pmm, ct * (2 * m + 1) * pmm
                    plm = p1
                    for ll in range(m + 2, l + 1):
                        plm = ((2 * ll - 1) * ct * p1 - (ll + m - 1) * p0) / (ll - m)
                        p0, p1 = p1, plm
                ratio = 1.0
                for i in range(l - m + 1, l + m + 1):
                    ratio /= i
                Y = math.sqrt((2 * l + 1) / (4 * math.pi) * ratio) * plm * np.exp(1j * m * az)
                out[j, l + m] = radial * Y
                if m > 0:
                    out[j, l - m] = radial * ((-1) ** m) * np.conj(Y)
        bank.append(out)
    return bank
