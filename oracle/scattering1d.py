"""numpy oracle for the 1-D scattering hot path (TEST INFRASTRUCTURE - see oracle/__init__.py).

Restates, in plain numpy (float64), the arithmetic below the frontend:
  * reflect pad            kymatio/scattering1d/backend/numpy_backend.py:44-78
  * subsample_fourier      kymatio/scattering1d/backend/numpy_backend.py:8-41
  * rfft / ifft / irfft    kymatio/scattering1d/backend/numpy_backend.py:86-108
  * cdgmm / modulus        kymatio/backend/numpy_backend.py:45-95
  * unpad                  kymatio/scattering1d/backend/numpy_backend.py:80-84
  * the cascade            kymatio/scattering1d/core/scattering1d.py:40-107
  * unpad / sort / stack   kymatio/scattering1d/frontend/base_frontend.py:106-155

The filter bank (kymatio/scattering1d/filter_bank.py:51-404) is constructor-time code outside the hot path and is
NOT restated: the caller passes the filter dictionaries (phi, psi1, psi2 with 'j' and 'levels') together with
the padding/unpadding indices the frontend derives (base_frontend.py:83-98).  The FFTs live in scipy.fft (unpinned
third-party dependency, requirements.txt:1-5) and are called through the same entry points as the reference.
Pinned against reference outputs (tests/golden/golden_1d_*.npz, ref_fixture_1d.npz) in tests/test_oracle_1d3d.py.
"""
import numpy as np
import scipy.fft

__all__ = ["pad", "subsample_fourier", "scattering1d", "meta_order"]


def pad(x, pad_left, pad_right):
    # numpy_backend.py:44-78 (mode='reflect'); refuses pads >= the signal length like the reference
    if pad_left >= x.shape[-1] or pad_right >= x.shape[-1]:
        raise ValueError("Indefinite padding size (larger than tensor).")
    return np.pad(x, ((0, 0),) * (x.ndim - 1) + ((pad_left, pad_right),), mode="reflect")


def subsample_fourier(x, k):
    # numpy_backend.py:8-41: mean of the k aliases
    N = x.shape[-1]
    return x.reshape(x.shape[:-1] + (k, N // k)).mean(axis=-2)


def meta_order(psi1, psi2, max_order=2):
    """Channel keys in the frontend's output order: sorted by (order, n) (base_frontend.py:150)."""
    keys = [()] + [(n1,) for n1 in range(len(psi1))]
    if max_order == 2:
        keys += [(n1, n2) for n1 in range(len(psi1)) for n2 in range(len(psi2)) if psi2[n2]["j"] > psi1[n1]["j"]]
    return keys


def scattering1d(x, phi, psi1, psi2, log2_stride, pad_left, pad_right, ind_start, ind_end, max_order=2):
    """x: (B, N) real -> (B, K, (ind_end - ind_start)[log2_stride]); average='local' (T > 0)."""
    x = np.asarray(x, dtype=np.float64)
    lv = lambda f: np.asarray(f, dtype=np.float64).reshape(-1)                 # noqa: E731
    U0 = pad(x, pad_left, pad_right)
    U0_hat = scipy.fft.fft(U0, axis=-1)                                         # rfft: numpy_backend.py:86-90
    ls = int(log2_stride)
    i0, i1 = ind_start[ls], ind_end[ls]
    low = lambda U_hat, level, k: scipy.fft.ifft(subsample_fourier(U_hat * lv(phi["levels"][level]), 2 ** k), axis=-1).real  # noqa: E731
    out = {(): low(U0_hat, 0, ls)[..., i0:i1]}                                  # core:47-51
    for n1, p1 in enumerate(psi1):                                              # core:57-79
        j1 = p1["j"]
        k1 = min(j1, ls)
        U1 = np.abs(scipy.fft.ifft(subsample_fourier(U0_hat * lv(p1["levels"][0]), 2 ** k1), axis=-1))
        U1_hat = scipy.fft.fft(U1, axis=-1)
        out[(n1,)] = low(U1_hat, k1, max(ls - k1, 0))[..., i0:i1]
        if max_order == 2:
            for n2, p2 in enumerate(psi2):                                      # core:81-107
                j2 = p2["j"]
                if j2 > j1:
                    s2 = min(j2, ls)
                    k2 = max(s2 - k1, 0)
                    U2 = np.abs(scipy.fft.ifft(subsample_fourier(U1_hat * lv(p2["levels"][k1]), 2 ** k2), axis=-1))
                    U2_hat = scipy.fft.fft(U2, axis=-1)
                    out[(n1, n2)] = low(U2_hat, k1 + k2, max(ls - s2, 0))[..., i0:i1]
    return np.stack([out[k] for k in meta_order(psi1, psi2, max_order)], axis=-2)
