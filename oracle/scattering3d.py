"""numpy oracle for the 3-D solid-harmonic scattering hot path (TEST INFRASTRUCTURE - see oracle/__init__.py).

Restates, in plain numpy (float64 / complex128):
  * rfft / ifft            kymatio/scattering3d/backend/numpy_backend.py:62-70  (fftn / ifftn over the last 3 axes)
  * cdgmm3d                kymatio/backend/numpy_backend.py:65-95               (complex filter branch)
  * modulus_rotation       kymatio/scattering3d/backend/numpy_backend.py:5-29   (sqrt(prev^2 + |x|^2))
  * compute_integrals      kymatio/scattering3d/backend/numpy_backend.py:31-54
  * stack                  kymatio/scattering3d/backend/numpy_backend.py:56-60
  * the cascade + order    kymatio/scattering3d/core/scattering3d.py:24-73

The solid-harmonic filter bank (kymatio/scattering3d/filter_bank.py:8-184) is constructor-time code outside the
hot path and is NOT restated: the caller passes `filters[l]` as complex arrays of shape (J+1, 2l+1, M, N, O).
Pinned against reference outputs (tests/golden/golden_3d_*.npz, ref_fixture_3d.npz) in tests/test_oracle_1d3d.py.
"""
import numpy as np
import scipy.fft

__all__ = ["scattering3d"]


def _integrals(U, powers):
    # numpy_backend.py:31-54 (the reference casts the result to float32; kept in float64 here)
    return np.stack([(U ** q).reshape(U.shape[0], -1).sum(axis=1) for q in powers], axis=-1)


def scattering3d(x, filters, L, J, integral_powers, max_order=2, rotation_covariant=True):
    """x: (B, M, N, O) real -> (B, n_j, L+1, P), n_j = (J+1) [+ (J+1)J/2 when max_order == 2]."""
    x = np.asarray(x, dtype=np.float64)
    ax = (-3, -2, -1)
    U0_hat = scipy.fft.fftn(x, axes=ax)

    def band(U_hat, l, j):                                                     # core:31-39 / 48-56
        fl = np.asarray(filters[l][j])
        if not rotation_covariant:
            return np.abs(scipy.fft.ifftn(U_hat * fl[0], axes=ax))
        acc = None
        for m in range(fl.shape[0]):
            u = np.abs(scipy.fft.ifftn(U_hat * fl[m], axes=ax)) ** 2
            acc = u if acc is None else acc + u                                # == nested sqrt(prev^2 + |x|^2)
        return np.sqrt(acc)

    order1, order2 = [], []
    for l in range(L + 1):
        o1, o2 = [], []
        for j1 in range(J + 1):
            U1 = band(U0_hat, l, j1)
            o1.append(_integrals(U1, integral_powers))
            if max_order > 1:
                U1_hat = scipy.fft.fftn(U1, axes=ax)
                for j2 in range(j1 + 1, J + 1):
                    o2.append(_integrals(band(U1_hat, l, j2), integral_powers))
        order1.append(o1)
        order2.append(o2)
    S = [a + b for a, b in zip(order1, order2)] if max_order == 2 else order1    # core:62-67
    S = [v for grp in zip(*S) for v in grp]                                     # (l, j) -> (j, l)   core:69-70
    S = np.stack(S, axis=1)                                                     # stack: numpy_backend.py:56-60
    return S.reshape((S.shape[0], S.shape[1] // (L + 1), L + 1) + S.shape[2:])
