"""Host restatement of the index arithmetic of the banded low-pass adjoint (csrc/bwd2d.cuh: k2d_bwd_col, csrc/tile2d.cuh:
tile_bwd_body): the spatial low-pass matrix is G[y][yo] = a[(kl (yo+1) - y) mod n] inside the tap radius R (plan2d.cuh:
analyse_lowpass), and the kernels enumerate, per input sample y, the outputs it contributes to from j0 = ceil((y-R)/kl) ..
j1 = floor((y+R)/kl) (mod n/kl, minus the unpadded border) instead of scanning the dense row."""
import numpy as np
import pytest


def _window(y, R, kl, n, o):
    """The enumeration of k2d_bwd_col (C integer division semantics restated with explicit floor / ceil)."""
    mper = n // kl
    j0, j1 = y - R, y + R
    j0 = (j0 + kl - 1) // kl if j0 >= 0 else -((-j0) // kl)
    j1 = j1 // kl if j1 >= 0 else -((-j1 + kl - 1) // kl)
    out = []
    for jj in range(j0, j1 + 1):
        yo = jj % mper - 1
        if 0 <= yo < o:
            out.append(yo)
    return out, j1 - j0 + 1


@pytest.mark.parametrize("n,kl,R", [(256, 16, 34), (272, 8, 17), (128, 8, 17), (64, 4, 9), (32, 2, 5), (136, 4, 9),
                                    (40, 4, 3), (20, 2, 2), (240, 8, 20), (256, 16, 7), (272, 8, 1)])
def test_banded_window_equals_dense_row_support(n, kl, R):
    rng = np.random.RandomState(n + kl + R)
    o = n // kl - 2
    a = np.zeros(n)
    for t in range(-R, R + 1):
        a[t % n] = rng.uniform(0.1, 1.0)            # every tap inside the radius is nonzero
    G = np.zeros((n, o))
    for y in range(n):
        for yo in range(o):
            t = (kl * (yo + 1) - y) % n
            if min(t, n - t) <= R:
                G[y, yo] = a[t]
    assert 2 * R + 1 < n
    for y in range(n):
        win, count = _window(y, R, kl, n, o)
        assert count <= (2 * R) // kl + 2          # the register budget the kernel checks ((2R)/kl + 2 <= 8)
        assert len(set(win)) == len(win)
        assert sorted(win) == sorted(np.nonzero(G[y])[0].tolist()), (y, win)
    # and the adjoint they implement: gA = G @ T for a random T, accumulated through the windows
    T = rng.randn(o, 5)
    gA = np.zeros((n, 5))
    for y in range(n):
        for yo in _window(y, R, kl, n, o)[0]:
            gA[y] += G[y, yo] * T[yo]
    assert np.allclose(gA, G @ T)
