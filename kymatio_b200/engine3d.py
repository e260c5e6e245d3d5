"""Host engine of the fused 3-D solid-harmonic scattering path (``scat3d_*`` in include/scat_b200.h).

Replaces the per-primitive loop of kymatio/scattering3d/core/scattering3d.py:24-73 by, per band (l, j):
    col_prod  (cdgmm3d with all 2l+1 filters + inverse transform along M)
    plane     (2-D inverse of every (N, O) plane, sum_m |.|^2, sqrt, compute_integrals; parents: 2-D forward)
    col_fwd   (parents: forward transform along M -> U1_hat)
The engine owns only torch tensors (tables, workspaces, the float64 integral accumulator); there is no CPU path.
"""
import ctypes
import os

import torch

from . import _lib


class Unsupported(Exception):
    pass


def _stream(dev):
    return ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def band_layout(L, J, max_order):
    """Index of every band in the reference's output axis n_j (core/scattering3d.py:62-73): first the J+1
    first-order scales, then the (j1, j2 > j1) pairs in loop order."""
    first = {j1: j1 for j1 in range(J + 1)}
    second, n = {}, J + 1
    if max_order > 1:
        for j1 in range(J + 1):
            for j2 in range(j1 + 1, J + 1):
                second[(j1, j2)] = n
                n += 1
    return first, second, n


class Engine3D:
    _tables = {}

    def __init__(self, M, N, O, device):
        self.lib = _lib.load()
        self.device = torch.device(device)
        self.shape = (int(M), int(N), int(O))
        if not self.lib.scat3d_supported(*self.shape):
            raise Unsupported("no fused 3-D kernel instance for this volume shape")
        with torch.cuda.device(self.device):
            self.tab = torch.empty(self.lib.scat3d_tables_bytes(*self.shape), dtype=torch.uint8, device=self.device)
            _lib.check(self.lib.scat3d_tables_init(self.tab.data_ptr(), *self.shape, _stream(self.device)))

    def rfft(self, x):
        """x: (B, M, N, O) float32 real volumes -> (B, M, N, O, 2) natural-order spectrum (core/scattering3d.py:24)."""
        M, N, O = self.shape
        out = torch.empty((x.shape[0], M, N, O, 2), dtype=torch.float32, device=self.device)
        if x.shape[0]:
            with torch.cuda.device(self.device):
                _lib.check(self.lib.scat3d_rfft(self.tab.data_ptr(), x.data_ptr(), out.data_ptr(), x.shape[0], M, N, O,
                                                _stream(self.device)))
        return out

    def forward(self, U0_hat, filters, rotation_covariant, L, J, max_order, powers):
        """U0_hat: (B, M, N, O, 2) float32 spectrum of the input volumes; filters[l]: (J+1, 2l+1, M, N, O, 2).
        Returns (B, n_j, L+1, P) float32 in the reference's layout."""
        lib, dev = self.lib, self.device
        M, N, O = self.shape
        B, P = U0_hat.shape[0], len(powers)
        first, second, n_j = band_layout(L, J, max_order)
        istride = n_j * (L + 1) * P
        vol = M * N * O * 8
        nm_max = max((f.shape[1] if rotation_covariant else 1) for f in filters[:L + 1])
        budget = int(os.environ.get("SCAT_B200_WS3D_MB", "16384")) << 20
        Bc = max(1, min(B, budget // (vol * (nm_max + 1))))
        with torch.cuda.device(dev):
            st = _stream(dev)
            acc = torch.zeros((B, istride), dtype=torch.float64, device=dev)
            pw = torch.tensor([float(q) for q in powers], dtype=torch.float32, device=dev)
            Y = torch.empty(Bc * nm_max * vol, dtype=torch.uint8, device=dev)
            U1 = torch.empty(Bc * vol, dtype=torch.uint8, device=dev)
            tab = self.tab.data_ptr()
            for b0 in range(0, B, Bc):
                nb = min(Bc, B - b0)
                u0 = U0_hat.data_ptr() + b0 * vol
                ap = acc.data_ptr() + b0 * istride * 8

                def band(parent_ptr, l, j, spec_ptr, slot):
                    f = filters[l][j]
                    nm = f.shape[0] if rotation_covariant else 1
                    _lib.check(lib.scat3d_col_prod(tab, parent_ptr, f.data_ptr(), Y.data_ptr(), nb, nm, M, N, O, st))
                    _lib.check(lib.scat3d_plane(tab, Y.data_ptr(), spec_ptr, ap, istride, (slot * (L + 1) + l) * P,
                                                pw.data_ptr(), P, nb, nm, M, N, O, st))

                for l in range(L + 1):
                    for j1 in range(J + 1):
                        kids = [j2 for j2 in range(j1 + 1, J + 1)] if max_order > 1 else []
                        band(u0, l, j1, U1.data_ptr() if kids else None, first[j1])
                        if kids:
                            _lib.check(lib.scat3d_col_fwd(tab, U1.data_ptr(), U1.data_ptr(), nb, M, N, O, st))
                            for j2 in kids:
                                band(U1.data_ptr(), l, j2, None, second[(j1, j2)])
        return acc.to(torch.get_default_dtype()).reshape(B, n_j, L + 1, P)
