"""Stage-level check of the fused 1-D kernels against torch.fft on the GPU (run through gpurun)."""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kymatio_b200 import _lib  # noqa: E402
from kymatio_b200.engine1d import _FINSEG, _Tables, _split, _stream, circular_support  # noqa: E402

dev = torch.device("cuda:0")
lib = _lib.load()
tabs = _Tables(dev)


def rel(a, b):
    return float((a - b).abs().max() / b.abs().max())


def run(N, k, NI, B, M, full_support=False):
    Npar = N * k
    st = _stream(dev)
    torch.manual_seed(N + k)
    parent = torch.randn(B, Npar, 2, device=dev)
    filt = torch.zeros(NI, Npar, device=dev)
    for i in range(NI):
        if full_support:
            filt[i] = torch.randn(Npar, device=dev)
        else:
            lo = (37 + 1000 * i) % Npar
            ln = max(3, Npar // (3 + i))
            idx = (torch.arange(ln, device=dev) + lo) % Npar
            filt[i, idx] = torch.rand(ln, device=dev) + 0.1
    supp = [circular_support(filt[i].cpu().numpy(), 1e-7) for i in range(NI)]
    filt_dev = torch.tensor([filt[i].data_ptr() for i in range(NI)], dtype=torch.int64, device=dev)
    supp_dev = torch.tensor(supp, dtype=torch.int32, device=dev).reshape(-1, 2).contiguous()
    G = B * NI
    Y = torch.empty(G, N, 2, device=dev)
    tab = tabs.for_length(N)
    _lib.check(lib.scat1d_col_prod(tab.data_ptr(), parent.data_ptr(), Npar, 0, filt_dev.data_ptr(), supp_dev.data_ptr(),
                                   Y.data_ptr(), G, NI, Npar, N, 0.0, st))
    # reference
    pc = torch.view_as_complex(parent)
    X = (pc[:, None, :] * filt[None]).reshape(B, NI, k, N).mean(2)
    u = torch.fft.ifft(X)
    U = u.abs()
    U1 = torch.fft.fft(U)
    Yl = Y.clone()
    _lib.check(lib.scat1d_row_mod(tab.data_ptr(), Y.data_ptr(), G, N, None, 0, 0.0, st))
    out = torch.empty(G, N, 2, device=dev)
    _lib.check(lib.scat1d_col_fwd(tab.data_ptr(), Y.data_ptr(), out.data_ptr(), G, N, 0.0, st))
    torch.cuda.synchronize()
    got = torch.view_as_complex(out).reshape(B, NI, N)
    e1 = rel(torch.view_as_real(got), torch.view_as_real(U1))
    # leaf + finish
    Na, Nb = _split(N)
    nparts = (Na + 15) // 16
    Fc = min(N // 2 + 1, 48)
    phi = torch.zeros(N, device=dev)
    w = torch.exp(-0.5 * (torch.arange(Fc, device=dev, dtype=torch.float32) / (Fc / 5.0)) ** 2)
    phi[:Fc] = w
    phi[N - Fc + 1:] = w[1:].flip(0)
    part = torch.empty(G, nparts, Fc, 2, device=dev)
    _lib.check(lib.scat1d_row_mod(tab.data_ptr(), Yl.data_ptr(), G, N, part.data_ptr(), Fc, 0.0, st))
    K = NI + 2
    chan = torch.arange(NI, dtype=torch.int32, device=dev) + 1
    S = torch.zeros(B, K, M, device=dev)
    ft = tabs.for_lowpass(M)
    rows = [(0, nparts * Fc, Fc, phi.data_ptr(), chan.data_ptr(), 2, nparts, N, Fc, NI, 0)]
    segs = torch.from_numpy(np.array(rows, dtype=_FINSEG).view(np.uint8).copy()).to(dev)
    _lib.check(lib.scat1d_finish(ft.data_ptr(), None, out.data_ptr(), part.data_ptr(), segs.data_ptr(), 1, G, M,
                                 S.data_ptr(), K * M, 0, M, 0.0, st))
    S2 = torch.zeros(B, K, M, device=dev)
    rows2 = [(0, N, 0, phi.data_ptr(), chan.data_ptr(), 1, 1, N, Fc, NI, 0)]
    segs2 = torch.from_numpy(np.array(rows2, dtype=_FINSEG).view(np.uint8).copy()).to(dev)
    _lib.check(lib.scat1d_finish(ft.data_ptr(), None, out.data_ptr(), part.data_ptr(), segs2.data_ptr(), 1, G, M,
                                 S2.data_ptr(), K * M, 0, M, 0.0, st))
    torch.cuda.synchronize()
    low = (U1 * phi).reshape(B, NI, N // M, M).mean(2)
    ref = torch.fft.ifft(low).real
    e2 = rel(S[:, 1:1 + NI], ref)
    e3 = rel(S2[:, 1:1 + NI], ref)
    e4 = 0.0
    if N <= lib.scat1d_tile_max():
        spec = torch.empty(G, N, 2, device=dev)
        tpart = torch.empty(G, Fc, 2, device=dev)
        _lib.check(lib.scat1d_tile(tab.data_ptr(), parent.data_ptr(), Npar, 0, filt_dev.data_ptr(), supp_dev.data_ptr(),
                                   spec.data_ptr(), tpart.data_ptr(), Fc, G, NI, Npar, N, 0.0, st))
        torch.cuda.synchronize()
        tg = torch.view_as_complex(spec).reshape(B, NI, N)
        e4 = max(rel(torch.view_as_real(tg), torch.view_as_real(U1)),
                 rel(tpart.reshape(B, NI, Fc, 2), torch.view_as_real(U1[..., :Fc].contiguous())))
    print(f"N={N:7d} k={k:3d} NI={NI} Na={Na} Nb={Nb} M={M}: spectrum {e1:.2e}  leaf-finish {e2:.2e}  spec-finish {e3:.2e}"
          f"  tile {e4:.2e}", "OK" if max(e1, e2, e3, e4) < 2e-5 else "FAIL", flush=True)
    return max(e1, e2, e3, e4) < 2e-5


def run_rfft(N, B=3):
    torch.manual_seed(N)
    x = torch.randn(B, N, device=dev)
    out = torch.empty(B, N, 2, device=dev)
    _lib.check(lib.scat1d_rfft(tabs.for_length(N).data_ptr(), x.data_ptr(), out.data_ptr(), out.data_ptr(), B, N, _stream(dev)))
    torch.cuda.synchronize()
    ref = torch.view_as_real(torch.fft.fft(x))
    e = rel(out, ref)
    print(f"rfft N={N:7d}: {e:.2e}", "OK" if e < 2e-6 else "FAIL", flush=True)
    return e < 2e-6


ok = True
for N in (16, 64, 512, 2048, 8192, 65536, 131072, 262144):
    ok &= run_rfft(N)
for (N, k, NI, B, M, full) in [(16, 1, 1, 2, 8, True), (32, 2, 2, 3, 16, True), (64, 1, 3, 2, 32, False),
                               (128, 4, 2, 2, 32, False), (256, 2, 3, 3, 64, False), (512, 1, 2, 2, 128, True),
                               (1024, 8, 3, 2, 256, False), (2048, 2, 2, 2, 512, False), (4096, 1, 2, 2, 512, False),
                               (8192, 4, 3, 2, 512, False), (16384, 2, 2, 2, 512, False), (32768, 4, 2, 2, 512, False),
                               (65536, 2, 3, 2, 512, False), (131072, 1, 2, 2, 512, False), (262144, 1, 1, 2, 1024, False)]:
    ok &= run(N, k, NI, B, M, full)
print("ALL OK" if ok else "SOME FAILED")
sys.exit(0 if ok else 1)
