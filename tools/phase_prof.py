"""Per-phase cycle breakdown of the 2-D tile kernels (profiling build, `make -C kymatio_b200/csrc prof`).

    SCAT_B200_LIB=kymatio_b200/lib/libscat_b200_prof.so python tools/phase_prof.py [batch] [J] [size]

Prints, per tile-kernel kind, the share of CTA time spent in: support staging, product+periodise load, row inverse
FFT, column inverse FFT + modulus, low-pass 4a, low-pass 4b, forward FFT + store / tail."""
import ctypes
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kymatio_b200 import Scattering2D, _lib  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
J = int(sys.argv[2]) if len(sys.argv) > 2 else 3
N = int(sys.argv[3]) if len(sys.argv) > 3 else 256
PH = ["supp", "load", "ifft_rows", "ifft_cols+mod", "low4a", "low4b", "fwd_fft/tail", "-"]
lib = _lib.load()
S = Scattering2D(J, (N, N), L=8).cuda()
x = torch.randn(B, N, N, device="cuda")
buf = (ctypes.c_uint64 * 192)()
with torch.no_grad():
    for _ in range(2):
        S(x)
    n = lib.scat_phase_prof_read(buf, 192, 1)
    assert n == 192, "not the profiling build (set SCAT_B200_LIB=.../libscat_b200_prof.so)"
    S(x)
    lib.scat_phase_prof_read(buf, 192, 1)
rows = {}
for kid in range(24):
    v = [int(buf[kid * 8 + p]) for p in range(8)]
    tot = sum(v)
    if not tot:
        continue
    size = {0: 136, 1: 68, 2: "other"}[kid // 8]
    name = f"tile{size}_K{4 if kid & 4 else 2}{'_spec' if kid & 2 else ''}{'_o2' if kid & 1 else '_o1'}"
    rows[name] = {"total_Mcycles": tot / 1e6, **{PH[p]: round(v[p] / tot, 4) for p in range(7)}}
    print(name, "total Mcyc %.1f" % (tot / 1e6), " ".join(f"{PH[p]}={v[p] / tot:.3f}" for p in range(7)))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(rows, open(f"gpurun_out/phase_prof_B{B}_J{J}_N{N}.json", "w"), indent=1)
