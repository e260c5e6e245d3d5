"""Per-phase cycle breakdown of the backward tile kernels (profiling build, `make -C kymatio_b200/csrc prof`).

    SCAT_B200_LIB=kymatio_b200/lib/libscat_b200_prof.so python tools/phase_prof_bwd.py [batch] [J] [size]
"""
import ctypes
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kymatio_b200 import Scattering2D, _lib  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
J = int(sys.argv[2]) if len(sys.argv) > 2 else 4
N = int(sys.argv[3]) if len(sys.argv) > 3 else 224
PH = ["stage", "load", "low_adj_vert", "ifft2", "low_adj_horz+mod_bwd", "fft2", "scatter", "-"]
lib = _lib.load()
S = Scattering2D(J, (N, N), L=8).cuda()
x = torch.randn(B, N, N, device="cuda")


def step():
    xi = x.detach().requires_grad_(True)
    S(xi).sum().backward()


NB = 32 * 8
buf = (ctypes.c_uint64 * NB)()
step()
n = lib.scat_phase_prof_read(buf, NB, 1)
assert n == NB, "not the profiling build (set SCAT_B200_LIB=.../libscat_b200_prof.so)"
step()
lib.scat_phase_prof_read(buf, NB, 1)
rows = {}
for kid in range(24, 32):
    v = [int(buf[kid * 8 + p]) for p in range(8)]
    tot = sum(v)
    if not tot:
        continue
    size = [">=128", ">=64", ">=32", "<32"][(kid - 24) // 2]
    name = f"bwd_tile{size}_K{'>2' if kid & 1 else '2'}"
    rows[name] = {"total_Mcycles": tot / 1e6, **{PH[p]: round(v[p] / tot, 4) for p in range(7)}}
    print(name, "total Mcyc %.1f" % (tot / 1e6), " ".join(f"{PH[p]}={v[p] / tot:.3f}" for p in range(7)))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(rows, open(f"gpurun_out/phase_prof_bwd_B{B}_J{J}_N{N}.json", "w"), indent=1)
