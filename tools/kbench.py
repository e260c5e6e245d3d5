"""Quick A/B bench of one 2-D config with per-kernel timing (CUDA events inside the library).
usage: python tools/kbench.py [label] [batch] [J] [size] [steps]   (env SCAT_B200_* select kernel variants)"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kymatio_b200 import Scattering2D, _lib  # noqa: E402

label = sys.argv[1] if len(sys.argv) > 1 else "run"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
J = int(sys.argv[3]) if len(sys.argv) > 3 else 3
N = int(sys.argv[4]) if len(sys.argv) > 4 else 256
steps = int(sys.argv[5]) if len(sys.argv) > 5 else 10
S = Scattering2D(J, (N, N), L=8).cuda()
torch.manual_seed(0)
x = torch.randn(B, N, N, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
with torch.no_grad():
    for _ in range(3):
        y = S(x)
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for e0, e1 in ev:
        flush.zero_()
        e0.record(); y = S(x); e1.record()
    torch.cuda.synchronize()
    ms = sorted(e0.elapsed_time(e1) for e0, e1 in ev)
    _lib.timing_enable(True)
    for _ in range(2):
        flush.zero_()
        S(x)
    rows = _lib.timing_report()
    _lib.timing_enable(False)
med = ms[len(ms) // 2]
out = {"label": label, "B": B, "J": J, "N": N, "ms_median": med, "ms_min": ms[0], "img_per_s": B / med * 1e3,
       "checksum": float(y.double().abs().sum()),
       "env": {k: v for k, v in os.environ.items() if k.startswith("SCAT_B200_") and k != "SCAT_B200_LIB"},
       "kernels": {r["label"]: round(r["ms"] / 2, 4) for r in sorted(rows, key=lambda r: -r["ms"])}}
print(json.dumps(out), flush=True)
