"""Forward+backward (input gradient) timing of one 2-D config with the per-kernel split.
usage: python tools/bwd_bench.py B J N      (SCAT_B200_ORDER1_FUSED=0 selects the per-primitive first-order graph)"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kymatio_b200 import Scattering2D, _lib  # noqa: E402

B, J, N = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
S = Scattering2D(J, (N, N)).cuda()
x = torch.randn(B, N, N, device="cuda")


def step():
    xi = x.detach().requires_grad_(True)
    S(xi).sum().backward()
    return xi.grad


for _ in range(2):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    g = step()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
_lib.timing_enable(True)
step()
rows = _lib.timing_report()
_lib.timing_enable(False)
rows.sort(key=lambda r: -r["ms"])
print(json.dumps({"order1_fused": os.environ.get("SCAT_B200_ORDER1_FUSED", "1"), "B": B, "J": J, "N": N, "ms": ms,
                  "img_per_s": B / ms * 1e3, "lib_ms": sum(r["ms"] for r in rows),
                  "top": {r["label"]: round(r["ms"], 3) for r in rows[:16]}}))
