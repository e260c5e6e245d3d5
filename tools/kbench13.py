"""Per-kernel timing of the C3 (1-D) and C4 (3-D) configs through the kymatio.torch frontends (backend torch_b200).
usage: python tools/kbench13.py c3|c4 [batch]"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import import_reference  # noqa: E402
from kymatio_b200 import _lib, kymatio_plugin  # noqa: E402

assert import_reference()
kymatio_plugin.install()
from kymatio.torch import HarmonicScattering3D, Scattering1D  # noqa: E402

which = sys.argv[1]
torch.manual_seed(0)
if which == "c3":
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 512
    S = Scattering1D(J=8, shape=2 ** 16, Q=(8, 1), backend="torch_b200").cuda()
    x = torch.randn(B, 2 ** 16, device="cuda")
else:
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 16
    S = HarmonicScattering3D(J=2, shape=(128, 128, 128), L=2, backend="torch_b200").cuda()
    x = torch.randn(B, 128, 128, 128, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
steps = 5
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
agg = {}
for r in rows:
    k = r["label"].split(":")[0]
    a = agg.setdefault(k, [0.0, 0, 0.0])
    a[0] += r["ms"] / 2; a[1] += r.get("count", 0) // 2 if isinstance(r.get("count", 0), int) else 0; a[2] += r.get("bytes", 0.0) / 2
print(json.dumps({"which": which, "B": B, "ms_median": med, "units_per_s": B / med * 1e3, "lib_ms": sum(v[0] for v in agg.values()),
                  "kernels": {k: {"ms": round(v[0], 3), "n": v[1], "GBps": round(v[2] / max(v[0], 1e-9) / 1e6, 1)} for k, v in
                              sorted(agg.items(), key=lambda kv: -kv[1][0])}}), flush=True)
