"""C4 (HarmonicScattering3D J=2 L=2 128^3) through the unmodified kymatio frontend: fused torch_b200 with per-kernel
timing; optionally the reference torch GPU backend."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import import_reference  # noqa: E402

assert import_reference()
import kymatio_b200.kymatio_plugin as plugin  # noqa: E402
from kymatio_b200 import _lib  # noqa: E402
plugin.install()
from kymatio.torch import HarmonicScattering3D  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16


def timeit(fn, n=3, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


with torch.no_grad():
    x = torch.randn(B, 128, 128, 128, device="cuda")
    Sb = HarmonicScattering3D(2, (128, 128, 128), L=2, backend="torch_b200").cuda()
    ms = timeit(lambda: Sb(x))
    res = {"config": "C4 3D J=2 L=2 128^3", "batch": B, "fused_ms": ms, "fused_per_s": B / ms * 1e3}
    _lib.timing_enable(True)
    Sb(x)
    rows = _lib.timing_report()
    _lib.timing_enable(False)
    agg = {}
    for r in rows:
        a = agg.setdefault(r["label"], [0.0, 0.0, 0])
        a[0] += r["ms"]; a[1] += r["bytes"]; a[2] += r["count"]
    res["kernel_ms_sum"] = sum(r["ms"] for r in rows)
    res["kernels"] = {k: {"ms": round(v[0], 3), "GBps": round(v[1] / v[0] / 1e6) if v[0] else 0, "launches": v[2]}
                      for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])}
    if "--ref" in sys.argv:
        xs = x[:2]
        Sr = HarmonicScattering3D(2, (128, 128, 128), L=2, backend="torch").cuda()
        msr = timeit(lambda: Sr(xs), n=2)
        res["reference_torch_gpu_per_s"] = 2 / msr * 1e3
        a, b = Sb(xs), Sr(xs)
        res["max_rel_vs_ref_torch"] = float(((a - b).abs() / b.abs().clamp_min(1e-30)).max())
    print(json.dumps(res), flush=True)
