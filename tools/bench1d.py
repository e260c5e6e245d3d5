"""C3 (Scattering1D J=8 Q=(8,1) N=2^16) through the unmodified kymatio frontend: fused torch_b200 vs the eager
torch_b200 primitives vs the reference's torch GPU backend; per-kernel timing of the fused path."""
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
from kymatio.torch import Scattering1D  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
ref_too = "--ref" in sys.argv


def timeit(fn, n=5, warm=2):
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
    x = torch.randn(B, 2 ** 16, device="cuda")
    Sb = Scattering1D(8, 2 ** 16, Q=(8, 1), backend="torch_b200").cuda()
    ms = timeit(lambda: Sb(x), n=5)
    res = {"config": "C3 1D J=8 Q=(8,1) N=2^16", "batch": B, "fused_ms": ms, "fused_per_s": B / ms * 1e3,
           "pass_model_GBps": B / ms * 1e3 * 173451264 / 1e9}
    _lib.timing_enable(True)
    Sb(x)
    rows = _lib.timing_report()
    _lib.timing_enable(False)
    tot = sum(r["ms"] for r in rows)
    agg = {}
    for r in rows:
        k = r["label"].split(":")[0]
        a = agg.setdefault(k, [0.0, 0.0, 0])
        a[0] += r["ms"]; a[1] += r["bytes"]; a[2] += r["count"]
    res["kernel_ms_sum"] = tot
    res["kernels"] = {k: {"ms": v[0], "GBps": v[1] / v[0] / 1e6 if v[0] else 0, "launches": v[2]} for k, v in agg.items()}
    res["top"] = sorted([(r["label"], round(r["ms"], 3), round(r["bytes"] / max(r["ms"], 1e-9) / 1e6)) for r in rows],
                        key=lambda t: -t[1])[:12]
    if ref_too:
        xs = x[:32]
        Sr = Scattering1D(8, 2 ** 16, Q=(8, 1), backend="torch").cuda()
        msr = timeit(lambda: Sr(xs), n=3)
        res["reference_torch_gpu_per_s"] = 32 / msr * 1e3
        plugin.install(fused=False)
        Se = Scattering1D(8, 2 ** 16, Q=(8, 1), backend="torch_b200").cuda()
        mse = timeit(lambda: Se(xs), n=3)
        res["eager_b200_per_s"] = 32 / mse * 1e3
        plugin.install(fused=True)
        a, b = Sb(xs), Sr(xs)
        res["max_rel_vs_ref_torch"] = float((a - b).abs().max() / b.abs().max())
    print(json.dumps(res), flush=True)
