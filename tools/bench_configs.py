"""Side measurements for the other BASELINE.json configs (not the headline bench): torch_b200 vs the
reference's own torch GPU backend on the same device, reduced batches where noted.  Prints one JSON
object per line."""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import import_reference  # noqa: E402

assert import_reference(), "reference not installed under baseline/_ref"
import kymatio_b200.kymatio_plugin as plugin  # noqa: E402
plugin.install()
from kymatio.torch import Scattering1D, Scattering2D, HarmonicScattering3D  # noqa: E402


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


def report(name, unit, B, ms_ref, ms_b200, note=""):
    print(json.dumps({"config": name, "batch": B, "unit": unit,
                      "reference_torch_gpu": {"ms": ms_ref, "per_s": B / ms_ref * 1e3},
                      "torch_b200": {"ms": ms_b200, "per_s": B / ms_b200 * 1e3},
                      "speedup": ms_ref / ms_b200, "note": note}), flush=True)


which = sys.argv[1:] or ["c1", "c2", "c3", "c4", "c5"]
with torch.no_grad():
    if "c1" in which:
        B = 128
        x = torch.randn(B, 32, 32, device="cuda")
        Sr, Sb = Scattering2D(2, (32, 32), backend="torch").cuda(), Scattering2D(2, (32, 32), backend="torch_b200").cuda()
        report("C1 2D J=2 L=8 32x32", "images/s", B, timeit(lambda: Sr(x)), timeit(lambda: Sb(x), n=20))
    if "c2" in which:
        B = 64
        x = torch.randn(B, 256, 256, device="cuda")
        Sr, Sb = Scattering2D(3, (256, 256), backend="torch").cuda(), Scattering2D(3, (256, 256), backend="torch_b200").cuda()
        report("C2 2D J=3 L=8 256x256 (headline shape)", "images/s", B, timeit(lambda: Sr(x), n=3), timeit(lambda: Sb(x), n=20),
               "reduced batch 64")
    if "c3" in which:
        B = 32
        x = torch.randn(B, 2 ** 16, device="cuda")
        Sr = Scattering1D(8, 2 ** 16, Q=(8, 1), backend="torch").cuda()
        Sb = Scattering1D(8, 2 ** 16, Q=(8, 1), backend="torch_b200").cuda()
        report("C3 1D J=8 Q=(8,1) N=2^16", "signals/s", B, timeit(lambda: Sr(x), n=3), timeit(lambda: Sb(x), n=3),
               "fused 1-D path; reduced batch 32 (see tools/bench1d.py for batch 256)")
    if "c4" in which:
        B = 2
        x = torch.randn(B, 128, 128, 128, device="cuda")
        Sr = HarmonicScattering3D(2, (128, 128, 128), L=2, backend="torch").cuda()
        Sb = HarmonicScattering3D(2, (128, 128, 128), L=2, backend="torch_b200").cuda()
        report("C4 3D J=2 L=2 128^3", "volumes/s", B, timeit(lambda: Sr(x), n=2, warm=1), timeit(lambda: Sb(x), n=2, warm=1),
               "fused 3-D path; reduced batch 2 (see tools/bench3d.py for batch 16)")
if "c5" in which:
    B = 64
    Sr, Sb = Scattering2D(4, (224, 224), backend="torch").cuda(), Scattering2D(4, (224, 224), backend="torch_b200").cuda()

    def fb(S):
        x = torch.randn(B, 224, 224, device="cuda", requires_grad=True)
        S(x).sum().backward()
    report("C5 2D J=4 L=8 224x224 forward+backward", "images/s", B, timeit(lambda: fb(Sr), n=2, warm=1),
           timeit(lambda: fb(Sb), n=2, warm=1), "backward: fused second-order tiles + recomputed first-order graph; reduced batch 64")
    with torch.no_grad():
        x = torch.randn(B, 224, 224, device="cuda")
        report("C5 2D J=4 L=8 224x224 forward only", "images/s", B, timeit(lambda: Sr(x), n=2, warm=1), timeit(lambda: Sb(x), n=10))
