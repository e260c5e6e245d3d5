"""Parity (vs the reference torch backend on the same GPU) and throughput at image sizes far from the compiled ones."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import import_reference  # noqa: E402
import kymatio_b200.kymatio_plugin as plugin  # noqa: E402

assert import_reference()
plugin.install()
from kymatio.torch import Scattering2D  # noqa: E402

for (J, shape, B) in [(4, (512, 512), 4), (5, (1024, 1024), 2), (3, (300, 420), 4), (2, (96, 96), 16), (4, (384, 256), 4)]:
    torch.manual_seed(0)
    x = torch.randn(B, *shape, device="cuda")
    t0 = time.perf_counter()
    Sb = Scattering2D(J=J, shape=shape, backend="torch_b200").cuda()
    t1 = time.perf_counter()
    Sr = Scattering2D(J=J, shape=shape, backend="torch").cuda()
    t2 = time.perf_counter()
    with torch.no_grad():
        yb, yr = Sb(x), Sr(x)
        torch.cuda.synchronize()
        rel = float((yb - yr).abs().max() / yr.abs().max())
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        ev[0].record()
        for _ in range(3):
            Sb(x)
        ev[1].record()
        ev[2].record()
        for _ in range(3):
            Sr(x)
        ev[3].record()
        torch.cuda.synchronize()
    print(f"J={J} {shape} B={B}: max rel {rel:.2e}  torch_b200 {ev[0].elapsed_time(ev[1]) / 3:.2f} ms  reference torch {ev[2].elapsed_time(ev[3]) / 3:.2f} ms"
          f"  constructors {1e3 * (t1 - t0):.0f} / {1e3 * (t2 - t1):.0f} ms", flush=True)
    assert rel <= 1e-4
