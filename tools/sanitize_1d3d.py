"""Small fused 1-D / 3-D forwards for compute-sanitizer (memcheck / racecheck)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import import_reference  # noqa: E402

assert import_reference()
import kymatio_b200.kymatio_plugin as plugin  # noqa: E402
plugin.install()
from kymatio.torch import Scattering1D, HarmonicScattering3D  # noqa: E402

with torch.no_grad():
    S = Scattering1D(J=5, shape=2048, Q=(4, 1), backend="torch_b200").cuda()
    y = S(torch.randn(2, 2048, device="cuda"))
    S3 = HarmonicScattering3D(J=1, shape=(16, 16, 16), L=1, backend="torch_b200").cuda()
    z = S3(torch.randn(2, 16, 16, 16, device="cuda"))
    torch.cuda.synchronize()
print("ok", tuple(y.shape), tuple(z.shape), float(y.abs().mean()), float(z.abs().mean()))
