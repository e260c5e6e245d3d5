"""Run a few forwards of the headline config (for ncu captures)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kymatio_b200 import Scattering2D  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
n = int(sys.argv[2]) if len(sys.argv) > 2 else 2
S = Scattering2D(3, (256, 256)).cuda()
x = torch.randn(B, 256, 256, device="cuda")
for _ in range(n):
    y = S(x)
torch.cuda.synchronize()
print(y.shape, float(y.abs().mean()))
