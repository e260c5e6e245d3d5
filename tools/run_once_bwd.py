"""A few forward+backward steps of the C5 shape (for ncu captures of the backward kernels)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kymatio_b200 import Scattering2D  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
n = int(sys.argv[2]) if len(sys.argv) > 2 else 2
S = Scattering2D(4, (224, 224)).cuda()
x = torch.randn(B, 224, 224, device="cuda")
for _ in range(n):
    xi = x.detach().requires_grad_(True)
    S(xi).sum().backward()
torch.cuda.synchronize()
print(float(xi.grad.abs().mean()))
