"""Small forward/backward runs for compute-sanitizer (memcheck / racecheck / initcheck)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kymatio_b200 import Scattering2D
for (J, shape, B) in [(2, (32, 32), 2), (3, (64, 64), 1), (2, (33, 47), 1), (3, (256, 256), 1)]:
    S = Scattering2D(J, shape).cuda()
    x = torch.randn(B, *shape, device="cuda", requires_grad=True)
    y = S(x)
    if shape[0] <= 64:
        y.sum().backward()
    torch.cuda.synchronize()
    print(J, shape, float(y.abs().mean()))
