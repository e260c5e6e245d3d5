"""Small forward/backward runs for compute-sanitizer (memcheck / racecheck / initcheck): the 2-D forward and the fused
backward blocks at the config shapes (C2: 272-padded, C5: 256-padded static tiles, streaming first-order chain, kept
spectra), generic sizes, and the device-side filter synthesis."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kymatio_b200 import Scattering2D
from kymatio_b200.filter_bank_gpu import solid_harmonic_filter_bank_gpu, gaussian_filter_bank_gpu
for (J, shape, B, bwd) in [(2, (32, 32), 2, True), (3, (64, 64), 1, True), (2, (33, 47), 1, True), (3, (256, 256), 1, True),
                           (4, (224, 224), 1, True), (3, (256, 256), 2, False)]:
    S = Scattering2D(J, shape).cuda()          # filters synthesised on the device
    x = torch.randn(B, *shape, device="cuda", requires_grad=bwd)
    y = S(x)
    if bwd:
        y.sum().backward()
    torch.cuda.synchronize()
    print(J, shape, float(y.abs().mean()), float(x.grad.abs().mean()) if bwd else None)
b = solid_harmonic_filter_bank_gpu(9, 8, 12, 1, 3, 1.0)
g = gaussian_filter_bank_gpu(9, 8, 12, 2, 1.0)
torch.cuda.synchronize()
print("filters3d", [tuple(t.shape) for t in b], tuple(g.shape))
