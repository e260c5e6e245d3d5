import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kymatio_b200 import Scattering2D, _lib
J, shape, B = (4, (224, 224), 64) if len(sys.argv) < 2 else (3, (256, 256), 64)
S = Scattering2D(J, shape).cuda()
def fb():
    x = torch.randn(B, *shape, device="cuda", requires_grad=True)
    S(x).sum().backward()
for _ in range(2): fb()
torch.cuda.synchronize()
_lib.timing_enable(True); fb(); rows = _lib.timing_report(); _lib.timing_enable(False)
rows.sort(key=lambda r: -r["ms"])
tot = sum(r["ms"] for r in rows)
print("library kernels total %.2f ms" % tot)
for r in rows[:22]: print("  %-40s x%-3d %8.3f ms" % (r["label"], r["count"], r["ms"]))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); fb(); e1.record(); torch.cuda.synchronize()
print("fwd+bwd wall (device) %.2f ms" % e0.elapsed_time(e1))
