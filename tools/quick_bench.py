"""Quick device-resident timing of the headline forward (no L2 flush, for A/B tuning runs)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kymatio_b200 import Scattering2D, _lib
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
S = Scattering2D(3, (256, 256)).cuda()
x = torch.randn(B, 256, 256, device="cuda")
for _ in range(3): S(x)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 10
e0.record()
for _ in range(n): S(x)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
_lib.timing_enable(True); S(x); rows = _lib.timing_report(); _lib.timing_enable(False)
rows.sort(key=lambda r: -r["ms"])
print(f"{os.environ.get('TAG','')} {ms:.3f} ms/step {B/ms*1e3:.0f} img/s | " + " ".join(f"{r['label'].split(':G')[0]}={r['ms']:.2f}" for r in rows[:7]))
