"""Kernel labels of one C2 forward in LAUNCH order (-> gpurun_out/launch_labels.json), to pair the rows of an ncu capture
of tools/run_once.py with the labels bench.py reports."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kymatio_b200 import Scattering2D, _lib
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
S = Scattering2D(3, (256, 256)).cuda()
x = torch.randn(B, 256, 256, device="cuda")
S(x)
_lib.timing_enable(True)
S(x)
rows = _lib.timing_report()
_lib.timing_enable(False)
os.makedirs("gpurun_out", exist_ok=True)
json.dump({"batch": B, "labels": [r["label"] for r in rows]}, open("gpurun_out/launch_labels.json", "w"))
print([r["label"] for r in rows])
