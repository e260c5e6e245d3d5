"""Parity of the C2-shape forward against the committed reference-generated golden (max rel / per-channel L2)."""
import os, sys, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from kymatio_b200 import Scattering2D
from parity import parity_report
d = np.load("tests/golden/golden_2d_c2_J3_256.npz")
S = Scattering2D(3, (256, 256)).cuda()
y = S(torch.from_numpy(d["x"]).cuda()).cpu().numpy()
r = parity_report(y, d["Sx64"])
r["env"] = {k: v for k, v in os.environ.items() if k.startswith("SCAT_B200_") and k != "SCAT_B200_LIB"}
print(json.dumps(r))
