"""First-contact diagnostics on the GPU box: parity per golden case (no early exit) + rough timing."""
import glob
import os
import sys
import time
import traceback

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from kymatio_b200 import Scattering2D, _lib  # noqa: E402
from parity import parity_report  # noqa: E402

print(torch.cuda.get_device_name(0), "lib version", _lib.load().scat_version())
for f in sorted(glob.glob(os.path.join(ROOT, "tests/golden/golden_2d_*.npz"))):
    d = np.load(f)
    name = os.path.basename(f)[10:-4]
    try:
        S = Scattering2D(int(d["J"]), tuple(int(v) for v in d["shape"]), L=int(d["L"]),
                         max_order=int(d["max_order"]), pre_pad=bool(d["pre_pad"])).cuda()
        y = S(torch.from_numpy(d["x"]).cuda())
        torch.cuda.synchronize()
        y = y.cpu().numpy()
        r = parity_report(y, d["Sx64"])
        L, J = int(d["L"]), int(d["J"])
        r0 = parity_report(y[:, :1], d["Sx64"][:, :1])["max_rel"]
        r1 = parity_report(y[:, 1:1 + L * J], d["Sx64"][:, 1:1 + L * J])["max_rel"]
        r2 = parity_report(y[:, 1 + L * J:], d["Sx64"][:, 1 + L * J:])["max_rel"] if y.shape[1] > 1 + L * J else 0
        print(f"{name:16s} {str(y.shape):20s} max_rel={r['max_rel']:.2e} chan_l2={r['chan_l2_max']:.2e}"
              f" @ch{r['chan_argmax']}  S0={r0:.1e} S1={r1:.1e} S2={r2:.1e}", flush=True)
    except Exception:
        print(name, "FAILED")
        traceback.print_exc()

if "--time" in sys.argv:
    for (J, shape, B) in [(3, (256, 256), 64), (3, (256, 256), 256), (2, (32, 32), 128)]:
        S = Scattering2D(J, shape).cuda()
        x = torch.randn(B, *shape, device="cuda")
        for _ in range(3):
            S(x)
        torch.cuda.synchronize()
        t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
        n = 5
        t0.record()
        for _ in range(n):
            S(x)
        t1.record(); torch.cuda.synchronize()
        ms = t0.elapsed_time(t1) / n
        print(f"J={J} shape={shape} B={B}: {ms:.2f} ms/batch  {B / ms * 1e3:.0f} img/s", flush=True)
