#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_autograd2d_gpu.py -x -q > gpurun_out/r02i_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02i_pytest.log
tail -15 gpurun_out/r02i_pytest.log
cat > /tmp/bwd_bench.py <<'PY'
import os, sys, time, json, torch
sys.path.insert(0, os.getcwd())
from kymatio_b200 import Scattering2D, _lib
B = int(sys.argv[1]); J = int(sys.argv[2]); N = int(sys.argv[3])
S = Scattering2D(J, (N, N)).cuda()
x = torch.randn(B, N, N, device="cuda")
def step():
    xi = x.detach().requires_grad_(True)
    S(xi).sum().backward()
    return xi.grad
for _ in range(2): step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3): g = step()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
_lib.timing_enable(True); step(); rows = _lib.timing_report(); _lib.timing_enable(False)
rows.sort(key=lambda r: -r["ms"])
print(json.dumps({"env": os.environ.get("SCAT_B200_ORDER1_FUSED", "1"), "B": B, "J": J, "N": N, "ms": ms, "img_per_s": B / ms * 1e3, "lib_ms": sum(r["ms"] for r in rows),
                  "top": {r["label"]: round(r["ms"], 3) for r in rows[:14]}}))
PY
for f in 0 1; do SCAT_B200_ORDER1_FUSED=$f timeout 300 python /tmp/bwd_bench.py 64 4 224 >> gpurun_out/r02i_bwd.jsonl 2>> gpurun_out/r02i_err.log; done
for f in 0 1; do SCAT_B200_ORDER1_FUSED=$f timeout 300 python /tmp/bwd_bench.py 64 3 256 >> gpurun_out/r02i_bwd.jsonl 2>> gpurun_out/r02i_err.log; done
cat gpurun_out/r02i_bwd.jsonl; tail -5 gpurun_out/r02i_err.log
timeout 200 python tools/kbench.py base >> gpurun_out/r02i_kbench.jsonl 2>> gpurun_out/r02i_err.log
python - <<'PY'
import json
for l in open('gpurun_out/r02i_kbench.jsonl'):
    d=json.loads(l); ks=d['kernels']
    print(d['label'][-75:], '%.3f ms %.0f img/s chk %.8e'%(d['ms_median'], d['img_per_s'], d['checksum']))
    print('     ', ' '.join('%s=%.3f'%(k.split(':G')[0],v) for k,v in list(ks.items())[:12]))
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k2d_tile -s 10 -c 1 -f -o gpurun_out/r02i_tile python tools/run_once.py 256 3 > gpurun_out/r02i_ncu.log 2>&1
bash tools/ncu_export.sh gpurun_out/r02i_tile.ncu-rep gpurun_out/r02i_tile
ls -la gpurun_out | tail -5
