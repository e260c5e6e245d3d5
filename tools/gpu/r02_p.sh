#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_autograd2d_gpu.py tests/test_scattering2d_gpu.py tests/test_shape_sweep_gpu.py -x -q > gpurun_out/r02p_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02p_pytest.log
tail -6 gpurun_out/r02p_pytest.log | cut -c1-300
timeout 300 python tools/bwd_bench.py 64 4 224 >> gpurun_out/r02p_bwd.jsonl 2>> gpurun_out/r02p_err.log
timeout 300 python tools/bwd_bench.py 64 3 256 >> gpurun_out/r02p_bwd.jsonl 2>> gpurun_out/r02p_err.log
cat gpurun_out/r02p_bwd.jsonl | cut -c1-900
timeout 200 python tools/kbench.py c1 128 2 32 50 >> gpurun_out/r02p_kbench.jsonl 2>> gpurun_out/r02p_err.log
timeout 200 python tools/kbench.py c2 >> gpurun_out/r02p_kbench.jsonl 2>> gpurun_out/r02p_err.log
python - <<'PY'
import json
for l in open('gpurun_out/r02p_kbench.jsonl'):
    d=json.loads(l); ks=d['kernels']
    print(d['label'][-75:], '%.3f ms %.0f img/s chk %.8e'%(d['ms_median'], d['img_per_s'], d['checksum']))
    print('     ', ' '.join('%s=%.3f'%(k.split(':G')[0],v) for k,v in list(ks.items())[:12]))
PY
tail -3 gpurun_out/r02p_err.log
