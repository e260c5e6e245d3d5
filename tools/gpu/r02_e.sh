#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 120 python tools/parity_c2.py > gpurun_out/r02e_parity.jsonl 2> gpurun_out/r02e_err.log; echo "parity rc=$?"
cat gpurun_out/r02e_parity.jsonl; tail -3 gpurun_out/r02e_err.log
timeout 600 python -m pytest tests/test_scattering2d_gpu.py tests/test_shape_sweep_gpu.py tests/test_autograd2d_gpu.py tests/test_kymatio_plugin_gpu.py -x -q > gpurun_out/r02e_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02e_pytest.log
tail -5 gpurun_out/r02e_pytest.log
run() { env "$@" timeout 200 python tools/kbench.py "$*" >> gpurun_out/r02e_kbench.jsonl 2>> gpurun_out/r02e_err.log; }
run SCAT_B200_TMA=0
run SCAT_B200_TMA=1
run SCAT_B200_TMA=1 SCAT_B200_TMA_CTAS=2
run SCAT_B200_TMA=1 SCAT_B200_TMA_CTAS=1
python - <<'PY'
import json
for l in open('gpurun_out/r02e_kbench.jsonl'):
    d=json.loads(l); ks=d['kernels']
    print(d['label'][-75:], '%.3f ms %.0f img/s chk %.8e'%(d['ms_median'], d['img_per_s'], d['checksum']))
    print('     ', ' '.join('%s=%.3f'%(k.split(':G')[0],v) for k,v in list(ks.items())[:12]))
PY
tail -3 gpurun_out/r02e_err.log
