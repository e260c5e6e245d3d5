#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 120 python tools/parity_c2.py > gpurun_out/r02g_parity.jsonl 2> gpurun_out/r02g_err.log; cat gpurun_out/r02g_parity.jsonl; tail -3 gpurun_out/r02g_err.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02g_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02g_pytest.log
tail -6 gpurun_out/r02g_pytest.log
run() { env "$@" timeout 200 python tools/kbench.py "$*" >> gpurun_out/r02g_kbench.jsonl 2>> gpurun_out/r02g_err.log; }
run SCAT_B200_X=0
run SCAT_B200_TILE_THREADS=640
run SCAT_B200_TILE_THREADS=576
run SCAT_B200_TILE_THREADS=544
timeout 200 python tools/kbench.py c5 256 4 224 >> gpurun_out/r02g_kbench.jsonl 2>> gpurun_out/r02g_err.log
python - <<'PY'
import json
for l in open('gpurun_out/r02g_kbench.jsonl'):
    d=json.loads(l); ks=d['kernels']
    print(d['label'][-75:], '%.3f ms %.0f img/s chk %.8e'%(d['ms_median'], d['img_per_s'], d['checksum']))
    print('     ', ' '.join('%s=%.3f'%(k.split(':G')[0],v) for k,v in list(ks.items())[:12]))
PY
SCAT_B200_LIB=$PWD/kymatio_b200/lib/libscat_b200_prof.so timeout 300 python tools/phase_prof.py 256 3 256 > gpurun_out/r02g_phase.log 2>&1; cat gpurun_out/r02g_phase.log
tail -3 gpurun_out/r02g_err.log
