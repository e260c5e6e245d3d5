#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/kbench13.py c3 > gpurun_out/r02t_c3.json 2> gpurun_out/r02t_err.log
timeout 300 python tools/kbench13.py c4 > gpurun_out/r02t_c4.json 2>> gpurun_out/r02t_err.log
tail -3 gpurun_out/r02t_err.log
python - <<'PY'
import json
for f in ('c3','c4'):
    d=json.loads(open(f'gpurun_out/r02t_{f}.json').read().strip().splitlines()[-1])
    print(f, d['B'], 'ms', round(d['ms_median'],2), 'lib', round(d['lib_ms'],2))
    for k,v in d['kernels'].items(): print('   %-40s %8.3f ms  n=%d  %7.0f GB/s'%(k, v['ms'], v['n'], v['GBps']))
PY
