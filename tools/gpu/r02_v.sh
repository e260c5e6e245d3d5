#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/r02v_bench.json 2> gpurun_out/r02v_bench.err; echo "bench rc=$?"
tail -3 gpurun_out/r02v_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02v_bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'])
print({k:(round(v['ms_per_step'],3), round(v['value'])) for k,v in d['configs'].items()})
PY
nvidia-smi --query-gpu=memory.used,memory.total --format=csv
