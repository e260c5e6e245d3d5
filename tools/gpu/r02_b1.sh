#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/r02f_bench.json 2> gpurun_out/r02f_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02f_bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'])
for k in d['kernels'][:7]: print(k['label'], round(k['ms_per_step'],3), round(k['pass_model_frac'],2), round(k.get('dram_frac') or 0,2), k.get('issue_frac'), k.get('l1tex_frac'), k.get('binding'))
print({k:(round(v['ms_per_step'],3), round(v['value'])) for k,v in d['configs'].items()})
PY
