#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02s_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02s_pytest.log
tail -8 gpurun_out/r02s_pytest.log | cut -c1-300
timeout 900 python bench.py > gpurun_out/r02s_bench.json 2> gpurun_out/r02s_bench.err; echo "bench rc=$?"
tail -3 gpurun_out/r02s_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02s_bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e'], d['clocks'])
print({k:(round(v['ms_per_step'],3), round(v['value'])) for k,v in d['configs'].items()})
print(d['roofline']['frac'], d['roofline']['step_pass_model'], d['cpu_baseline'])
PY
