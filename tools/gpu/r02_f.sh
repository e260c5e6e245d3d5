#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02f_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02f_pytest.log
tail -8 gpurun_out/r02f_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02f_bench.json 2> gpurun_out/r02f_bench.err; echo "bench rc=$?"
tail -5 gpurun_out/r02f_bench.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r02f_bench.json').read().strip().splitlines()[-1])
    print('value',d['value'],'e2e',d['e2e']['value'],'ceil',d['e2e'].get('copy_only_ceiling'), 'frontend', d['config']['frontend'])
    for k,v in d.get('configs',{}).items(): print(k, {a:b for a,b in v.items() if a in('value','unit','ms_per_step','pass_model_frac','parity_max_rel_vs_golden','graph_equals_eager')})
    print(d.get('cpu_baseline'), d.get('reference_torch_gpu'))
except Exception as e: print('parse fail', e)
PY
python tools/launch_labels.py 256 > gpurun_out/r02f_labels.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:k2d_ -s 24 -c 12 -f -o gpurun_out/r02f_c2_full python tools/run_once.py 256 3 > gpurun_out/r02f_ncu.log 2>&1
ncu -i gpurun_out/r02f_c2_full.ncu-rep --page raw --csv > gpurun_out/r02f_c2_raw.csv 2>/dev/null
rm -f gpurun_out/r02f_c2_full.ncu-rep
ncu --metrics gpu__time_duration.sum --clock-control none -s 24 -c 24 --csv --log-file gpurun_out/r02f_launches.csv python tools/run_once.py 256 4 > /dev/null 2>&1
ls -la gpurun_out
