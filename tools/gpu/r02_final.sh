#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02f_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02f_pytest.log
tail -4 gpurun_out/r02f_pytest.log | cut -c1-300
timeout 900 python bench.py > gpurun_out/r02f_bench.json 2> gpurun_out/r02f_bench.err; echo "bench rc=$?"
tail -2 gpurun_out/r02f_bench.err | cut -c1-300
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02f_bench_ref.json 2>> gpurun_out/r02f_bench.err; echo "ref rc=$?"
cut -c1-600 gpurun_out/r02f_bench_ref.json
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02f_bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['e2e'].get('copy_only_ceiling'), d['clocks'])
print({k:(round(v['ms_per_step'],3), round(v['value']), v.get('parity_max_rel_vs_golden')) for k,v in d['configs'].items()})
print(d['roofline']['frac'], d['roofline']['step_pass_model'], d['cpu_baseline'])
PY
SCAT_B200_LIB=kymatio_b200/lib/libscat_b200_prof.so timeout 300 python tools/phase_prof_bwd.py 64 4 224 2>&1 | tail -6
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
