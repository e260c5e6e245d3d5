#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02k_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02k_pytest.log
tail -6 gpurun_out/r02k_pytest.log | cut -c1-300
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02k_bench.json 2> gpurun_out/r02k_bench.err; echo "bench rc=$?"
tail -5 gpurun_out/r02k_bench.err | cut -c1-300
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02k_bench_ref.json 2>> gpurun_out/r02k_bench.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r02k_bench.json').read().strip().splitlines()[-1])
    print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'ceil',d['e2e'].get('copy_only_ceiling'))
    for k,v in d.get('configs',{}).items(): print(k, {a:b for a,b in v.items() if a in('value','unit','ms_per_step','pass_model_frac','parity_max_rel_vs_golden','graph_equals_eager')})
    for k in d['kernels']: print(k['label'], round(k['ms_per_step'],3), round(k['pass_model_frac'],2), k.get('dram_frac'), k.get('binding'))
    print(json.dumps(d['roofline'])[:1500])
except Exception as e: print('parse fail', e)
PY
cat gpurun_out/r02k_bench_ref.json | cut -c1-400
for f in 1; do SCAT_B200_ORDER1_FUSED=$f timeout 300 python tools/bwd_bench.py 64 4 224 >> gpurun_out/r02k_bwd.jsonl 2>> gpurun_out/r02k_bench.err; done
cat gpurun_out/r02k_bwd.jsonl | cut -c1-900
python tools/launch_labels.py 256 > gpurun_out/r02k_labels.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:k2d_ -s 24 -c 12 -f -o gpurun_out/r02k_c2_full python tools/run_once.py 256 3 > gpurun_out/r02k_ncu.log 2>&1
ncu -i gpurun_out/r02k_c2_full.ncu-rep --page raw --csv > gpurun_out/r02k_c2_raw.csv 2>/dev/null
rm -f gpurun_out/r02k_c2_full.ncu-rep
ncu --metrics gpu__time_duration.sum --clock-control none -s 24 -c 24 --csv --log-file gpurun_out/r02k_launches.csv python tools/run_once.py 256 4 > /dev/null 2>&1
cuobjdump -sass kymatio_b200/lib/libscat_b200.so | grep -oE "UBLKCP[A-Z.]*|UBLKPF[A-Z.0-9]*|SYNCS[A-Z.0-9]*|UTMALDG|LDGSTS|HMMA[A-Z.0-9]*|FFMA2|FADD2|UTC[A-Z]*MMA" | sort | uniq -c > gpurun_out/r02k_sass_mnemonics.txt; cat gpurun_out/r02k_sass_mnemonics.txt
ls -la gpurun_out | tail
