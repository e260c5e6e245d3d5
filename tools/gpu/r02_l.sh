#!/bin/bash
set -x
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r02l_bench8.json 2> gpurun_out/r02l_bench8.err; echo "bench rc=$?"
tail -5 gpurun_out/r02l_bench8.err | cut -c1-300
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r02l_bench8.json').read().strip().splitlines()[-1])
    print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'ceil',d['e2e'].get('copy_only_ceiling')); print('gather', json.dumps(d.get('gather'), indent=1)); print({k:(v.get('value'),v.get('ms_per_step')) for k,v in d.get('configs',{}).items()})
except Exception as e: print('parse fail', e)
PY
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 tools/dist_check.py > gpurun_out/r02l_dist8.log 2>&1; echo "dist rc=$?"; grep -v "^\*\|OMP" gpurun_out/r02l_dist8.log | tail -5 | cut -c1-300
