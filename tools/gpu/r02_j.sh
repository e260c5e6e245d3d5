#!/bin/bash
set -x
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m pytest tests/test_parallel_nccl_gpu.py -x -q > gpurun_out/r02j_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02j_pytest.log
tail -30 gpurun_out/r02j_pytest.log | cut -c1-300
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py > gpurun_out/r02j_dist.log 2>&1; echo "dist rc=$?"; tail -12 gpurun_out/r02j_dist.log | cut -c1-300
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02j_bench2.json 2> gpurun_out/r02j_bench2.err; echo "bench rc=$?"
tail -5 gpurun_out/r02j_bench2.err | cut -c1-300
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r02j_bench2.json').read().strip().splitlines()[-1])
    print('value',d['value'],'e2e',d['e2e']); print('gather', json.dumps(d.get('gather'), indent=1)); print({k:(v.get('value'),v.get('ms_per_step')) for k,v in d.get('configs',{}).items()})
except Exception as e: print('parse fail', e)
PY
