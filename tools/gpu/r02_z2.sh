#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_parallel_nccl_gpu.py -x -q > gpurun_out/r02z_pytest2.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/r02z_pytest2.log | cut -c1-300
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02z_bench_2gpu.json 2> gpurun_out/r02z_bench_2gpu.err; echo "bench rc=$?"
tail -3 gpurun_out/r02z_bench_2gpu.err | cut -c1-300
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02z_bench_2gpu.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, d['e2e']['value'], d.get('gather'))
print({k:(round(v['ms_per_step'],3), round(v['value'])) for k,v in d['configs'].items()})
PY
