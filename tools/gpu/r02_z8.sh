#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r02z_bench_8gpu.json 2> gpurun_out/r02z_bench_8gpu.err; echo "bench rc=$?"
tail -3 gpurun_out/r02z_bench_8gpu.err | cut -c1-300
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02z_bench_8gpu.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, 'e2e', d['e2e']['value'], d['e2e'].get('copy_only_ceiling'))
g=d.get('gather') or {}
print({k:g.get(k) for k in ('ms_per_step_with_gather','ms_per_step_local_only','gather_ms')}, (g.get('peer_store') or {}))
print({k:(round(v['ms_per_step'],3), round(v['value'])) for k,v in d['configs'].items()})
PY
