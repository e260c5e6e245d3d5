#!/bin/bash
mkdir -p gpurun_out
for cfg in "1 2" "2 2" "1 3" "2 3"; do
  set -- $cfg
  BENCH_E2E_CHUNKS=$1 BENCH_E2E_STREAMS=$2 timeout 300 python bench.py --no-configs --steps 20 --warmup 3 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('chunks $1 streams $2', 'value %.0f e2e %.0f ceil %.0f'%(d['value'], d['e2e']['value'], d['e2e'].get('copy_only_ceiling',0)))"
done
