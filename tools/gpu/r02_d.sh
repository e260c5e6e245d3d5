#!/bin/bash
set -x
mkdir -p gpurun_out
run() { env "$@" timeout 200 python tools/kbench.py "$*" >> gpurun_out/r02d_kbench.jsonl 2>> gpurun_out/r02d_kbench.err; }
run SCAT_B200_STAGGER_NS=0
run SCAT_B200_STAGGER_NS=2000
run SCAT_B200_STAGGER_NS=5000
run SCAT_B200_STAGGER_NS=10000
run SCAT_B200_SUPP_THR=1e-6
run SCAT_B200_SUPP_THR=1e-5
run SCAT_B200_SUPP_THR=1e-5 SCAT_B200_STAGGER_NS=5000
for t in 1e-7 1e-6 1e-5 1e-4; do SCAT_B200_SUPP_THR=$t python tools/parity_c2.py >> gpurun_out/r02d_parity.jsonl 2>> gpurun_out/r02d_kbench.err; done
cat gpurun_out/r02d_parity.jsonl
python - <<'PY'
import json
for l in open('gpurun_out/r02d_kbench.jsonl'):
    d=json.loads(l); ks=d['kernels']
    print(d['label'][-75:], '%.3f ms %.0f img/s chk %.8e'%(d['ms_median'], d['img_per_s'], d['checksum']))
    print('     ', ' '.join('%s=%.3f'%(k.split(':G')[0],v) for k,v in list(ks.items())[:12]))
PY
tail -3 gpurun_out/r02d_kbench.err
