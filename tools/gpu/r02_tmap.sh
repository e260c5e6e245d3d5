#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/r02tm.json
SCAT_B200_IMRF_TMAP=0 timeout 200 python tools/kbench.py c5_off 256 4 224 >> gpurun_out/r02tm.json 2>gpurun_out/r02tm.err
SCAT_B200_IMRF_TMAP=1 timeout 200 python tools/kbench.py c5_on 256 4 224 >> gpurun_out/r02tm.json 2>>gpurun_out/r02tm.err
timeout 200 python tools/kbench.py c2 >> gpurun_out/r02tm.json 2>>gpurun_out/r02tm.err
tail -3 gpurun_out/r02tm.err
python - <<'PY'
import json
for l in open('gpurun_out/r02tm.json'):
    d=json.loads(l); ks=d['kernels']
    print(d['label'], '%.3f ms %.0f img/s chk %.10e'%(d['ms_median'], d['img_per_s'], d['checksum']))
    print('     ', ' '.join('%s=%.3f'%(k.split(':G')[0],v) for k,v in list(ks.items())[:8]))
PY
