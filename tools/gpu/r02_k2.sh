#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_scattering2d_gpu.py tests/test_autograd2d_gpu.py -x -q 2>&1 | tail -2
rm -f gpurun_out/r02k2.json
timeout 200 python tools/kbench.py c2 > gpurun_out/r02k2.json 2>gpurun_out/r02k2.err; tail -2 gpurun_out/r02k2.err
timeout 200 python tools/kbench.py c5 256 4 224 >> gpurun_out/r02k2.json 2>>gpurun_out/r02k2.err
python - <<'PY'
import json
for l in open('gpurun_out/r02k2.json'):
    d=json.loads(l); ks=d['kernels']
    print(d['label'], '%.3f ms %.0f img/s chk %.8e'%(d['ms_median'], d['img_per_s'], d['checksum']))
    print('     ', ' '.join('%s=%.3f'%(k.split(':G')[0],v) for k,v in list(ks.items())[:12]))
PY
