#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_scattering2d_gpu.py tests/test_autograd2d_gpu.py tests/test_kymatio_plugin_gpu.py tests/test_shape_sweep_gpu.py -x -q 2>&1 | tail -3
rm -f gpurun_out/r02k3.json
timeout 200 python tools/kbench.py c1 128 2 32 50 >> gpurun_out/r02k3.json 2>gpurun_out/r02k3.err; tail -2 gpurun_out/r02k3.err
python - <<'PY'
import json
for l in open('gpurun_out/r02k3.json'):
    d=json.loads(l); ks=d['kernels']
    print(d['label'], '%.3f ms %.0f img/s chk %.8e'%(d['ms_median'], d['img_per_s'], d['checksum']))
    print('     ', ' '.join('%s=%.3f'%(k.split(':G')[0],v) for k,v in list(ks.items())[:12]))
PY
timeout 300 python tools/bwd_bench.py 128 2 32 2>>gpurun_out/r02k3.err | cut -c1-500
