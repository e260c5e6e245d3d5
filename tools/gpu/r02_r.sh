#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_filters_gpu.py tests/test_kymatio_plugin_gpu.py tests/test_scattering3d_gpu.py -x -q -s > gpurun_out/r02r_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02r_pytest.log
tail -30 gpurun_out/r02r_pytest.log | cut -c1-400
timeout 300 python - <<'PY' 2>&1 | tail -20
import time, torch, numpy as np
from kymatio_b200.filter_bank_gpu import filter_bank_2d_gpu, solid_harmonic_filter_bank_gpu, gaussian_filter_bank_gpu
from kymatio_b200.filter_bank2d import filter_bank_2d
from kymatio_b200 import _lib
for cfg in [(272,272,3,8),(256,256,4,8),(1040,1040,4,8)]:
    filter_bank_2d_gpu(*cfg); torch.cuda.synchronize()
    t0=time.perf_counter(); fb=filter_bank_2d_gpu(*cfg); torch.cuda.synchronize(); t1=time.perf_counter()
    if cfg[0] <= 272:
        t2=time.perf_counter(); filter_bank_2d.__wrapped__ if False else None
        from kymatio_b200.filter_bank2d import _filter_bank_cached
        _filter_bank_cached.cache_clear(); filter_bank_2d(*cfg); t3=time.perf_counter()
        print(cfg, 'gpu %.1f ms  numpy(product, vectorised) %.1f ms'%((t1-t0)*1e3,(t3-t2)*1e3))
    else:
        print(cfg, 'gpu %.1f ms'%((t1-t0)*1e3))
solid_harmonic_filter_bank_gpu(128,128,128,2,2,1.0); torch.cuda.synchronize()
t0=time.perf_counter(); b=solid_harmonic_filter_bank_gpu(128,128,128,2,2,1.0); g=gaussian_filter_bank_gpu(128,128,128,3,1.0); torch.cuda.synchronize(); t1=time.perf_counter()
print('3-D 128^3 J=2 L=2 bank on device: %.1f ms (%.0f MB)'%((t1-t0)*1e3, sum(x.numel()*8 for x in b)/1e6))
PY
