#!/bin/bash
# round 2, GPU call B: L2 bulk prefetch + two-half leaf tile: parity, A/B timing, phase profile, ncu export
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_scattering2d_gpu.py tests/test_shape_sweep_gpu.py tests/test_autograd2d_gpu.py tests/test_kymatio_plugin_gpu.py -x -q > gpurun_out/r02b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02b_pytest.log
tail -5 gpurun_out/r02b_pytest.log
run() { env "$@" timeout 200 python tools/kbench.py "$*" >> gpurun_out/r02b_kbench.jsonl 2>> gpurun_out/r02b_kbench.err; }
run SCAT_B200_PREFETCH=0 SCAT_B200_TILE2H=0
run SCAT_B200_PREFETCH=1 SCAT_B200_TILE2H=0
run SCAT_B200_PREFETCH=1 SCAT_B200_TILE2H=1 SCAT_B200_TILE2H_THREADS=384
run SCAT_B200_PREFETCH=1 SCAT_B200_TILE2H=1 SCAT_B200_TILE2H_THREADS=320
run SCAT_B200_PREFETCH=1 SCAT_B200_TILE2H=1 SCAT_B200_TILE2H_THREADS=288
run SCAT_B200_PREFETCH=0 SCAT_B200_TILE2H=1 SCAT_B200_TILE2H_THREADS=384
run SCAT_B200_PREFETCH=1 SCAT_B200_TILE2H=2 SCAT_B200_TILE2H_THREADS=384
run SCAT_B200_PREFETCH=1 SCAT_B200_TILE2H=2 SCAT_B200_TILE2H_THREADS=256
python - <<'PY'
import json
for l in open('gpurun_out/r02b_kbench.jsonl'):
    d=json.loads(l); ks=d['kernels']
    print(d['label'][-60:], '%.3f ms %.0f img/s chk %.6e'%(d['ms_median'], d['img_per_s'], d['checksum']), ' '.join('%s=%.3f'%(k.split(':G')[0],v) for k,v in list(ks.items())[:8]))
PY
SCAT_B200_LIB=$PWD/kymatio_b200/lib/libscat_b200_prof.so SCAT_B200_TILE2H=0 timeout 300 python tools/phase_prof.py 256 3 256 > gpurun_out/r02b_phase_pf.log 2>&1
SCAT_B200_LIB=$PWD/kymatio_b200/lib/libscat_b200_prof.so SCAT_B200_TILE2H=2 timeout 300 python tools/phase_prof.py 256 3 256 > gpurun_out/r02b_phase_2h.log 2>&1
cat gpurun_out/r02b_phase_pf.log gpurun_out/r02b_phase_2h.log
SCAT_B200_TILE2H=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k2d_ -s 24 -c 12 -f -o gpurun_out/r02b_c2_full python tools/run_once.py 256 3 > gpurun_out/r02b_ncu.log 2>&1
tail -2 gpurun_out/r02b_ncu.log
bash tools/ncu_export.sh gpurun_out/r02b_c2_full.ncu-rep gpurun_out/r02b_c2
SCAT_B200_TILE2H=2 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k2d_tile2h -s 6 -c 3 -f -o gpurun_out/r02b_2h_full python tools/run_once.py 256 3 > gpurun_out/r02b_ncu2.log 2>&1
bash tools/ncu_export.sh gpurun_out/r02b_2h_full.ncu-rep gpurun_out/r02b_2h
ls -la gpurun_out; du -sh gpurun_out
