#!/bin/bash
# round 2, GPU call A: new parity-gate tests + phase profile + ncu full of every kernel of one C2 forward
set -x
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r02a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02a_pytest.log
tail -5 gpurun_out/r02a_pytest.log
SCAT_B200_LIB=$PWD/kymatio_b200/lib/libscat_b200_prof.so timeout 300 python tools/phase_prof.py 256 3 256 > gpurun_out/r02a_phase_c2.log 2>&1
SCAT_B200_LIB=$PWD/kymatio_b200/lib/libscat_b200_prof.so timeout 300 python tools/phase_prof.py 64 4 224 > gpurun_out/r02a_phase_c5.log 2>&1
cat gpurun_out/r02a_phase_c2.log gpurun_out/r02a_phase_c5.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k2d_ -s 24 -c 12 -f -o gpurun_out/r02a_c2_full python tools/run_once.py 256 3 > gpurun_out/r02a_ncu.log 2>&1
tail -3 gpurun_out/r02a_ncu.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err
tail -c 600 gpurun_out/r02a_bench.json
ls -la gpurun_out | tail -8
