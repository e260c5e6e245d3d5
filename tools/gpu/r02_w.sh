#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k2d_colpass_imrf -s 1 -c 1 -f -o gpurun_out/r02w_imrf python tools/run_once.py 256 3 > gpurun_out/r02w_ncu.log 2>&1
tail -2 gpurun_out/r02w_ncu.log
bash tools/ncu_export.sh gpurun_out/r02w_imrf.ncu-rep gpurun_out/r02w_imrf
ls -la gpurun_out | grep r02w
