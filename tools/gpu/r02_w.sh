#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k2d_tile -s 4 -c 1 -f -o gpurun_out/r02w_tile python tools/run_once.py 256 3 > gpurun_out/r02w_ncu.log 2>&1
tail -3 gpurun_out/r02w_ncu.log
bash tools/ncu_export.sh gpurun_out/r02w_tile.ncu-rep gpurun_out/r02w_tile
ls -la gpurun_out | grep r02w
