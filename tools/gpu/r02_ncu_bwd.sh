#!/bin/bash
mkdir -p gpurun_out
# second step's backward kernels (first step = 16 tile_bwd/bwd_col/bwd_row/tile_adj/hlow launches)
timeout 900 ncu --set full --clock-control none -k regex:"k2d_tile_bwd|k2d_bwd_col|k2d_bwd_row|k2d_tile_adj|k2d_hlow_bwd" -s 16 -c 16 -f -o gpurun_out/r02_bwd_full python tools/run_once_bwd.py 64 2 > gpurun_out/r02_ncu_bwd.log 2>&1
tail -2 gpurun_out/r02_ncu_bwd.log
(echo "# ncu --set full --clock-control none -k regex:k2d_tile_bwd|k2d_bwd_col|k2d_bwd_row|k2d_tile_adj|k2d_hlow_bwd -s 16 -c 16 python tools/run_once_bwd.py 64 2   (backward kernels of one C5-shape step, batch 64)"; python tools/ncu_summary.py gpurun_out/r02_bwd_full.ncu-rep) > gpurun_out/r02_ncu_bwd_summary.txt
rm -f gpurun_out/r02_bwd_full.ncu-rep
grep -c "^----" gpurun_out/r02_ncu_bwd_summary.txt
