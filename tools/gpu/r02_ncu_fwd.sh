#!/bin/bash
mkdir -p gpurun_out
export SCAT_B200_GPU_FILTERS=0     # numpy filter bank: the constructor then launches no k2d_ kernels (keeps the -s count)
python tools/launch_labels.py 256 > gpurun_out/r02n_labels.log 2>&1; tail -1 gpurun_out/r02n_labels.log | cut -c1-400
timeout 900 ncu --set full --clock-control none -k regex:k2d_ -s 24 -c 12 -f -o gpurun_out/r02n_c2_full python tools/run_once.py 256 3 > gpurun_out/r02n_ncu.log 2>&1
tail -2 gpurun_out/r02n_ncu.log
ncu -i gpurun_out/r02n_c2_full.ncu-rep --page raw --csv > gpurun_out/r02n_c2_raw.csv 2>/dev/null
(echo "# ncu --set full --clock-control none -k regex:k2d_ -s 24 -c 12 python tools/run_once.py 256 3   (one C2 forward, batch 256, launch order; SCAT_B200_GPU_FILTERS=0 so that the constructor launches no k2d_ kernels)"; python tools/ncu_summary.py gpurun_out/r02n_c2_full.ncu-rep) > gpurun_out/r02n_ncu_full_summary.txt
rm -f gpurun_out/r02n_c2_full.ncu-rep
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k2d_ -s 24 -c 24 --csv --log-file gpurun_out/r02n_launches.csv python tools/run_once.py 256 4 > /dev/null 2>&1
grep -c "^----" gpurun_out/r02n_ncu_full_summary.txt; wc -l gpurun_out/r02n_launches.csv
