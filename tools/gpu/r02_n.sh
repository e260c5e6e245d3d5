#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29540 tools/dist_check.py > gpurun_out/r02n_dist8.log 2>&1; echo "dist rc=$?"; grep -v "^\*\|OMP\|^$" gpurun_out/r02n_dist8.log | tail -6 | cut -c1-300
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 tools/gather_bench.py 256 > gpurun_out/r02n_gather8.log 2>&1; echo "rc=$?"; grep -v "^\*\|OMP\|^$" gpurun_out/r02n_gather8.log | tail -6 | cut -c1-700
