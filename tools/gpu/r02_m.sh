#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 tools/dist_check.py > gpurun_out/r02m_dist2.log 2>&1; echo "dist rc=$?"; grep -v "^\*\|OMP\|^$" gpurun_out/r02m_dist2.log | tail -12 | cut -c1-300
