#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parallel_nccl_gpu.py -x -q > gpurun_out/r02z_pytest2.log 2>&1; echo "pytest rc=$?"
tail -25 gpurun_out/r02z_pytest2.log | cut -c1-300
