#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_autograd2d_gpu.py -x -q > gpurun_out/r02q_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02q_pytest.log
tail -3 gpurun_out/r02q_pytest.log | cut -c1-300
timeout 300 python tools/bwd_bench.py 64 4 224 >> gpurun_out/r02q_bwd.jsonl 2>> gpurun_out/r02q_err.log
timeout 300 python tools/bwd_bench.py 64 3 256 >> gpurun_out/r02q_bwd.jsonl 2>> gpurun_out/r02q_err.log
cat gpurun_out/r02q_bwd.jsonl | cut -c1-900
tail -3 gpurun_out/r02q_err.log
