#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_scattering1d_gpu.py -x -q > gpurun_out/r02o_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02o_pytest.log
tail -25 gpurun_out/r02o_pytest.log | cut -c1-300
