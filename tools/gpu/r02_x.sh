#!/bin/bash
mkdir -p gpurun_out
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 3 python tools/sanitize_run.py > gpurun_out/r02x_memcheck.log 2>&1; echo "memcheck rc=$?"
tail -5 gpurun_out/r02x_memcheck.log
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 3 python tools/sanitize_run.py > gpurun_out/r02x_racecheck.log 2>&1; echo "racecheck rc=$?"
tail -5 gpurun_out/r02x_racecheck.log
