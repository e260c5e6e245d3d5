#!/bin/bash
mkdir -p gpurun_out
SCAT_B200_LIB=kymatio_b200/lib/libscat_b200_prof.so timeout 300 python tools/phase_prof_bwd.py 64 4 224 2>&1 | tail -12
SCAT_B200_LIB=kymatio_b200/lib/libscat_b200_prof.so timeout 300 python tools/phase_prof_bwd.py 64 3 256 2>&1 | tail -12
