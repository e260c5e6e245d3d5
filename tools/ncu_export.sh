#!/bin/bash
# usage: tools/ncu_export.sh <rep.ncu-rep> <out-prefix>: raw metrics csv + per-kernel source csv (gzip), then drop the .ncu-rep
# if it is too big to travel back (gpurun_out is capped at 64 MiB)
REP=$1; OUT=$2
ncu -i $REP --page raw --csv > ${OUT}_raw.csv 2>/dev/null
ncu -i $REP --page source --csv --print-source cuda,sass > ${OUT}_source.csv 2>/dev/null
gzip -f ${OUT}_source.csv
SZ=$(stat -c %s $REP)
if [ $SZ -gt 30000000 ]; then rm -f $REP; fi
