#!/bin/bash
# usage: tools/gpu/submit.sh <timeout_s> <script> [gpus]  - retries while the pod answers busy (exit 3)
T=$1; S=$2; G=${3:-1}
for i in $(seq 1 40); do
  if [ "$G" = "1" ]; then /usr/local/graft/bin/gpurun --timeout $T -- "bash $S"; else /usr/local/graft/bin/gpurun --gpus $G --timeout $T -- "bash $S"; fi
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3
