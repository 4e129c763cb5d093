#!/bin/bash
# usage: tools/gpu_try.sh <timeout> <gpus> '<command>'  -- retries gpurun while the pod answers busy (rc 3)
T=$1; G=$2; shift 2
for i in $(seq 1 40); do
  if [ "$G" = "1" ]; then /usr/local/graft/bin/gpurun --timeout $T -- "$@" > /tmp/gpu_try.log 2>&1; else /usr/local/graft/bin/gpurun --gpus $G --timeout $T -- "$@" > /tmp/gpu_try.log 2>&1; fi
  rc=$?
  if grep -q "status=transient" /tmp/gpu_try.log || [ $rc -eq 3 ]; then sleep 90; continue; fi
  break
done
tail -40 /tmp/gpu_try.log
