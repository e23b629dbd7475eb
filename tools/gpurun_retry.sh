#!/bin/bash
# usage: tools/gpurun_retry.sh [--gpus N] <timeout-seconds> '<command>'   -- retries while the pod has no free GPU slot
GP=""
if [ "$1" == "--gpus" ]; then GP="--gpus $2"; shift; shift; fi
T=$1; shift
for attempt in $(seq 1 40); do
  out=$(/usr/local/graft/bin/gpurun $GP --timeout $T -- "$@" 2>&1)
  if echo "$out" | grep -q "status=transient"; then sleep 60; continue; fi
  echo "$out"; exit 0
done
echo "gave up: no GPU slot"; exit 3
