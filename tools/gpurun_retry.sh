#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout-seconds> [--gpus N] <command...>   -- retries while the pod answers "transient"/busy (nothing charged)
T=$1; shift
G=""
if [ "$1" = "--gpus" ]; then G="--gpus $2"; shift 2; fi
for i in $(seq 1 40); do
  out=$(gpurun --timeout $T $G -- "$@" 2>&1)
  if echo "$out" | grep -q "status=transient\|exit code 3\|rc=3"; then sleep 75; continue; fi
  echo "$out"; exit 0
done
echo "$out"; exit 3
