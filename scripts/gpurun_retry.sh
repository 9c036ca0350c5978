#!/bin/bash
# Development aid: run a gpurun call, retrying while the pod answers "busy" (exit code 3).
# usage: [GPUS=2] gpurun_retry.sh LOG TIMEOUT 'command'
log=$1; shift; to=$1; shift
for i in $(seq 1 20); do
  if [ -n "$GPUS" ]; then gpurun --gpus "$GPUS" --timeout "$to" -- "$@" > "$log" 2>&1; else gpurun --timeout "$to" -- "$@" > "$log" 2>&1; fi
  rc=$?
  if [ $rc -ne 3 ] && ! grep -q "status=transient" "$log"; then exit $rc; fi
  sleep 90
done
