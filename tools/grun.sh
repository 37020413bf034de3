#!/bin/bash
# gpurun with retries while the pod is busy (exit code 3 = nothing charged).
# usage: [GPUS=2] tools/grun.sh <timeout> '<command>'
t=$1; shift
extra=""
if [ -n "$GPUS" ]; then extra="--gpus $GPUS"; fi
for i in $(seq 1 60); do
  /usr/local/graft/bin/gpurun --timeout "$t" $extra -- "$@"
  rc=$?
  if [ $rc -ne 3 ] && [ $rc -ne 2 ]; then exit $rc; fi
  sleep 45
done
exit 3
