#!/bin/bash
# gpurun with retries while the pod is busy (exit code 3 / "transient").  usage: gpurun_retry.sh <logfile> <timeout> [--gpus N] -- '<command>'
LOG=$1; shift; TMO=$1; shift
EXTRA=""
if [ "$1" == "--gpus" ]; then EXTRA="--gpus $2"; shift; shift; fi
shift   # the --
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $TMO $EXTRA -- "$1" > $LOG 2>&1
  if ! grep -q "status=transient" $LOG; then break; fi
  sleep 150
done
