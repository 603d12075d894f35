#!/bin/bash
# usage: tools/gpu_retry.sh <logfile> [--gpus N] <timeout_s> '<command>'   -- retries while the pod answers busy (rc 3)
log=$1; shift
gpus=""
if [ "$1" == "--gpus" ]; then gpus="--gpus $2"; shift 2; fi
to=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun $gpus --timeout $to -- "$@" > $log 2>&1
  rc=$?
  if [ $rc -ne 3 ] && ! grep -q "status=transient" $log; then exit $rc; fi
  sleep 45
done
exit 3
