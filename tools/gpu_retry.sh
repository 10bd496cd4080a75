#!/bin/bash
# usage: [GPUS=N] tools/gpu_retry.sh OUTFILE TIMEOUT 'command'  — retries while the pod answers busy (exit 3)
out=$1; to=$2; shift 2
g=""
if [ -n "$GPUS" ]; then g="--gpus $GPUS"; fi
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun $g --timeout $to -- "$@" > $out 2>&1
  rc=$?
  if [ $rc -ne 3 ] && ! grep -q "status=transient" $out; then exit $rc; fi
  sleep 90
done
exit 3
