#!/bin/bash
# usage: tools/gpu_retry.sh <timeout_s> <logfile> <command...>   — gpurun with retries while the pod answers "busy" (exit 3)
t=$1; log=$2; shift 2
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $t -- "$@" > $log 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then echo "gpurun rc=$rc (try $i)" >> $log; exit $rc; fi
  sleep 45
done
echo "gpurun: gave up after 40 busy answers" >> $log
exit 3
