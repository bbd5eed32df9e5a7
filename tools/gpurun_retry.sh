#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout> <out-file> <command...> : retries while the pod answers "busy" (exit 3), nothing is charged for those
t=$1; out=$2; shift 2
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun --timeout $t -- "$@" > $out 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then echo "gpurun rc=$rc after $i tries" >> $out; exit $rc; fi
  sleep 170
done
exit 3
