#!/bin/bash
# usage: [GPUS=N] run_retry.sh <timeout> <command...>   -- retries while the pod answers "transient"/busy (exit 3), nothing is charged then
T=$1; shift
G=""
if [ -n "$GPUS" ] && [ "$GPUS" != "1" ]; then G="--gpus $GPUS"; fi
for i in 1 2 3 4 5 6 7 8 9 10 11 12 13 14 15; do
  /usr/local/graft/bin/gpurun $G --timeout $T -- "$@" > /tmp/gpurun_last.log 2>&1
  rc=$?
  if grep -q "status=transient" /tmp/gpurun_last.log || [ $rc -eq 3 ]; then sleep 120; continue; fi
  break
done
tail -60 /tmp/gpurun_last.log
