#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout-s> <logfile> <command...>   -- retries while the pod answers busy (exit 3), nothing is charged for those
T=$1; LOG=$2; shift 2
for attempt in 1 2 3 4 5 6 7 8 9 10 11 12; do
  /usr/local/graft/bin/gpurun --timeout $T -- "$@" > $LOG 2>&1
  rc=$?
  if grep -q "status=transient" $LOG || [ $rc -eq 3 ]; then sleep 90; continue; fi
  break
done
tail -30 $LOG
