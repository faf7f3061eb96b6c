#!/bin/bash
# retry a gpurun call while the pod answers busy (exit code 3); usage: tools/gpurun_retry.sh LOGFILE TIMEOUT [--gpus N] -- CMD
LOG=$1; shift; TO=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $TO "$@" > $LOG 2>&1
  rc=$?
  if [ $rc -ne 3 ] && ! grep -q "status=transient" $LOG; then exit $rc; fi
  sleep 90
done
exit 3
