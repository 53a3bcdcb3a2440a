#!/bin/bash
# submit.sh <timeout_s> <script> [--gpus N]: run a job script under gpurun, retrying while the pod answers "busy" (exit 3)
T=$1; S=$2; shift 2
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun "$@" --timeout $T -- "bash $S"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3
