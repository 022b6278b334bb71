#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout_s> <gpus> '<command>'  -- retries while the pod answers "busy" (exit 3)
T=$1; G=$2; CMD=$3
for i in $(seq 1 40); do
  if [ "$G" = "1" ]; then /usr/local/graft/bin/gpurun --timeout $T -- "$CMD"; else /usr/local/graft/bin/gpurun --gpus $G --timeout $T -- "$CMD"; fi
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  echo "[retry] busy, attempt $i; sleeping 60 s"
  sleep 60
done
exit 3
