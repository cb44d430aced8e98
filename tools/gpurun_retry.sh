#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout_s> <gpus> <command...>  — retries while the pod answers busy
T=$1; G=$2; shift 2
for i in $(seq 1 40); do
  if [ "$G" = "1" ]; then /usr/local/graft/bin/gpurun --timeout $T -- "$@"; else /usr/local/graft/bin/gpurun --gpus $G --timeout $T -- "$@"; fi
  rc=$?
  if [ $rc -ne 3 ] && ! grep -q '"status": "transient"' gpurun_out/.last_call.json 2>/dev/null; then exit $rc; fi
  echo "[retry $i] busy, sleeping 150 s"; sleep 150
done
exit 3
