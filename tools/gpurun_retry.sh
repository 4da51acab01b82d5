#!/bin/bash
# Retries a gpurun call while the pod answers "busy" (exit code 3 / status transient; nothing is charged for those).
# usage: tools/gpurun_retry.sh <timeout_s> <gpus> '<command>'
T=$1; G=$2; shift 2
for i in $(seq 1 40); do
  if [ "$G" = "1" ]; then /usr/local/graft/bin/gpurun --timeout "$T" -- "$@"; else /usr/local/graft/bin/gpurun --gpus "$G" --timeout "$T" -- "$@"; fi
  rc=$?
  st=$(python -c "import json;print(json.load(open('/root/repo/gpurun_out/.last_call.json')).get('status'))" 2>/dev/null)
  if [ "$st" != "transient" ]; then exit $rc; fi
  sleep 90
done
exit 3
