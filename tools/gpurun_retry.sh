#!/bin/bash
# gpurun with retries while the pod answers "busy" (exit code 3 / status=transient): usage tools/gpurun_retry.sh <timeout> '<command>'
to=$1; shift
for i in $(seq 1 20); do
  out=$(/usr/local/graft/bin/gpurun --timeout "$to" -- "$@" 2>&1)
  if echo "$out" | grep -q "status=transient"; then sleep 90; continue; fi
  echo "$out"; exit 0
done
echo "$out"; echo "gave up after 20 tries"
