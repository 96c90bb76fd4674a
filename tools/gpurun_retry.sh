#!/bin/bash
# developer aid (build container): gpurun with retries while the pod has no free slot.  usage: tools/gpurun_retry.sh <timeout-s> '<command>'
t=$1; shift
for i in $(seq 1 40); do
  out=$(/usr/local/graft/bin/gpurun --timeout $t -- "$@" 2>&1); rc=$?
  if echo "$out" | grep -q "status=transient\|status=busy"; then sleep 45; continue; fi
  echo "$out"; exit $rc
done
echo "gpurun: gave up after retries"; exit 3
