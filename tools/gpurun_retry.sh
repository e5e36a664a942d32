#!/bin/bash
# usage: tools/gpurun_retry.sh [gpurun options] -- 'command'   (retries while the pod answers "transient"/busy)
for i in $(seq 1 20); do
  out=$(gpurun "$@" 2>&1); rc=$?
  if echo "$out" | grep -q "status=transient"; then sleep 90; continue; fi
  if [ $rc -eq 3 ]; then sleep 90; continue; fi
  echo "$out"; exit $rc
done
echo "$out"; exit 3
