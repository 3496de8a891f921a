#!/bin/bash
# gpurun with retries while the pod answers "transient" (rc 3 / nothing charged).  usage: tools/gpurun_retry.sh [gpurun args] -- cmd
for i in $(seq 1 20); do
  out=$(/usr/local/graft/bin/gpurun "$@" 2>&1); rc=$?
  if echo "$out" | grep -q "status=transient"; then sleep 150; continue; fi
  echo "$out"; exit $rc
done
echo "gave up after 20 transient answers"; exit 3
