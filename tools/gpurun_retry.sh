#!/bin/bash
# usage: tools/gpurun_retry.sh <outfile> <gpurun args...>   -- retries while the pod answers busy (exit 3 / transient)
out=$1; shift
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun "$@" > "$out" 2>&1
  rc=$?
  if grep -q "status=transient" "$out" || [ $rc -eq 3 ]; then sleep 90; continue; fi
  break
done
echo "gpurun_retry done rc=$rc" >> "$out"
