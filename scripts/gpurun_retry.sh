#!/bin/bash
# gpurun with retries while the pod has no free GPU slot (exit 3 / "transient"): scripts/gpurun_retry.sh [gpurun args] -- 'cmd'
for attempt in $(seq 1 40); do
  out=$(/usr/local/graft/bin/gpurun "$@" 2>&1)
  echo "$out" | tail -80
  if echo "$out" | grep -q "status=transient"; then
    echo "[retry] attempt $attempt: no slot, sleeping 90 s"
    sleep 90
    continue
  fi
  break
done
