#!/bin/bash
# Retry a gpurun call until the pod has a slot (status "transient" is not charged).  usage: scripts/gpu_retry.sh <timeout s> <command>
T=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $T -- "$@" > /tmp/gpurun_try.log 2>&1
  if ! grep -q "status=transient" /tmp/gpurun_try.log; then cat /tmp/gpurun_try.log; exit 0; fi
  sleep 150
done
echo "no GPU slot after 40 tries"; exit 3
