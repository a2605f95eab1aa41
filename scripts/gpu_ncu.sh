#!/bin/bash
# usage: scripts/gpu_ncu.sh <tag> <kernel-regex> [extra bench args]   (runs on the GPU box via gpurun)
set -u
mkdir -p gpurun_out
TAG=$1; KREGEX=$2; shift 2
SMALL="python bench.py --batch 32 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e $*"
$SMALL > gpurun_out/plain_${TAG}.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_${TAG}.log; exit 1; }
ncu --set full --clock-control none --import-source on -k "regex:${KREGEX}" -s 3 -c 1 -f -o gpurun_out/prof_${TAG} $SMALL > gpurun_out/ncu_full_${TAG}.log 2>&1
echo "ncu full rc=$?"
