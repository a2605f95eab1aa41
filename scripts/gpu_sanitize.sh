#!/bin/bash
# Runs on the GPU box (via gpurun): compute-sanitizer memcheck / racecheck / initcheck over scripts/sanitize_cases.py.
# Logs: gpurun_out/sanitize_<tool>_<case>.log  (summarised in profiles/r02_sanitizer.md)
set -u
mkdir -p gpurun_out
CS=${CS:-/usr/local/cuda/bin/compute-sanitizer}
export GPC_JIT=${GPC_JIT:-1}
run() {  # tool case timeout
  local log=gpurun_out/sanitize_$1_$2.log
  timeout $3 $CS --tool $1 --print-limit 20 --error-exitcode 3 python scripts/sanitize_cases.py $2 > $log 2>&1
  echo "$1 $2: rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $log | tail -1) | $(grep -E '^(smoke|stress|hd):' $log | tail -1)"
}
python scripts/sanitize_cases.py smoke && python scripts/sanitize_cases.py stress && python scripts/sanitize_cases.py hd || exit 1
run memcheck smoke 600
run racecheck smoke 900
run initcheck smoke 600
run memcheck stress 900
run racecheck stress 1200
run memcheck hd 900
run racecheck hd 1200
