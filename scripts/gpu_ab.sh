#!/bin/bash
# Runs on the GPU box (gpurun -- scripts/gpu_ab.sh [pytest -k expression]): A/B of the in-tree library and every
# variant library under opengpc_b200/variants/ -- a parity subset first (a variant that fails it is not timed), then
# the resident bench of both forests (kernel times only).  One line per (library, forest) in gpurun_out/ab.txt.
set -u
mkdir -p gpurun_out
KEXPR=${1:-"golden or random or kat or repeat or fuzz"}
: > gpurun_out/ab.txt
for lib in opengpc_b200/libgpc_b200.so opengpc_b200/variants/*.so; do
  [ -f "$lib" ] || continue
  name=$(basename $lib)
  if ! GPC_B200_LIB=$PWD/$lib timeout 600 python -m pytest tests -m gpu -x -q -k "$KEXPR" > gpurun_out/ab_${name}.log 2>&1; then
    echo "$name PARITY FAILED: $(tail -3 gpurun_out/ab_${name}.log | tr '\n' ' ')" | tee -a gpurun_out/ab.txt
    continue
  fi
  cfgs="1"; [ "$name" = libgpc_b200.so ] && cfgs="1 0"       # variants: the headline config only
  for cfg in $cfgs; do
    out=$(GPC_B200_LIB=$PWD/$lib python bench.py --config $cfg --no-cpu-baseline --no-e2e --steps 20 --warmup 3 2>&1 | tail -1)
    echo "$name config $cfg $(echo "$out" | python -c '
import json, sys
d = json.loads(sys.stdin.read())
print(round(d["value"]), round(d["ms_per_step"], 4), {k: round(v, 4) for k, v in d["roofline"]["kernel_ms_per_step"].items()}, d.get("verified"))
' 2>&1 | tail -1)" | tee -a gpurun_out/ab.txt
  done
done
