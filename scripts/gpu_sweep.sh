#!/bin/bash
# Runs on the GPU box: bench (kernel times only) every library variant under opengpc_b200/variants/ for every
# "nib_log2 slot_log2" setting of the fast row matcher given as arguments (default: the built-in sizing).
mkdir -p gpurun_out
[ $# -eq 0 ] && set -- "0 0"
for lib in opengpc_b200/libgpc_b200.so opengpc_b200/variants/*.so; do
  [ -f "$lib" ] || continue
  for cfg in "$@"; do
    a=${cfg% *}; b=${cfg#* }
    out=$(GPC_B_NIB_LOG2=$a GPC_B_SLOT_LOG2=$b GPC_B200_LIB=$PWD/$lib python bench.py --no-cpu-baseline --no-e2e --steps 10 2>&1 | tail -1)
    echo "$(basename $lib) nib/slot=$cfg $(echo "$out" | grep -o '"value": [0-9.]*' | head -1) $(echo "$out" | grep -o '"kernel_ms_per_step": {[^}]*}')"
  done
done | tee gpurun_out/sweep.txt
