#!/bin/bash
# Runs on the GPU box: bench every variant library under opengpc_b200/variants/ (kernel times only).
for lib in opengpc_b200/libgpc_b200.so opengpc_b200/variants/*.so; do
  for forest in tau zero; do
    out=$(GPC_B200_LIB=$PWD/$lib python bench.py --no-cpu-baseline --no-e2e --forest $forest --steps 10 2>&1 | tail -1)
    echo "$(basename $lib) $forest $(echo "$out" | grep -o '"value": [0-9.]*' | head -1) $(echo "$out" | grep -o '"kernel_ms_per_step": {[^}]*}')"
  done
done
