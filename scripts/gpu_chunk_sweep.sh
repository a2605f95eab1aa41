#!/bin/bash
# Runs on the GPU box: end-to-end throughput of gpc_match_batch against the pipeline chunk size (GPC_CHUNK_PAIRS).
mkdir -p gpurun_out; : > gpurun_out/chunk_sweep.txt
for ch in 8 16 24 32 64; do
  out=$(GPC_CHUNK_PAIRS=$ch python bench.py --no-cpu-baseline --steps 10 2>&1 | tail -1)
  echo "chunk $ch $(echo "$out" | python -c '
import json, sys
d = json.loads(sys.stdin.read()); e = d["e2e"]
print(round(d["value"]), "e2e", round(e["value"]), "ceiling", round(e.get("copy_ceiling_pairs_per_s", 0)), "frac", round(e.get("frac_of_ceiling", 0), 3))')" | tee -a gpurun_out/chunk_sweep.txt
done
