#!/bin/bash
# usage: scripts/gpu_iter.sh <tag> <kernel-regex> [pytest -k expression]   (runs on the GPU box via gpurun)
# Quick development loop: a parity subset, a bench line, the ncu launch list (+ warp instructions) and one full capture.
set -u
mkdir -p gpurun_out
TAG=$1; KREGEX=$2; KEXPR=${3:-"golden or random or kat or repeat or sort_matcher or fuzz"}
timeout 900 python -m pytest tests -m gpu -x -q -k "$KEXPR" 2>&1 | tail -6
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err || tail -5 gpurun_out/bench_${TAG}.err
python - <<PY
import json
d = json.load(open("gpurun_out/bench_${TAG}.json"))
print("value", round(d["value"]), "e2e", round(d["e2e"]["value"]), d["roofline"]["kernel_ms_per_step"])
print({k: round(v) for k, v in d.get("other_configs", {}).items()})
PY
SMALL="python bench.py --batch 32 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e"
$SMALL > gpurun_out/plain_${TAG}.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_${TAG}.log; exit 1; }
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_${TAG}.csv $SMALL > gpurun_out/ncu_list_${TAG}.log 2>&1
echo "ncu list rc=$?"
ncu --set full --clock-control none --import-source on -k "regex:${KREGEX}" -s 1 -c 2 -f -o gpurun_out/prof_${TAG} $SMALL > gpurun_out/ncu_full_${TAG}.log 2>&1
echo "ncu full rc=$?"
