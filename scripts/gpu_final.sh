#!/bin/bash
# Runs on the GPU box (gpurun -- bash scripts/gpu_final.sh <tag>): the round's closing evidence in one call --
# the whole GPU suite, the bench lines of every config (plain runs), then the ncu launch list and captures.
set -u
mkdir -p gpurun_out
TAG=${1:-r02}
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/gputests_${TAG}.txt
for cfg in 2 3 4; do
  python bench.py --config $cfg > gpurun_out/bench_config${cfg}_${TAG}.json 2> gpurun_out/bench_config${cfg}_${TAG}.err; echo "config $cfg rc=$?"
done
python bench.py --config 4 --extended > gpurun_out/bench_config4_extended_${TAG}.json 2> gpurun_out/bench_config4_extended_${TAG}.err; echo "extended rc=$?"
bash scripts/gpu_profile.sh ${TAG}
