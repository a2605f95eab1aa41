#!/bin/bash
# Runs on the GPU box (via gpurun): plain bench, then ncu launch list + one full capture of the
# two hot kernels.  Outputs land in gpurun_out/.
set -u
mkdir -p gpurun_out
TAG=${1:-r01}
SMALL="python bench.py --batch 32 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e"
python bench.py > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; echo "bench rc=$?"; tail -c 2500 gpurun_out/bench_${TAG}.json
python bench.py --forest zero --no-cpu-baseline > gpurun_out/bench_zero_${TAG}.json 2>> gpurun_out/bench_${TAG}.err; echo "bench zero rc=$?"
$SMALL > gpurun_out/plain_${TAG}.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_${TAG}.csv $SMALL > gpurun_out/ncu_list_${TAG}.log 2>&1
echo "ncu list rc=$?"
$SMALL > gpurun_out/plain2_${TAG}.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k 'regex:preprocess_hash|match_rows' -s 6 -c 2 -f -o gpurun_out/prof_${TAG} $SMALL > gpurun_out/ncu_full_${TAG}.log 2>&1
echo "ncu full rc=$?"
ls -la gpurun_out | tail -20
