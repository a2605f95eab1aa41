#!/bin/bash
# Runs on the GPU box (via gpurun): plain bench lines, then the ncu launch list and one full capture
# of the hot kernels (each only after the same command has exited 0 without ncu).  Outputs: gpurun_out/.
set -u
mkdir -p gpurun_out
TAG=${1:-r01}
SMALL="python bench.py --batch 32 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e"
python bench.py > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; echo "bench rc=$?"; tail -c 3000 gpurun_out/bench_${TAG}.json
python bench.py --config 0 --no-cpu-baseline > gpurun_out/bench_tau_${TAG}.json 2>> gpurun_out/bench_${TAG}.err; echo "bench tau rc=$?"
$SMALL > gpurun_out/plain_${TAG}.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_${TAG}.csv $SMALL > gpurun_out/ncu_list_${TAG}.log 2>&1
echo "ncu list rc=$?"
$SMALL > gpurun_out/plain2_${TAG}.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k 'regex:smooth_sobel|hash_tiles|match_rows' -s 9 -c 3 -f -o gpurun_out/prof_${TAG} $SMALL > gpurun_out/ncu_full_${TAG}.log 2>&1
echo "ncu full rc=$?"
$SMALL --config 0 > gpurun_out/plain3_${TAG}.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k 'regex:hash_tiles' -s 3 -c 1 -f -o gpurun_out/prof_tau_${TAG} $SMALL --config 0 > gpurun_out/ncu_full_tau_${TAG}.log 2>&1
echo "ncu full tau rc=$?"
ls -la gpurun_out | tail -12
