#!/bin/bash
# Runs on the GPU box (via gpurun): plain bench lines, then the ncu launch list and one full capture of the hot kernels
# AT THE BENCHMARKED BATCH (256 pairs per step; each only after the same command has exited 0 without ncu).
# Outputs: gpurun_out/.   usage: scripts/gpu_profile.sh <tag>
set -u
mkdir -p gpurun_out
TAG=${1:-r02}
STEP="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e"
python bench.py > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; echo "bench rc=$?"; tail -c 1500 gpurun_out/bench_${TAG}.json
python bench.py --config 0 > gpurun_out/bench_tau_${TAG}.json 2>> gpurun_out/bench_${TAG}.err; echo "bench tau rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_${TAG}.json 2>> gpurun_out/bench_${TAG}.err; echo "bench reference arm rc=$?"
$STEP > gpurun_out/plain_${TAG}.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_${TAG}.csv $STEP > gpurun_out/ncu_list_${TAG}.log 2>&1
echo "ncu list rc=$?"
# the step is launched 4 times (3 warm-up + 1 timed) + once more by nothing else: skip the first 3 launches of each kernel
$STEP > gpurun_out/plain2_${TAG}.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k 'regex:smooth_sobel|hash_tiles|match_rows_fast|match_rows_tail|emit_supports' -s 15 -c 5 -f -o gpurun_out/prof_${TAG} $STEP > gpurun_out/ncu_full_${TAG}.log 2>&1
echo "ncu full rc=$?"
$STEP --config 0 > gpurun_out/plain3_${TAG}.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k 'regex:hash_tiles' -s 3 -c 1 -f -o gpurun_out/prof_tau_${TAG} $STEP --config 0 > gpurun_out/ncu_full_tau_${TAG}.log 2>&1
echo "ncu full tau rc=$?"
ls -la gpurun_out | tail -8
