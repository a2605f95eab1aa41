#!/bin/bash
# Runs on a multi-GPU box (gpurun --gpus N): the driver's scaling launch (torchrun, one rank per GPU) next to the
# one-process pool, for N = the GPUs present; at N = 8 also BASELINE configs[2] and configs[4].
# usage: scripts/gpu_scale.sh <tag>     outputs: gpurun_out/scale_<tag>_*.json
set -u
mkdir -p gpurun_out
TAG=${1:-r02}
N=$(nvidia-smi -L | wc -l)
run() {  # label, gpus, extra bench args
  local label=$1 n=$2; shift 2
  if [ "$n" -eq 1 ]; then python bench.py --gpus 1 "$@" > gpurun_out/scale_${TAG}_${label}_n${n}.json 2> gpurun_out/scale_${TAG}_${label}_n${n}.err
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 295$((10 + n)) bench.py --gpus $n "$@" > gpurun_out/scale_${TAG}_${label}_n${n}.json 2> gpurun_out/scale_${TAG}_${label}_n${n}.err; fi
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/scale_${TAG}_${label}_n${n}.json").read().strip().splitlines()[-1])
    e = d.get("e2e", {})
    print("${label} n=${n}: value %.0f  e2e %.0f  ceiling %s  verified %s" % (d["value"], e.get("value", 0), e.get("copy_ceiling_pairs_per_s"), d.get("verified", {}).get("digest_checked_pairs")))
except Exception as ex:
    print("${label} n=${n}: FAILED", ex)
PY
}
for n in 1 2 4 8; do
  [ "$n" -le "$N" ] || continue
  run default $n --steps 10 --warmup 3 --no-cpu-baseline
  python bench.py --pool --gpus $n --steps 10 > gpurun_out/scale_${TAG}_pool_n${n}.json 2> gpurun_out/scale_${TAG}_pool_n${n}.err
  python -c "
import json; d=json.loads(open('gpurun_out/scale_${TAG}_pool_n${n}.json').read().strip().splitlines()[-1]); print('pool n=${n}: e2e %.0f' % d['value'], d['verified'])" 2>&1 | tail -1
done
if [ "$N" -ge 8 ]; then
  run config2 8 --config 2 --steps 5 --warmup 3 --no-cpu-baseline
  run config4 8 --config 4 --steps 5 --warmup 3 --no-cpu-baseline
  python bench.py --pool --gpus 8 --config 2 --steps 5 > gpurun_out/scale_${TAG}_pool_config2_n8.json 2> gpurun_out/scale_${TAG}_pool_config2_n8.err
  python -c "
import json; d=json.loads(open('gpurun_out/scale_${TAG}_pool_config2_n8.json').read().strip().splitlines()[-1]); print('pool config2 n=8: e2e %.0f' % d['value'], d['verified'])" 2>&1 | tail -1
fi
