#!/bin/bash
# Runs on the GPU box: global-mode throughput and launch list for the default library and every variant.
mkdir -p gpurun_out
for lib in opengpc_b200/libgpc_b200.so opengpc_b200/variants/*.so; do
  n=$(basename $lib .so)
  echo "== $n: $(GPC_B200_LIB=$PWD/$lib timeout 200 python scripts/micro/global_mode_timing.py 2>&1 | tail -1)"
  GPC_B200_LIB=$PWD/$lib ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/global_launches_$n.csv python scripts/micro/global_mode_timing.py > /dev/null 2>&1
done
