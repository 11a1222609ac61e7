#!/bin/bash
# usage (under gpurun): bash tools/sweep.sh name[@ENV=val[,ENV=val]] ...  -> times lum_b200/exp_<name>.so on the C4 and demo worlds
mkdir -p gpurun_out
for spec in "$@"; do
  n=${spec%%@*}; e=""; [[ "$spec" == *@* ]] && e=$(echo "${spec#*@}" | tr ',' ' ')
  echo "== $spec"
  env $e LUMCU_LIB=$PWD/lum_b200/exp_$n.so python tools/quick_time.py synthetic-256x256x64 5 2>&1 | tail -1
  env $e LUMCU_LIB=$PWD/lum_b200/exp_$n.so LUMCU_KERNEL_VARIANT=0 python tools/quick_time.py demo 50 2>&1 | tail -1
done | tee gpurun_out/sweep_$(date +%H%M%S).log
