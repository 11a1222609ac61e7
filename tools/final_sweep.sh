#!/bin/bash
# Last sweep of round 2: tuning builds of liblumcu.so from knobs that exist in the shipped source (tools/final_variants.txt,
# "name:flags"), built here side by side as lum_b200/exp_<name>.so (git-ignored; they travel to the GPU box), then timed
# there by tools/final_sweep.py.
#   here:          bash tools/final_sweep.sh build
#   under gpurun:  python tools/final_sweep.py <tag>
set -u
cd "$(dirname "$0")/.."
NV="nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -Xcompiler -fPIC,-ffp-contract=off,-Wall -I include -shared"
mapfile -t VARIANTS < <(grep -v '^#' tools/final_variants.txt | grep ':')
pids=()
for v in "${VARIANTS[@]}"; do
  n=${v%%:*}; f=${v#*:}
  ( $NV $f -o lum_b200/exp_$n.so lum_b200/csrc/lumcu.cu 2>&1 | grep -v "minnctapersm will be ignored"
    echo "exp_$n.so ($f) $(cuobjdump -res-usage lum_b200/exp_$n.so 2>/dev/null | grep -A1 'pool_kernelILi64ELi1[0-9]*ELi[0-9]ELb0' | grep -o 'REG:[0-9]* STACK:[0-9]* SHARED:[0-9]*' | head -1)" ) &
  pids+=($!)
  if [ ${#pids[@]} -ge 7 ]; then wait ${pids[0]}; pids=("${pids[@]:1}"); fi
done
wait
