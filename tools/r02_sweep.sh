#!/bin/bash
# Round-2 kernel sweep: build compile-time variants of liblumcu.so side by side (here, no GPU), then on a GPU box time
# each on the C4 world (pooled kernel), the C4 world through the per-workgroup kernel and the demo world, run the
# parity tests on the ones named in $PARITY, and take one `ncu --set full` capture of the default and of the fastest.
#   here:          bash tools/r02_sweep.sh build
#   under gpurun:  bash tools/r02_sweep.sh run <tag>    -> gpurun_out/<tag>/
# Variants: "name:flags" in tools/r02_variants.txt (one per line, # comments).
set -u
cd "$(dirname "$0")/.."
mapfile -t VARIANTS < <(grep -v '^#' tools/r02_variants.txt | grep ':')
NV="nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -Xcompiler -fPIC,-ffp-contract=off,-Wall -I include -shared"
if [ "${1:-}" = build ]; then
  pids=()
  for v in "${VARIANTS[@]}"; do
    n=${v%%:*}; f=${v#*:}
    ( $NV $f -o lum_b200/exp_$n.so lum_b200/csrc/lumcu.cu 2>&1 | grep -v "minnctapersm will be ignored" ; \
      echo "built exp_$n.so ($f) $(cuobjdump -res-usage lum_b200/exp_$n.so 2>/dev/null | grep -A1 'pool_kernelILi64ELi128ELi3ELb0\|pool_kernelILi64ELi64ELi3ELb0' | grep -o 'REG:[0-9]* STACK:[0-9]*' | head -1)" ) &
    pids+=($!)
    if [ ${#pids[@]} -ge 6 ]; then wait ${pids[0]}; pids=("${pids[@]:1}"); fi
  done
  wait
  exit 0
fi
TAG=${2:-r02sweep}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/gpu.txt 2>&1
python tools/quick_time.py synthetic-256x256x64 1 > /dev/null 2>&1      # generates + caches the scene
{
for v in "${VARIANTS[@]}"; do
  n=${v%%:*}
  export LUMCU_LIB=$PWD/lum_b200/exp_$n.so
  a=$(timeout 120 python tools/quick_time.py synthetic-256x256x64 8 2>&1 | tail -1)
  b=$(LUMCU_KERNEL_VARIANT=0 timeout 120 python tools/quick_time.py demo 200 2>&1 | tail -1)
  echo "$n | $a | $b"
done
} | tee $OUT/times.log
unset LUMCU_LIB
for n in ${PARITY:-}; do
  echo "== parity $n" >> $OUT/parity.log
  LUMCU_LIB=$PWD/lum_b200/exp_$n.so timeout 300 python -m pytest tests/test_parity_gpu.py tests/test_glossy_gpu.py -m gpu -q -x -k "not full_size" 2>&1 | tail -2 >> $OUT/parity.log
done
[ -x tools/gather_ceiling ] && timeout 300 tools/gather_ceiling > $OUT/gather_ceiling.json 2> $OUT/gather_ceiling.err
# full ncu capture: the default and the fastest variant on the C4 world
best=$(sort -t'|' -k2 $OUT/times.log | python -c "
import sys,re
r=[]
for l in sys.stdin:
    m=re.search(r'^(\S+) \|.*?mean ([0-9.]+)',l)
    if m: r.append((float(m.group(2)),m.group(1)))
print(min(r)[1] if r else 'base')")
echo "fastest: $best" | tee -a $OUT/times.log
for n in base $best; do
  LUMCU_LIB=$PWD/lum_b200/exp_$n.so timeout 600 ncu --set full --clock-control none --import-source on -k regex:radiance_update -s 3 -c 1 \
      -o $OUT/prof_$n python tools/quick_time.py synthetic-256x256x64 3 > $OUT/prof_$n.log 2>&1
  [ "$best" = base ] && break
done
cat $OUT/times.log; cat $OUT/parity.log 2>/dev/null; cat $OUT/gather_ceiling.json 2>/dev/null | head -30
