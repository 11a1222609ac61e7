#!/bin/bash
# First GPU call of the next round: build the opt-in variants next to the default, check
# each against the parity suite and time it on the C4 and demo worlds.
#   here:          bash tools/next_round_sweep.sh build
#   under gpurun:  bash tools/next_round_sweep.sh run      -> gpurun_out/next_round_sweep.log
set -e
cd "$(dirname "$0")/.."
VARIANTS=("base:" "base10:-DLUMCU_MINB=10" "spec2wg:-DLUMCU_SPEC2_WG" "refahead:-DLUMCU_REFINE_AHEAD" "both:-DLUMCU_SPEC2_WG -DLUMCU_REFINE_AHEAD" "roll:-DLUMCU_ROLL" "roll10:-DLUMCU_ROLL -DLUMCU_MINB=10" "refahead10:-DLUMCU_REFINE_AHEAD -DLUMCU_MINB=10" "hashshare10:-DLUMCU_HASH_SHARE -DLUMCU_MINB=10" "all10:-DLUMCU_ROLL -DLUMCU_REFINE_AHEAD -DLUMCU_HASH_SHARE -DLUMCU_SPEC2_WG -DLUMCU_MINB=10" "roll_refahead:-DLUMCU_ROLL -DLUMCU_REFINE_AHEAD" "nospec2:-DLUMCU_NO_SPEC2")
if [ "$1" = build ]; then
  for v in "${VARIANTS[@]}"; do
    n=${v%%:*}; f=${v#*:}
    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false \
         -Xcompiler -fPIC,-ffp-contract=off,-Wall -I include $f -shared -o lum_b200/exp_$n.so lum_b200/csrc/lumcu.cu 2>&1 | grep -v "minnctapersm will be ignored" || true
    echo "built lum_b200/exp_$n.so ($f)"
  done
  exit 0
fi
mkdir -p gpurun_out
{
for v in "${VARIANTS[@]}"; do
  n=${v%%:*}
  echo "== $n"
  export LUMCU_LIB=$PWD/lum_b200/exp_$n.so
  timeout 120 python -m pytest tests/test_parity_gpu.py tests/test_glossy_gpu.py -m gpu -q -x -k "not full_size" 2>&1 | tail -1
  python tools/quick_time.py synthetic-256x256x64 5 2>&1 | tail -1
  LUMCU_KERNEL_VARIANT=0 python tools/quick_time.py synthetic-256x256x64 3 2>&1 | tail -1
  LUMCU_KERNEL_VARIANT=0 python tools/quick_time.py demo 50 2>&1 | tail -1
done
} | tee gpurun_out/next_round_sweep.log
