#!/bin/bash
# the 1-bit-per-voxel occupancy mirror for the coarse march, built (-DLUMCU_OCC) and measured against
# the shipped kernel on the C4 world.  usage (under gpurun): bash tools/occ_experiment.sh <tag>
OUT=gpurun_out/${1:-occ}; mkdir -p $OUT
( LUMCU_LIB=$PWD/lum_b200/exp_occ.so timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -x -q 2>&1 | tail -2 ) > $OUT/parity_occ.log
for rep in 1 2; do for v in base occ; do
  LUMCU_OCC_ONCE=1 LUMCU_LIB=$PWD/lum_b200/exp_$v.so python tools/quick_time.py synthetic-256x256x64 8 2>&1 | tail -1 | sed "s/^/$v /"
  LUMCU_OCC_ONCE=1 LUMCU_LIB=$PWD/lum_b200/exp_$v.so python tools/quick_time.py synthetic-512x512x64 3 2>&1 | tail -1 | sed "s/^/$v /"
done; done | tee $OUT/times.log
for v in base occ; do
  LUMCU_OCC_ONCE=1 LUMCU_LIB=$PWD/lum_b200/exp_$v.so timeout 600 ncu --set full --clock-control none --import-source on -k regex:radiance_update -s 3 -c 1 \
      -o $OUT/prof_$v python tools/quick_time.py synthetic-256x256x64 3 > $OUT/prof_$v.log 2>&1
done
cat $OUT/parity_occ.log
