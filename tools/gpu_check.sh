#!/bin/bash
# One 1-GPU box visit: GPU test-suite, smoke, quick timings, bench.  usage (under gpurun): bash tools/gpu_check.sh <tag> [bench args]
TAG=${1:-chk}; shift
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/gpu.txt 2>&1
( timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 ) > $OUT/pytest_gpu.log
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 ) > $OUT/smoke.log
( python tools/quick_time.py synthetic-256x256x64 8; LUMCU_KERNEL_VARIANT=0 python tools/quick_time.py demo 200 ) > $OUT/times.log 2>&1
( timeout 900 python bench.py "$@" 2> $OUT/bench.err | tail -3 ) > $OUT/bench.json
cat $OUT/pytest_gpu.log $OUT/smoke.log $OUT/times.log; cut -c1-3000 $OUT/bench.json; tail -5 $OUT/bench.err
