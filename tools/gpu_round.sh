#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench (both arms), ncu launch list + full capture of the top kernel.
# usage (under gpurun): bash tools/gpu_round.sh <tag>
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/gpu.txt 2>&1
nproc > $OUT/nproc.txt
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > $OUT/pytest_gpu.log
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 ) > $OUT/smoke.log
( timeout 900 python bench.py 2> $OUT/bench.err | tail -3 ) > $OUT/bench.json
( timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2> $OUT/bench_ref.err | tail -1 ) > $OUT/bench_ref.json
# launch list of the same command (short K/W), cold-cache serialised: compare shares
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-demo > $OUT/launches_bench.log 2>&1
# full capture of the top kernel (synthetic world) and of the demo-world kernel
timeout 900 ncu --set full --clock-control none --import-source on -k regex:radiance_update -s 3 -c 1 \
    -o $OUT/prof_synth python tools/quick_time.py synthetic-256x256x64 3 > $OUT/prof_synth.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:radiance_update -s 3 -c 1 \
    -o $OUT/prof_demo python tools/quick_time.py demo 3 > $OUT/prof_demo.log 2>&1
ls -la $OUT
cat $OUT/pytest_gpu.log $OUT/smoke.log $OUT/bench.json $OUT/bench_ref.json
