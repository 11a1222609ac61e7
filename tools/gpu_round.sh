#!/bin/bash
# One 1-GPU box visit at the end of a round: GPU test-suite, smoke, bench (both arms), ncu launch list of a short bench
# run, `ncu --set full` captures of the trace kernel (C4, 512x512x64, demo) and of the request builder, and
# compute-sanitizer memcheck + racecheck over the GPU tests (fuzz and shard tests included).
# usage (under gpurun): LUM_COMMIT=<hash> bash tools/gpu_round.sh <tag>
TAG=${1:-r02}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "commit ${LUM_COMMIT:-unknown}" > $OUT/commit.txt
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/gpu.txt 2>&1
nproc > $OUT/nproc.txt
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > $OUT/pytest_gpu.log
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 ) > $OUT/smoke.log
( timeout 900 python bench.py 2> $OUT/bench.err | tail -3 ) > $OUT/bench.json
( timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2> $OUT/bench_ref.err | tail -1 ) > $OUT/bench_ref.json
# launch list of the same command (short K/W), cold-cache serialised: compare shares
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-demo > $OUT/launches_bench.log 2>&1
# full captures of the trace kernel and of the request builder
timeout 900 ncu --set full --clock-control none --import-source on -k regex:radiance_update -s 3 -c 1 \
    -o $OUT/prof_synth python tools/quick_time.py synthetic-256x256x64 3 > $OUT/prof_synth.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:radiance_update -s 2 -c 1 \
    -o $OUT/prof_synth512 python tools/quick_time.py synthetic-512x512x64 2 > $OUT/prof_synth512.log 2>&1
LUMCU_KERNEL_VARIANT=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:radiance_update -s 3 -c 1 \
    -o $OUT/prof_demo python tools/quick_time.py demo 3 > $OUT/prof_demo.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:build_fused -c 1 \
    -o $OUT/prof_build python tools/quick_time.py synthetic-256x256x64 1 > $OUT/prof_build.log 2>&1
# sanitizers over the GPU tests of the shipped build (the full-size and 2-GPU tests excepted: hours under instrumentation)
SEL='not full_size and not two_gpu'
( echo "# commit ${LUM_COMMIT:-unknown}: compute-sanitizer --tool memcheck python -m pytest tests -m gpu -q -k '$SEL'"
  timeout 1500 compute-sanitizer --tool memcheck python -m pytest tests -m gpu -q -x -k "$SEL" 2>&1 | tail -8; echo "rc=$?" ) > $OUT/memcheck.txt
( echo "# commit ${LUM_COMMIT:-unknown}: compute-sanitizer --tool racecheck python -m pytest tests -m gpu -q -k '$SEL'"
  timeout 2400 compute-sanitizer --tool racecheck python -m pytest tests -m gpu -q -x -k "$SEL" 2>&1 | tail -8; echo "rc=$?" ) > $OUT/racecheck.txt
ls -la $OUT
cat $OUT/pytest_gpu.log $OUT/smoke.log; cut -c1-600 $OUT/bench.json; cut -c1-400 $OUT/bench_ref.json; cat $OUT/memcheck.txt $OUT/racecheck.txt
