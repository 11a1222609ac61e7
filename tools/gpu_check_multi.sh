#!/bin/bash
# One N-GPU box visit: the multi-GPU tests and bench.py at N ranks (fused exchange and the NCCL variant).
# usage (under gpurun --gpus N): bash tools/gpu_check_multi.sh <tag> <N> [bench args]
TAG=${1:-multi}; N=${2:-2}; shift; shift
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=index,name,clocks.sm --format=csv > $OUT/gpu.txt 2>&1
nvidia-smi topo -m > $OUT/topo.txt 2>&1
( timeout 900 python -m pytest tests/test_multigpu_gpu.py -m gpu -x -q 2>&1 | tail -25 ) > $OUT/pytest_multigpu.log
for ex in fused nccl; do
  ( timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
      bench.py --gpus $N --exchange $ex "$@" 2> $OUT/bench_$ex.err | tail -2 ) > $OUT/bench_$ex.json
done
cat $OUT/pytest_multigpu.log; for ex in fused nccl; do cut -c1-1200 $OUT/bench_$ex.json; tail -3 $OUT/bench_$ex.err; done
