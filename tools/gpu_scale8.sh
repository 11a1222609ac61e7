#!/bin/bash
# N-GPU runs of bench.py with chosen flag sets.  usage (gpurun --gpus N): bash tools/gpu_scale8.sh <tag> <N> "<flags 1>" "<flags 2>" ...
TAG=${1:-s8}; N=${2:-8}; shift; shift
OUT=gpurun_out/$TAG; mkdir -p $OUT
i=0
for flags in "$@"; do
  i=$((i+1))
  ( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 $flags 2> $OUT/run$i.err | tail -1 ) > $OUT/run$i.json
  python -c "
import json;l=json.loads(open('$OUT/run$i.json').read().strip().splitlines()[-1]);r=l['roofline']
print('$flags |', round(l['ms_per_step'],4), 'e2e', round(l['e2e']['ms_per_step'],4), 'phases', [round(x,3) for x in (l['e2e'].get('stream_ms_upload_build_update_download') or [])], 'kernel/rank', [round(x,3) for x in r['kernel_ms_per_rank']], l.get('multi_gpu_parity',{}).get('equal'))" || tail -5 $OUT/run$i.err
done
