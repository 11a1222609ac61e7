#!/usr/bin/env python3
"""BASELINE.json configs[4] (SURVEY.md 8d "C5"), reproducible: rays per probe 16..256 x worlds 64^2x16 .. 512^2x64 at
N = 1, and the four worlds at 64 rays per probe at N = 2, 4, 8 (strong scaling), through bench.py's own Runner (same
timed regions: CUDA events on the stage's stream, L2 flushed between steps, max over ranks; e2e = the pipelined
host-buffer frames).  One JSON per rank count under gpurun_out/c5/, stamped with the commit that produced it;
tools/c5_tables.py turns them into profiles/r02_c5_*.md.

    python tools/c5_sweep.py                                   (N = 1: every world x every ray count)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/c5_sweep.py
"""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
import bench

WORLDS = ["synthetic-64x64x16", "synthetic-128x128x32", "synthetic-256x256x64", "synthetic-512x512x64"]
RAYS = [16, 32, 64, 128, 256]
STEPS, WARMUP = 8, 3


def main():
    rank, local, n = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("LOCAL_RANK", "0"), ("WORLD_SIZE", "1")))
    torch.cuda.set_device(local)
    if n > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    try:
        head = subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True, cwd=ROOT).stdout.strip()
    except Exception:
        head = ""
    head = head or os.environ.get("LUM_COMMIT", "unknown")
    rows = []
    for wl in WORLDS:
        scene = bench.make_scene(wl)
        for rays in (RAYS if n == 1 else [64]):
            r = bench.Runner(scene, local, rank, n, torch, dist, rays=rays)
            tot, _ = r.timed(r.step_resident, STEPS, WARMUP)
            ms = tot / STEPS
            e2e, _ = r.timed(r.step_e2e, STEPS, WARMUP, span=True)
            row = dict(workload=wl, n_gpus=n, rays_per_probe=rays, probes=int(r.nreq), rays=int(r.nreq * rays),
                       ms_per_update=ms, rays_per_s=r.nreq * rays / (ms * 1e-3), e2e_ms_per_update=e2e / STEPS)
            if n == 1:
                ab = float(np.mean(r.algorithmic_bytes(2, 1)))
                row.update(algorithmic_bytes=ab, achieved_GBps=ab / (ms * 1e-3) / 1e9)
            else:
                row["multi_gpu_parity"] = r.parity_check(frames=2)["equal"]
            if rank == 0:
                rows.append(row)
                print(json.dumps(row), flush=True)
            if n > 1:
                dist.barrier()
            r.close()
            if n > 1:
                dist.barrier()
    if rank == 0:
        out = os.path.join(ROOT, "gpurun_out", "c5")
        os.makedirs(out, exist_ok=True)
        json.dump(dict(commit=head, steps=STEPS, warmup=WARMUP, gpu=torch.cuda.get_device_name(0), rows=rows),
                  open(os.path.join(out, f"c5_n{n}.json"), "w"), indent=1)
    if n > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
