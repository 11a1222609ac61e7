#!/usr/bin/env python3
"""gpurun_out/c5/c5_n{1,2,4,8}.json (tools/c5_sweep.py) -> profiles/r02_c5_sweep_n1.md and profiles/r02_c5_strong_scaling.md.
Refuses to mix commits: every table row comes from the commit named in the header.  usage: c5_tables.py (run here)"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = os.path.join(ROOT, "gpurun_out", "c5")
data = {n: json.load(open(os.path.join(src, f"c5_n{n}.json"))) for n in (1, 2, 4, 8) if os.path.exists(os.path.join(src, f"c5_n{n}.json"))}
commits = {d["commit"] for d in data.values()}
if len(commits) != 1:
    sys.exit(f"the C5 runs come from different commits: {commits}")
commit = commits.pop()
d1 = data[1]
with open(os.path.join(ROOT, "profiles", "r02_c5_sweep_n1.md"), "w") as f:
    f.write(f"C5 (BASELINE configs[4]) on one {d1['gpu']}, commit {commit}: `python tools/c5_sweep.py` ({d1['steps']} timed updates after "
            f"{d1['warmup']} warm-ups, CUDA events on the stage's stream, L2 flushed between updates; e2e = pipelined host-buffer frames).\n\n")
    f.write("| world (blocks) | probes | rays/probe | rays/update | ms/update | G rays/s | alg. GB/update | achieved GB/s | ms/update e2e |\n|---|---|---|---|---|---|---|---|---|\n")
    for r in d1["rows"]:
        f.write(f"| {r['workload'].split('-')[1]} | {r['probes']:,} | {r['rays_per_probe']} | {r['rays']:,} | {r['ms_per_update']:.3f} | "
                f"{r['rays_per_s'] / 1e9:.2f} | {r['algorithmic_bytes'] / 1e9:.3f} | {r['achieved_GBps']:.0f} | {r['e2e_ms_per_update']:.3f} |\n")
base = {r["workload"]: r for r in d1["rows"] if r["rays_per_probe"] == 64}
with open(os.path.join(ROOT, "profiles", "r02_c5_strong_scaling.md"), "w") as f:
    f.write(f"C5 strong scaling (BASELINE configs[4]) at 64 rays per probe on one 8x{d1['gpu']} box, commit {commit}: `tools/c5_sweep.py` under "
            f"torchrun at N = 2, 4, 8 (interleaved probe shards, one flag barrier + one gather/commit kernel per frame); parity = N replicas "
            f"word-for-word equal to the unsharded frames, checked in the same run.\n\n")
    f.write("| world (blocks) | probes | N | ms/update (resident) | G rays/s | parallel efficiency | ms/update (e2e) | e2e efficiency | parity |\n|---|---|---|---|---|---|---|---|---|\n")
    for wl, b in base.items():
        f.write(f"| {wl.split('-')[1]} | {b['probes']:,} | 1 | {b['ms_per_update']:.3f} | {b['rays_per_s'] / 1e9:.2f} | 100 % | {b['e2e_ms_per_update']:.3f} | 100 % | |\n")
        for n in (2, 4, 8):
            if n not in data:
                continue
            r = next(x for x in data[n]["rows"] if x["workload"] == wl)
            f.write(f"| {wl.split('-')[1]} | {r['probes']:,} | {n} | {r['ms_per_update']:.3f} | {r['rays_per_s'] / 1e9:.2f} | "
                    f"{100 * b['ms_per_update'] / (n * r['ms_per_update']):.0f} % | {r['e2e_ms_per_update']:.3f} | "
                    f"{100 * b['e2e_ms_per_update'] / (n * r['e2e_ms_per_update']):.0f} % | {'equal' if r['multi_gpu_parity'] else 'DIFFERENT'} |\n")
print("wrote profiles/r02_c5_sweep_n1.md, profiles/r02_c5_strong_scaling.md for commit", commit)
