#!/usr/bin/env python3
"""BASELINE.json configs[4] (SURVEY.md 8d "C5") on one GPU: rays/probe 16..256 x world 64^2x16 ..
512^2x64, full update of every selected probe, resident inputs, CUDA-event time of the update
(kernel + commit), median of `steps` after 2 warm-ups.  Prints a markdown table.
usage: sweep_c5.py [steps]"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lum_b200
from lum_b200 import scenes

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
WORLDS = [(64, 64, 16), (128, 128, 32), (256, 256, 64), (512, 512, 64)]
RAYS = [16, 32, 64, 128, 256]
rows = []
print("| world (blocks) | probes | rays/probe | rays/update | ms/update | Grays/s | alg. GB/update | achieved GB/s |")
print("|---|---|---|---|---|---|---|---|")
for wx, wy, wz in WORLDS:
    sc = scenes.synthetic_scene(wx, wy, wz)
    for R in RAYS:
        st = lum_b200.RadianceStage((wx, wy, wz), sc["bricks"].shape[0], rays_per_probe=R)
        st.upload_world(sc["world"]); st.upload_bricks(0, sc["bricks"]); st.upload_materials(sc["materials"])
        n = st.build_requests()
        st.set_debug(lum_b200.DEBUG_PRE_BLEND)
        st.update(0)
        S, H, N = st.download_debug(n, pre_blend=False)["totals"]
        st.set_debug(0)
        ms = []
        for i in range(steps + 2):
            st.update(i)
            if i >= 2:
                ms.append(st.last_update_ms())
        t = float(np.median(ms))
        gb = (3 * S + 52 * H + 12 * N) / 1e9
        rows.append(dict(world=[wx, wy, wz], probes=n, rays_per_probe=R, ms=t, grays=n * R / t / 1e6, alg_gb=gb, gbs=gb / (t * 1e-3)))
        print(f"| {wx}x{wy}x{wz} | {n:,} | {R} | {n * R:,} | {t:.3f} | {n * R / t / 1e6:.2f} | {gb:.3f} | {gb / (t * 1e-3):.0f} |", flush=True)
        st.close()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(rows, open("gpurun_out/sweep_c5.json", "w"), indent=1)
