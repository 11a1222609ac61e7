#!/usr/bin/env python3
"""Kernel-only timing of lumcu_update on a workload (tuning helper; env knobs LUMCU_* apply).
usage: quick_time.py [workload] [steps]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lum_b200
from lum_b200 import scenes

wl = sys.argv[1] if len(sys.argv) > 1 else "synthetic-256x256x64"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
cache = f"/tmp/scene_{wl}.npz"
if wl == "demo":
    sc = scenes.demo_scene()
elif os.path.exists(cache):
    sc = dict(np.load(cache))
else:
    wx, wy, wz = (int(v) for v in wl.split("-")[1].split("x"))
    sc = scenes.synthetic_scene(wx, wy, wz)
    np.savez(cache, **sc)
wz, wy, wx = sc["world"].shape
st = lum_b200.RadianceStage((wx, wy, wz), sc["bricks"].shape[0])
st.upload_world(sc["world"]); st.upload_bricks(0, sc["bricks"]); st.upload_materials(sc["materials"])
n = st.build_requests()
ms = []
for i in range(steps + 2):
    st.update(i)
    if i >= 2:
        ms.append(st.last_update_ms())
tag = " ".join(f"{k}={v}" for k, v in os.environ.items() if k.startswith("LUMCU_"))
print(f"{wl} [{tag}] probes={n} ms/update: mean {np.mean(ms):.3f} min {np.min(ms):.3f}  -> {n*64/np.mean(ms)/1e6:.2f} Grays/s")
