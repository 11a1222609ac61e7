#!/usr/bin/env python3
"""Golden outputs of the oracle on the demo world (configs 1-2): 4 accumulating frames,
time_seed = 1..4, zero initial field.  Stored: per-frame totals (S,H,N), the pre-blend mean
radiance of 64 fixed probes, and the final packed field.  Regenerate only when the oracle's
pinned semantics change (and say so in DESIGN.md)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
import oracle as O  # noqa: E402
from lum_b200 import scenes  # noqa: E402

d = scenes.demo_scene()
sc = O.Scene(d["world"], d["bricks"], d["materials"])
req = O.build_requests(d["world"])
idx = np.linspace(0, len(req) - 1, 64).astype(np.int64)
f = np.zeros(sc.cells, np.uint32)
totals, pre = [], []
FRAMES = 4
for frame in range(FRAMES):
    r = O.update(sc, f, req, 1 + frame)
    f = r["field"]
    totals.append(r["totals"])
    pre.append(r["pre_blend"][idx])
np.savez_compressed(os.path.join(HERE, "oracle_demo_golden.npz"), frames=FRAMES, probe_idx=idx,
                    totals=np.array(totals, np.int64), pre_blend=np.array(pre, np.float32), field=f)
print(totals, np.count_nonzero(f & 0x3FFFFFFF))
