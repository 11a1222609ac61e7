#!/usr/bin/env python3
"""Golden vectors produced by the REFERENCE's own shader: shaders/radiance.comp executed on the
CPU through oracle/_ref/libref_radiance_glsl.so (oracle/ref_radiance_glsl.cpp + glsl_shim.hpp;
needs /root/reference, so this runs in the build container only; the vectors travel).

Two chains, each starting from a zero field and feeding the reference's own output back in
(time_seed = frame index, as lumal.iFrame does, internal_render.cpp:1202):
  demo   Lum demo world (configs 1/2), 2,689 probes, 48 frames, size=1
  demo3  same world driven like the magicSize controller at 3: size=3, shift=frame%3, 12 frames
  synth  synthetic 24x24x8 world, 64-brick palette (the C4 generator at small size), 16 frames
For the checkpoint frames listed in *_frames the file stores the field BEFORE the frame, the field
AFTER it, the pre-blend mean radiance of every probe (f32) and, for the first two checkpoints,
the per-invocation count of world fetches (= 2 + 2*coarse iterations + refinement samples).
Consumers: tests/test_ref_glsl.py (oracle vs golden, CPU) and tests/test_parity_gpu.py (CUDA vs
golden)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
import oracle as O  # noqa: E402
from lum_b200 import scenes  # noqa: E402

assert O.ref_glsl_available(), "build oracle/_ref first (python __graft_entry__.py)"
out = {}


def chain(tag, sc, req, frames, checkpoints, size=1):
    f = np.zeros(sc.cells, np.uint32)
    out[f"{tag}_frames"] = np.array(checkpoints, np.int32)
    out[f"{tag}_size"] = np.int32(size)
    out[f"{tag}_requests"] = req.astype(np.int16)
    nfetch = 0
    for frame in range(frames):
        r = O.ref_glsl_update(sc, f, req, frame, size=size, shift=frame % size, want_rays=True)
        if frame in checkpoints:
            out[f"{tag}_before_{frame}"] = f.copy()
            out[f"{tag}_after_{frame}"] = r["field"].copy()
            out[f"{tag}_pre_{frame}"] = r["pre_blend"].copy()
            if nfetch < 2:
                assert r["ray_fetches"].max() < 65536
                out[f"{tag}_fetches_{frame}"] = r["ray_fetches"].astype(np.uint16)
                nfetch += 1
        f = r["field"]
    print(tag, "done; non-zero texels", np.count_nonzero(f & 0x3FFFFFFF))


d = scenes.demo_scene()
sc = O.Scene(d["world"], d["bricks"], d["materials"])
req = O.build_requests(d["world"])
chain("demo", sc, req, 48, [0, 1, 2, 15, 31, 47])
chain("demo3", sc, req, 12, [0, 4, 11], size=3)
s = scenes.synthetic_scene(24, 24, 8, palette=64)
sc = O.Scene(s["world"], s["bricks"], s["materials"])
req = O.build_requests(s["world"])
chain("synth", sc, req, 16, [0, 1, 15])
path = os.path.join(HERE, "ref_glsl_golden.npz")
np.savez_compressed(path, **out)
print(path, os.path.getsize(path), "bytes")
