#!/usr/bin/env python3
"""A tuning build's kernels on the host SIMT emulator against the oracle (no GPU): the pooled kernel compiled with the
given -D flags runs full and ragged batches of a dense world at 16 / 64 / 256 rays per probe and of the demo world.
usage: emu_check_variant.py LUMCU_PARK=6 LUMCU_TAILDIV=64 ..."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, ROOT)
import test_kernels_on_host as T
import oracle as O
from lum_b200 import scenes
from test_device_source_on_host import HERE

defs = sys.argv[1:]
tag = "_".join(d.replace("=", "") for d in defs) or "default"
lib = T.build(os.path.join(HERE, "_build", f"libkoh_{tag}.so"), defs)
demo = scenes.demo_scene()
small = scenes.synthetic_scene(24, 20, 12, palette=48, seed=1234)
req = O.build_requests(small["world"])
for rays, n in ((16, 50), (64, 20), (64, 18), (256, 7)):          # 18 probes at 64 rays: three full batches (the permuted order)
    T.check(lib, O, small, req[::97][:n], 11, T.POOL, rays=rays)
T.check(lib, O, demo, O.build_requests(demo["world"])[1000:1061], 5, T.POOL)
T.check(lib, O, demo, O.build_requests(demo["world"])[1000:1060], 6, T.POOL, size=3, shift=1, slack=7)
print("emulator: bit-exact with", defs or "defaults")
