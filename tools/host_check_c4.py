#!/usr/bin/env python3
"""The whole C4 update (4,191,908 probes x 64 rays of the synthetic 256x256x64 world, random field, seed 7) with the
kernels' per-thread source compiled for the host (tests/device_host) against the oracle, bit for bit -- every march
variant.  No GPU needed; about a minute per variant on 8 cores.  Test infrastructure, not a product path.
usage: host_check_c4.py [fast fused single general]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import oracle as O
import test_device_source_on_host as T
from lum_b200 import scenes

modes = sys.argv[1:] or ["fast", "fused", "single"]
lib = T.build_host_lib(T.OUT)
syn = scenes.synthetic_scene(256, 256, 64)
sc = T.scene_of(O, syn)
req = O.build_requests(syn["world"])
f0 = np.random.default_rng(4).integers(0, 2 ** 32, sc.cells, dtype=np.uint64).astype(np.uint32)
t = time.time()
ref = O.update(sc, f0, req, 7, want_rays=True)
print(f"{len(req)} probes; oracle {time.time() - t:.1f} s, S,H,N = {ref['totals']}, "
      f"{int(np.isnan(ref['pre_blend']).any(axis=1).sum())} NaN probe(s)")
for m in modes:
    t = time.time()
    T.assert_same(T.run(lib, O, syn, f0, req, 7, T.MODES[m]), ref)
    print(f"{m}: bit-exact on the whole update ({time.time() - t:.1f} s)")
