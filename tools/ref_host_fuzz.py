#!/usr/bin/env python3
"""The oracle's host-side functions against the REFERENCE's own C++ (oracle/_ref/libref_host.so: update_radiance's list
rule, get_shift, blockify_mesh, map_mesh's arithmetic, cut out of the tree verbatim) and its own map.comp
(libref_map_glsl.so) over MANY random inputs -- the live checks of tests/test_ref_host.py / test_ref_glsl.py beyond their
seeds.  Bar: equal (integers; floats bit for bit; bricks byte for byte).  Needs /root/reference; no GPU.
usage: ref_host_fuzz.py [first_case] [count]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle as O
from lum_b200 import scenes
from test_ref_host import PFIELDS_F, PFIELDS_I, bits, placement_dict

first = int(sys.argv[1]) if len(sys.argv) > 1 else 0
count = int(sys.argv[2]) if len(sys.argv) > 2 else 100
if not (O.ref_host_available() and O.ref_map_available()):
    raise SystemExit("oracle/_ref is not built (python __graft_entry__.py where /root/reference exists)")
demo = scenes.demo_scene()
models_vox = [demo[k] for k in ("model_tank_body", "model_tank_head", "model_robot", "model_tank") if k in demo]
t0 = time.time()
nreq = nplace = nvox = 0
for case in range(first, first + count):
    rng = np.random.default_rng(900000 + case)
    # ---- the request rule (internal_render.cpp:1142-1169) ----
    wx, wy, wz = (int(v) for v in rng.integers(1, 40, 3))
    dens = float(rng.choice([0.0, 0.003, 0.01, 0.1, 0.6, 1.0]))
    w = np.where(rng.random((wz, wy, wx)) < dens, rng.integers(-3, 40, (wz, wy, wx)), 0).astype(np.int16)
    nsp = int(rng.integers(0, 80))
    sp = np.stack([rng.integers(0, wx, nsp), rng.integers(0, wy, nsp), rng.integers(0, wz, nsp), rng.integers(0, 3, nsp)], 1).astype(np.int16)
    a, b = O.build_requests(w, sp if nsp else None), O.ref_host_build_requests(w, sp if nsp else None)
    assert np.array_equal(a, b), f"case {case}: request lists differ ({wx}x{wy}x{wz}, density {dens}, {nsp} specials)"
    nreq += len(a)
    # ---- placement (get_shift, map_mesh) + blockify ----
    models, pls = [], []
    for k in range(int(rng.integers(1, 5))):
        ax = rng.normal(size=3); ax /= np.linalg.norm(ax)
        ang = rng.uniform(-3.2, 3.2)
        rot = (np.float32(np.cos(ang / 2)), *(np.float32(np.sin(ang / 2)) * ax.astype(np.float32)))
        vox = models_vox[int(rng.integers(0, len(models_vox)))]
        size = [int(v) for v in vox.shape[::-1]] if k == 0 else [int(v) for v in rng.integers(1, 64, 3)]
        shift = [float(np.float32(v)) for v in rng.uniform(-60, 16 * 48 + 40, 3)]
        models.append((rot, shift, size))
        pa, pb = placement_dict(O.place_model(rot, shift, size, (48, 48, 16))), placement_dict(O.ref_host_place_model(rot, shift, size, (48, 48, 16)))
        for f in PFIELDS_F + ("inverse",):
            assert np.array_equal(bits(pa[f]), bits(pb[f])), f"case {case}: placement field {f} differs"
        for f in PFIELDS_I:
            assert np.array_equal(pa[f], pb[f]), f"case {case}: placement field {f} differs"
        pls.append(O.place_model(rot, shift, size, (48, 48, 16)))
        if k == 0:
            first_vox = vox
    nplace += len(models)
    cur, src, dst = O.blockify(demo["world"], 15, pls)
    rcur, rsrc, rdst, queued = O.ref_host_blockify(demo["world"], 15, models)
    assert np.array_equal(cur, rcur) and np.array_equal(src, rsrc) and np.array_equal(dst, rdst), f"case {case}: blockify differs"
    # ---- map.comp on the first model (the reference's shader text against the restatement) ----
    if len(dst) and int(dst.max()) < 1024:
        bricks = np.zeros((1024, 16, 16, 16), np.uint8)
        bricks[:15] = demo["bricks"]
        O.copy_bricks(bricks, src, dst)
        ref = bricks.copy()
        O.map_model(cur, bricks, first_vox, pls[0], 15)
        O.ref_glsl_map_model(cur, ref, first_vox, pls[0])
        assert np.array_equal(bricks, ref), f"case {case}: map.comp bricks differ"
        nvox += int(np.count_nonzero(bricks[15:]))
print(f"cases {first}..{first + count - 1}: {nreq} requests, {nplace} placements + blockify, {nvox} mapped voxels: all equal to the "
      f"reference's own C++ / map.comp ({time.time() - t0:.0f} s)")
