#!/usr/bin/env python3
"""The device source compiled for the host (tests/device_host) against the oracle on MANY random small worlds -- the
fuzz of tests/test_device_source_on_host.py::test_random_small_worlds beyond the 24 cases the suite runs (general, fast,
fused and single-step march; odd world sizes, out-of-world probes, magicSize dispatches, random fields and seeds).  No GPU.
--glossy / --map: the same for glossy_fragment() and map_model_invocation(), the bodies of the glossy and map kernels.
usage: host_fuzz.py [first_case] [count] [--glossy | --map | --rays]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, ROOT)
import test_device_source_on_host as T
import oracle as O

import ctypes as C
import numpy as np
from lum_b200 import scenes

argv = [a for a in sys.argv[1:] if not a.startswith("--")]
first = int(argv[0]) if len(argv) > 0 else 24
count = int(argv[1]) if len(argv) > 1 else 100
lib = T.build_host_lib(T.OUT)
t0 = time.time()
if "--glossy" in sys.argv:
    # glossy_fragment() -- the body of glossy_kernel -- on random worlds, cameras, fields and G-buffers (origins inside solids and
    # outside the world, arbitrary normals, roughness either side of the stencil rule; every fourth world loads a brick 0: the
    # general march) against oracle.glossy: coverage equal, colour bit for bit
    frags = 0
    for case in range(first, first + count):
        rng = np.random.default_rng(810000 + case)
        wx, wy, wz = (int(v) for v in rng.integers(1, [11, 10, 6]))
        P = int(rng.integers(2, 41))
        bricks = ((rng.random((P, 16, 16, 16)) < rng.choice([0.002, 0.05, 0.3, 0.9])) * rng.integers(1, 256, (P, 16, 16, 16))).astype(np.uint8)
        if case % 4:
            bricks[0] = 0
        world = ((rng.random((wz, wy, wx)) < rng.choice([0.1, 0.5, 1.0])) * rng.integers(1, P, (wz, wy, wx))).astype(np.int16)
        mats = rng.random((256, 6)).astype(np.float32)
        mats[:, 4] *= rng.choice([0.0, 1.0, 18.0])
        scn = dict(world=world, bricks=bricks, materials=mats)
        sc = O.Scene(world, bricks, mats)
        field = rng.integers(0, 2 ** 32, sc.cells, dtype=np.uint64).astype(np.uint32)
        h, w = int(rng.integers(1, 24)), int(rng.integers(1, 32))
        mn = rng.integers(0, 256, (h, w, 4), dtype=np.uint8)
        dp = rng.random((h, w), dtype=np.float32) * np.float32(rng.choice([0.05, 0.5]))
        cd = rng.normal(size=3); cd[2] = -abs(cd[2]) - 0.05
        if abs(cd[0]) + abs(cd[1]) < 1e-3:
            cd[0] = 0.3
        view = O.camera_view(rng.random(3) * [16 * wx + 80, 16 * wy + 80, 200] - 40, cd, w, h, pixels_in_voxel=float(rng.choice([0.2, 1.0, 5.0])))
        ref = O.glossy(sc, field, view, mn, dp)
        got, cov = T.run_glossy(lib, scn, field, view, mn, dp)
        assert np.array_equal(cov, ref["covered"]), case
        T.assert_frames_equal(got, ref["frame"])
        frags += h * w
    print(f"glossy cases {first}..{first + count - 1}: {frags} fragments, coverage equal and colour bit-exact ({time.time() - t0:.0f} s)")
elif "--map" in sys.argv:
    # map_model_invocation() -- the body of map_model_kernel -- for random placements of the demo's models against oracle.map_model
    demo = scenes.demo_scene()
    lib.device_on_host_map_model.restype = C.c_int
    p_ = lambda a: a.ctypes.data_as(C.c_void_p)
    names = [k for k in ("model_tank_body", "model_tank_head", "model_robot", "model_tank") if k in demo]
    vox = 0
    for case in range(first, first + count):
        rng = np.random.default_rng(820000 + case)
        model = np.ascontiguousarray(demo[names[int(rng.integers(0, len(names)))]], np.uint8)
        sz, sy, sx = model.shape
        ax = rng.normal(size=3); ax /= np.linalg.norm(ax)
        ang = rng.uniform(-3.2, 3.2)
        rot = (np.float32(np.cos(ang / 2)), *(np.float32(np.sin(ang / 2)) * ax.astype(np.float32)))
        shift = [float(np.float32(v)) for v in rng.uniform(-60, 16 * 48 + 40, 3)]
        shift[2] = float(np.float32(rng.uniform(-40, 16 * 16 + 30)))
        pl = O.place_model(rot, shift, (sx, sy, sz), (48, 48, 16))
        cur, src, dst = O.blockify(demo["world"], 15, [pl])
        bricks = np.zeros((1024, 16, 16, 16), np.uint8); bricks[:15] = demo["bricks"]
        O.copy_bricks(bricks, src, dst)
        got = bricks.copy()
        O.map_model(cur, bricks, model, pl, 15)
        cur = np.ascontiguousarray(cur, np.int16)
        inv = np.array(pl.inverse_trans, np.float32)
        sh = np.array(list(pl.map_shift), np.int32)
        ar = np.array(list(pl.map_area), np.int32)
        assert lib.device_on_host_map_model(C.c_int(48), C.c_int(48), C.c_int(16), C.c_int(1024), p_(cur), p_(got), p_(model), C.c_int(sx),
                                            C.c_int(sy), C.c_int(sz), p_(inv), p_(sh), p_(ar), C.c_int(15)) == 0
        assert np.array_equal(got, bricks), case
        vox += int(np.count_nonzero(bricks[15:]))
    print(f"map cases {first}..{first + count - 1}: {vox} mapped voxels, bricks byte-equal ({time.time() - t0:.0f} s)")
elif "--rays" in sys.argv:
    # the other ray counts of BASELINE configs[4] (the reference only defines 64): random worlds at 16 / 32 / 128 / 256 rays per
    # probe, random reactiveness, every march mode
    for case in range(first, first + count):
        rng = np.random.default_rng(830000 + case)
        rays = int(rng.choice([16, 32, 128, 256]))
        wx, wy, wz = (int(v) for v in rng.integers(1, [9, 8, 5]))
        P = int(rng.integers(2, 41))
        bricks = ((rng.random((P, 16, 16, 16)) < rng.choice([0.002, 0.05, 0.3, 0.9])) * rng.integers(1, 256, (P, 16, 16, 16))).astype(np.uint8)
        if case % 3:
            bricks[0] = 0
        world = ((rng.random((wz, wy, wx)) < rng.choice([0.1, 0.5, 1.0])) * rng.integers(1, P, (wz, wy, wx))).astype(np.int16)
        mats = rng.random((256, 6)).astype(np.float32)
        mats[:, 4] *= rng.choice([0.0, 1.0, 18.0])
        scn = dict(world=world, bricks=bricks, materials=mats)
        react = float(np.float32(rng.choice([0.01, 0.25, 1.0])))
        sc = T.scene_of(O, scn, rays=rays, reactiveness=react)
        inside = np.array([(x, y, z, 0) for z in range(wz) for y in range(wy) for x in range(wx)], np.int16)
        outside = np.array([(-1, 0, 0, 0), (wx, 0, 0, 0), (0, -2, wz, 0), (wx + 3, wy + 1, -1, 0)], np.int16)
        req = np.concatenate([inside, outside])
        req = req[rng.permutation(len(req))][: max(1, min(24, int(rng.integers(1, len(req) + 1))))]
        f0 = rng.integers(0, 2 ** 32, sc.cells, dtype=np.uint64).astype(np.uint32)
        size = min(int(rng.integers(1, 4)), len(req))
        shift = int(rng.integers(0, size))
        seed = int(rng.integers(-2 ** 31, 2 ** 31))
        ref = O.update(sc, f0, req, seed, size=size, shift=shift, want_rays=True)
        for mode in (("general", "fast", "fused", "single") if not bricks[0].any() else ("general",)):
            T.assert_same(T.run(lib, O, scn, f0, req, seed, T.MODES[mode], rays=rays, size=size, shift=shift, reactiveness=react), ref)
    print(f"ray-count cases {first}..{first + count - 1}: bit-exact in every march mode at 16 / 32 / 128 / 256 rays per probe ({time.time() - t0:.0f} s)")
else:
    for case in range(first, first + count):
        T.test_random_small_worlds(lib, O, case)
    print(f"cases {first}..{first + count - 1}: bit-exact in every march mode ({time.time() - t0:.0f} s)")
