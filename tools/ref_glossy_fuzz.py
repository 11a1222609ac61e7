#!/usr/bin/env python3
"""oracle.glossy against the REFERENCE's own shaders/glossy.frag (oracle/_ref/libref_glossy_glsl.so, the shader text
compiled in place) on MANY random worlds, cameras and G-buffers -- tests/test_glossy.py's live check beyond its three
seeds: random small worlds, random fields, G-buffers no rasteriser would produce (origins inside solids and outside the
world, arbitrary normals, so arbitrary mirror directions), roughness either side of the stencil rule.
Bar: coverage EQUAL, texel fetches per fragment EQUAL, colour within 1e-6 * max(1, |x|).  Needs /root/reference; no GPU.
usage: ref_glossy_fuzz.py [first_case] [count] [--odd]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle as O

REL_TOL = 1e-6
ODD = "--odd" in sys.argv
argv = [a for a in sys.argv[1:] if not a.startswith("--")]
first = int(argv[0]) if len(argv) > 0 else 0
count = int(argv[1]) if len(argv) > 1 else 100
if not O.ref_glossy_available():
    raise SystemExit("oracle/_ref is not built (python __graft_entry__.py where /root/reference exists)")
t0 = time.time()
worst, frags, covered = 0.0, 0, 0
for case in range(first, first + count):
    rng = np.random.default_rng(800000 + case)
    wx, wy, wz = (int(v) for v in rng.integers(1, [11, 10, 6]))
    P = int(rng.integers(2, 41))
    bricks = ((rng.random((P, 16, 16, 16)) < rng.choice([0.002, 0.05, 0.3, 0.9])) * rng.integers(1, 256, (P, 16, 16, 16))).astype(np.uint8)
    if not (ODD and case % 2):
        bricks[0] = 0                                       # (--odd: every other case loads a block 0: the general march)
    world = ((rng.random((wz, wy, wx)) < rng.choice([0.1, 0.5, 1.0])) * rng.integers(1, P, (wz, wy, wx))).astype(np.int16)
    mats = rng.random((256, 6)).astype(np.float32)
    mats[:, 4] *= rng.choice([0.0, 1.0, 18.0])
    sc = O.Scene(world, bricks, mats)
    field = rng.integers(0, 2 ** 32, sc.cells, dtype=np.uint64).astype(np.uint32) if case % 4 else np.zeros(sc.cells, np.uint32)
    h, w = int(rng.integers(1, 24)), int(rng.integers(1, 32))
    mn = rng.integers(0, 256, (h, w, 4), dtype=np.uint8)
    dp = rng.random((h, w), dtype=np.float32) * np.float32(rng.choice([0.05, 0.5]))
    cd = rng.normal(size=3); cd[2] = -abs(cd[2]) - 0.05
    if abs(cd[0]) + abs(cd[1]) < 1e-3:
        cd[0] = 0.3
    view = O.camera_view(rng.random(3) * [16 * wx + 80, 16 * wy + 80, 200] - 40, cd, w, h, pixels_in_voxel=float(rng.choice([0.2, 1.0, 5.0])))
    a = O.glossy(sc, field, view, mn, dp)
    b = O.ref_glsl_glossy(sc, field, view, mn, dp)
    where = f"case {case}: world {wx}x{wy}x{wz}, frame {w}x{h}"
    assert np.array_equal(a["covered"], b["covered"]), "coverage differs, " + where
    cov = a["covered"] > 0
    assert np.array_equal(np.where(cov, 2 + a["coarse"].astype(np.int64) + a["samples"], 0), b["fetches"]), "fetch counts differ, " + where
    fa, fb = a["frame"], b["frame"]
    assert np.array_equal(np.isfinite(fa), np.isfinite(fb)), "NaN pattern differs, " + where
    fin = np.isfinite(fb)
    if fin.any():
        worst = max(worst, float((np.abs(fa - fb) / np.maximum(1.0, np.abs(fb)))[fin].max()))
    assert worst <= REL_TOL, f"colour off by {worst:g}, " + where
    frags += h * w; covered += int(cov.sum())
print(f"cases {first}..{first + count - 1}: {frags} fragments ({covered} covered = mirror rays of the reference's glossy.frag): coverage and "
      f"fetch counts equal, colour within {worst:.2e} (bar {REL_TOL:g}) ({time.time() - t0:.0f} s)")
