#!/usr/bin/env python3
"""Generate tests/golden/ref_host_golden.npz from the REFERENCE's own host-side C++ of the radiance path.

Runs ONLY in the build container: oracle/_ref/libref_host.so is update_radiance's list rule, get_shift,
blockify_mesh and map_mesh's arithmetic cut out of /root/reference/src/renderer/src/internal_render.cpp and
compiled in place (oracle/Makefile, oracle/ref_host.cpp).  The vectors let the GPU box (no reference tree) and the
CPU suite check oracle.build_requests / place_model / blockify against the reference itself:

  demo_requests / demo_requests_sp   the demo world's list without and with the demo's 66 water specials
                                     (examples/demo.cpp:208-214: x in 8..29, y in 10..12, z = 1) + duplicates
  synth_requests                     a sparse random 40x24x12 world with out-of-order and repeated specials
  place_<i>_*                        24 model placements (rotations about every axis, fractional shifts, models
                                     poking out of the world): AABB, blockify border, map shift/area, inverse matrix
  blockify_<i>_*                     3 frames of demo.cpp-style scenes (2-3 overlapping models): current_world, copy pairs
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
import oracle as O  # noqa: E402
from lum_b200 import scenes  # noqa: E402


def quat(axis, a):
    ax = np.asarray(axis, np.float64); ax = ax / np.linalg.norm(ax)
    return (np.float32(np.cos(a / 2)), *(np.float32(np.sin(a / 2)) * ax.astype(np.float32)))


def placements():
    rng = np.random.default_rng(7)
    out = []
    for i in range(24):
        axis = [(0, 0, 1), (1, 0, 0), (0, 1, 0), tuple(rng.normal(size=3))][i % 4]
        rot = quat(axis, rng.uniform(-3.2, 3.2)) if i else (1.0, 0.0, 0.0, 0.0)
        size = [int(v) for v in rng.integers(1, 57, 3)]
        shift = [float(np.float32(v)) for v in rng.uniform(-40, 16 * 48 + 20, 3)]
        shift[2] = float(np.float32(rng.uniform(-20, 16 * 16 + 10)))
        out.append((rot, shift, size))
    out.append(((1.0, 0.0, 0.0, 0.0), [13.1 * 16, 14.1 * 16, 3.1 * 16], [40, 56, 24]))      # examples/demo.cpp:136
    return out


def water_specials():
    return np.array([(xx, yy, 1, 0) for xx in range(8, 30) for yy in range(10, 13)], np.int16)   # demo.cpp:208-214


def main():
    O.build()
    assert O.ref_host_available(), "oracle/_ref/libref_host.so missing (needs /root/reference)"
    d = scenes.demo_scene()
    g = {}
    g["demo_requests"] = O.ref_host_build_requests(d["world"])
    sp = np.concatenate([water_specials(), water_specials()[:5], np.array([(47, 47, 15, 0), (0, 0, 15, 0), (0, 0, 15, 0)], np.int16)])
    g["demo_specials"] = sp
    g["demo_requests_sp"] = O.ref_host_build_requests(d["world"], sp)
    rng = np.random.default_rng(3)
    w = np.where(rng.random((12, 24, 40)) < 0.02, rng.integers(1, 15, (12, 24, 40)), 0).astype(np.int16)
    ssp = np.stack([rng.integers(0, 40, 50), rng.integers(0, 24, 50), rng.integers(0, 12, 50), np.zeros(50, int)], 1).astype(np.int16)
    ssp = np.concatenate([ssp, ssp[:7]])
    g["synth_world"], g["synth_specials"] = w, ssp
    g["synth_requests"] = O.ref_host_build_requests(w, ssp)
    pl = placements()
    for i, (rot, shift, size) in enumerate(pl):
        p = O.ref_host_place_model(rot, shift, size, (48, 48, 16))
        g[f"place_{i}_in"] = np.array([*rot, *shift, *size], np.float64)
        for f in ("aabb_min", "aabb_max"):
            g[f"place_{i}_{f}"] = np.array(list(getattr(p, f)), np.float32)
        for f in ("block_min", "block_max", "map_shift", "map_area"):
            g[f"place_{i}_{f}"] = np.array(list(getattr(p, f)), np.int32)
        g[f"place_{i}_inverse"] = np.array(list(p.inverse_trans), np.float32)
    g["n_place"] = np.array(len(pl))
    for i in range(3):
        models = [pl[(5 * i + k) % len(pl)] for k in range(2 + i % 2)] + [pl[-1]]
        cur, src, dst, queued = O.ref_host_blockify(d["world"], 15, models)
        g[f"blockify_{i}_models"] = np.array([[*m[0], *m[1], *m[2]] for m in models], np.float64)
        g[f"blockify_{i}_world"], g[f"blockify_{i}_src"], g[f"blockify_{i}_dst"] = cur, src, dst
        g[f"blockify_{i}_queued"] = np.array(queued)
    g["n_blockify"] = np.array(3)
    out = os.path.join(HERE, "ref_host_golden.npz")
    np.savez_compressed(out, **g)
    print("wrote", out, os.path.getsize(out), "bytes;", len(g["demo_requests"]), "/", len(g["demo_requests_sp"]), "demo requests")


if __name__ == "__main__":
    main()
