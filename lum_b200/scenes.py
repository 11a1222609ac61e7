"""Workload definitions (inputs only, no arithmetic of the path).

demo_scene()       BASELINE.json configs 1-3: Lum's demo world, from the committed fixture
                   lum_b200/data/demo_scene.npz (generated from the reference's assets by
                   tests/golden/make_demo_fixture.py).
synthetic_scene()  BASELINE.json configs 4-5 (SURVEY.md section 8d "C4 inputs"): dense random
                   world, counter-based splitmix64 so any size is reproducible everywhere.
"""
from __future__ import annotations

import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEMO_FIXTURE = os.path.join(ROOT, "lum_b200", "data", "demo_scene.npz")
SYNTH_SEED = 0x4C554D31

_M64 = (1 << 64) - 1


def splitmix64(seed: int, start: int, n: int) -> np.ndarray:
    """Outputs start+1 .. start+n of the splitmix64 stream seeded with `seed` (counter form)."""
    idx = np.arange(start + 1, start + n + 1, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = np.uint64(seed & _M64) + idx * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return z


def _unit(u):  # top 24 bits -> [0,1)
    return (u >> np.uint64(40)).astype(np.float64) / float(1 << 24)


def demo_scene():
    """dict(world int16[16,48,48], bricks uint8[15,16,16,16], materials f32[256,6], models...)."""
    d = np.load(DEMO_FIXTURE)
    return {k: d[k] for k in d.files}


def synthetic_scene(wx=256, wy=256, wz=64, palette=1024, seed=SYNTH_SEED,
                    p_voxel=0.30, p_cell=0.25):
    """Dense random world of SURVEY.md 8(d) C4: palette bricks with each voxel solid w.p. 0.30
    (material 1..255), world cells non-empty w.p. 0.25 (brick 1..P-1), z=0 layer = brick 1,
    materials: colour U[0,1]^3, 5% emissive U[2,18], rough U[0,1].  Draw order: bricks, world
    (x fastest), materials -- all from one splitmix64 stream."""
    pos = 0
    u = splitmix64(seed, pos, palette * 4096); pos += palette * 4096
    solid = _unit(u) < p_voxel
    mat = (1 + (u & np.uint64(0xFFFFFFFF)) % np.uint64(255)).astype(np.uint8)
    bricks = np.where(solid, mat, 0).astype(np.uint8).reshape(palette, 16, 16, 16)
    bricks[0] = 0
    cells = wx * wy * wz
    u = splitmix64(seed, pos, cells); pos += cells
    full = _unit(u) < p_cell
    bid = (1 + (u & np.uint64(0xFFFFFFFF)) % np.uint64(palette - 1)).astype(np.int16)
    world = np.where(full, bid, 0).astype(np.int16).reshape(wz, wy, wx)
    world[0] = 1
    u = splitmix64(seed, pos, 256 * 8).reshape(256, 8)
    f = _unit(u).astype(np.float32)
    materials = np.zeros((256, 6), np.float32)
    materials[:, 0:3] = f[:, 0:3]
    materials[:, 3] = 1.0
    emissive = f[:, 3] < 0.05
    materials[:, 4] = np.where(emissive, 2.0 + 16.0 * f[:, 4], 0.0)
    materials[:, 5] = f[:, 5]
    materials[0] = 0.0
    return dict(world=world, bricks=bricks, materials=materials)
