#!/usr/bin/env python3
"""Generate lum_b200/data/demo_scene.npz (workload data of the package: bench.py and the demos load it) from the reference's demo assets.

Runs ONLY in the build container (needs /root/reference and oracle/_ref/ref_vox_dump, the
reference's own ogt_vox decoder compiled in place by oracle/Makefile).  The output is the
input data of BASELINE.json configs 1-3 in Lum's own layouts:

  world      int16 [16,48,48]      assets/scene via load_scene (load_stuff.cpp:187-219), ids clamped to [0,15]
  bricks     uint8 [15,16,16,16]   examples/demo.cpp:150-163 ids 1..14 via load_block (load_stuff.cpp:325-331); brick 0 = empty
  materials  float32 [256,6]       extract_palette_mem(tank_body.vox) (load_stuff.cpp:233-260; examples/demo.cpp:135)
  model_<name>  uint8 [sz,sy,sx]   raw .vox voxel grids (x fastest) for config 3 (tank_body, tank_head, robot, tank)

The GPU box has no /root/reference; tests and bench.py read only the committed .npz.
"""
import os
import struct
import subprocess
import sys
import tempfile

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
TOOL = os.path.join(HERE, "..", "..", "oracle", "_ref", "ref_vox_dump")

BLOCKS = ["dirt", "grass", "grassNdirt", "stone_dirt", "bush", "leaves", "iron", "lamp", "stone_brick",
          "stone_brick_cracked", "stone_pack", "bark", "wood", "planks"]  # ids 1..14
MODELS = ["tank_body", "tank_head", "tank_lf_rb_leg", "tank_rf_lb_leg", "robot", "tank"]
STATIC = 15


def dump(vox):
    with tempfile.NamedTemporaryFile(suffix=".bin") as t:
        subprocess.run([TOOL, vox, t.name], check=True)
        raw = open(t.name, "rb").read()
    sx, sy, sz = struct.unpack("<3i", raw[:12])
    n = sx * sy * sz
    vox_data = np.frombuffer(raw[12:12 + n], np.uint8).reshape(sz, sy, sx).copy()
    mats = np.frombuffer(raw[12 + n:12 + n + 256 * 24], np.float32).reshape(256, 6).copy()
    return vox_data, mats


def main():
    subprocess.run(["make", "-C", os.path.join(HERE, "..", "..", "oracle"), "ref"], check=True)
    raw = open(f"{REF}/assets/scene", "rb").read()
    sx, sy, sz = struct.unpack("<3i", raw[:12])
    assert (sx, sy, sz) == (48, 48, 16)
    world = np.frombuffer(raw[12:], np.int16).reshape(sz, sy, sx)
    world = np.clip(world, 0, STATIC).astype(np.int16)

    bricks = np.zeros((STATIC, 16, 16, 16), np.uint8)
    for i, name in enumerate(BLOCKS, start=1):
        v, _ = dump(f"{REF}/assets/{name}.vox")
        assert v.shape == (16, 16, 16), (name, v.shape)
        bricks[i] = v.transpose(2, 1, 0)  # voxels[x][y][z] = data[x + y*sx + z*sx*sy]

    out = dict(world=world, bricks=bricks)
    for name in MODELS:
        v, m = dump(f"{REF}/assets/{name}.vox")
        out[f"model_{name}"] = v
        if name == "tank_body":
            out["materials"] = m
    path = os.path.join(HERE, "..", "..", "lum_b200", "data", "demo_scene.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes")
    ids, cnt = np.unique(world, return_counts=True)
    print("block histogram", dict(zip(ids.tolist(), cnt.tolist())))
    print("emissive", [(i, float(out["materials"][i, 4])) for i in range(256) if out["materials"][i, 4] > 0])
    for k, v in out.items():
        print(k, v.shape, v.dtype)


if __name__ == "__main__":
    sys.exit(main())
