"""CPU tests of the feeder restatements in oracle/lum_oracle_feeders.c (model placement, blockify,
brick copies, map.comp) and of the host-side API library: properties that follow from the
reference text, plus the symbol check of include/clum.h.  No GPU needed."""
import math
import os
import re
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def quat_yaw(a):
    return (np.float32(math.cos(a / 2)), 0.0, 0.0, np.float32(math.sin(a / 2)))


def test_clum_exports_every_declared_symbol():
    """liblum.so loads on a GPU-less box and exports every function include/clum.h declares."""
    from lum_b200 import clum
    lib = clum.load_library()
    hdr = open(os.path.join(ROOT, "include", "clum.h")).read()
    declared = set(re.findall(r"\b(lum_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 50
    assert declared == set(clum.SYMBOLS), declared ^ set(clum.SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), name


def test_lumcu_exports_feeder_symbols():
    import lum_b200
    lib = lum_b200.load_library()
    hdr = open(os.path.join(ROOT, "include", "lumcu.h")).read()
    declared = set(re.findall(r"\b(lumcu_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(lum_b200.SYMBOLS), declared ^ set(lum_b200.SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), name


def test_api_fails_loudly_without_gpu():
    """No CPU fallback: lum_init on a box without a CUDA device aborts with the stage's message
    (reference behaviour on a failed Vulkan call: print + exit, macros.hpp:38-50)."""
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a GPU is present")
    code = ("from lum_b200 import clum\n"
            "r = clum.Renderer()\n"
            "print('unreachable')\n")
    p = subprocess.run([sys.executable, "-c", code], cwd=ROOT, capture_output=True, text=True)
    assert p.returncode != 0
    assert "unreachable" not in p.stdout
    assert "no CPU fallback" in p.stderr and "lumcu_create" in p.stderr


def test_place_model_identity(oracle_mod):
    O = oracle_mod
    p = O.place_model((1, 0, 0, 0), (13.1 * 16, 14.1 * 16, 3.1 * 16), (18, 23, 17), (48, 48, 16))
    s = np.array([13.1 * 16, 14.1 * 16, 3.1 * 16], np.float32)
    assert np.array_equal(np.array(p.aabb_min, np.float32), s)
    assert np.array_equal(np.array(p.aabb_max, np.float32), (s.astype(np.float64) + [17, 22, 16]).astype(np.float32))
    assert list(p.map_shift) == [209, 225, 49] and list(p.map_area) == [18, 23, 17]
    # blockify border: (int(min)-1)/16 .. (int(max)+1)/16  (internal_render.cpp:1034-1035)
    assert list(p.block_min) == [13, 14, 3] and list(p.block_max) == [14, 15, 4]
    inv = np.array(p.inverse_trans, np.float32).reshape(4, 4).T
    assert np.array_equal(inv[:3, :3], np.eye(3, dtype=np.float32)) and np.array_equal(inv[:3, 3], -s)


def test_place_model_inverse_is_inverse(oracle_mod):
    O = oracle_mod
    rng = np.random.default_rng(3)
    for _ in range(20):
        q = rng.normal(size=4); q = (q / np.linalg.norm(q)).astype(np.float32)
        shift = rng.uniform(-50, 700, 3).astype(np.float32)
        p = O.place_model(q, shift, (20, 31, 12), (48, 48, 16))
        w, x, y, z = (float(v) for v in q)
        R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                      [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                      [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])
        M = np.eye(4); M[:3, :3] = R; M[:3, 3] = shift
        inv = np.array(p.inverse_trans, np.float64).reshape(4, 4).T
        assert np.allclose(inv @ M, np.eye(4), atol=2e-4)
        # the AABB contains every transformed corner and is tight
        c = np.array([[i, j, k] for i in (0, 19) for j in (0, 30) for k in (0, 11)], np.float64)
        t = c @ R.T + shift
        assert np.allclose(np.array(p.aabb_min), t.min(0), atol=1e-3) and np.allclose(np.array(p.aabb_max), t.max(0), atol=1e-3)
        assert all(0 <= a <= b <= n - 1 for a, b, n in zip(p.block_min, p.block_max, (48, 48, 16)))


def test_blockify_allocates_in_zyx_order_and_keeps_dynamic_ids(oracle_mod, demo):
    O = oracle_mod
    ws = (48, 48, 16)
    a = O.place_model((1, 0, 0, 0), (13.1 * 16, 14.1 * 16, 3.1 * 16), (18, 23, 17), ws)
    b = O.place_model(quat_yaw(0.4), (14.0 * 16, 14.5 * 16, 3.5 * 16), (12, 12, 12), ws)
    cur, src, dst = O.blockify(demo["world"], 15, [a, b])
    assert np.array_equal(dst, np.arange(15, 15 + len(dst)))               # paletteCounter++ per allocation
    changed = np.argwhere(cur != demo["world"])
    assert len(changed) == len(dst)
    order = np.lexsort((changed[:, 2], changed[:, 1], changed[:, 0]))       # z, then y, then x ascending
    first = changed[order]
    na = int(np.prod([a.block_max[k] - a.block_min[k] + 1 for k in range(3)]))
    zyx = [(int(z), int(y), int(x)) for z, y, x in first]
    ids_a = [int(cur[z, y, x]) for z, y, x in zyx if a.block_min[2] <= z <= a.block_max[2]
             and a.block_min[1] <= y <= a.block_max[1] and a.block_min[0] <= x <= a.block_max[0]]
    assert ids_a == list(range(15, 15 + na))                                # model a's box, in loop order
    assert np.array_equal(src, [demo["world"][tuple(np.argwhere(cur == d)[0])] for d in dst])
    assert (cur[cur != demo["world"]] >= 15).all()
    # a block touched by both models is allocated once
    assert len(np.unique(cur[cur >= 15])) == len(dst)


def test_map_model_identity_places_voxels_exactly(oracle_mod, demo):
    """Identity rotation at an integer shift: model voxel (i,j,k) lands on world voxel shift+(i,j,k),
    only inside dynamic blocks, only where the model is solid (map.comp:41-65)."""
    O = oracle_mod
    ws = (48, 48, 16)
    model = demo["model_tank_head"]
    sz, sy, sx = model.shape
    shift = (200.0, 230.0, 60.0)
    p = O.place_model((1, 0, 0, 0), shift, (sx, sy, sz), ws)
    cur, src, dst = O.blockify(demo["world"], 15, [p])
    bricks = np.zeros((256, 16, 16, 16), np.uint8); bricks[:15] = demo["bricks"]
    O.copy_bricks(bricks, src, dst)
    before = bricks.copy()
    O.map_model(cur, bricks, model, p, 15)
    assert np.array_equal(bricks[:15], before[:15])                         # static bricks untouched
    exp = before.copy()
    for k, j, i in np.argwhere(model != 0):
        x, y, z = 200 + i, 230 + j, 60 + k
        bid = cur[z // 16, y // 16, x // 16]
        assert bid >= 15
        exp[bid, x % 16, y % 16, z % 16] = model[k, j, i]
    assert np.array_equal(bricks, exp)


def test_map_model_only_writes_dynamic_bricks_and_model_values(oracle_mod, demo):
    O = oracle_mod
    ws = (48, 48, 16)
    model = demo["model_robot"]
    sz, sy, sx = model.shape
    for rot, shift in [(quat_yaw(0.7), (300.0, 310.0, 40.0)), (quat_yaw(-2.1), (25.0, 30.0, 30.0)),
                       ((np.float32(0.5), np.float32(0.5), np.float32(0.5), np.float32(0.5)), (740.0, 750.0, 240.0))]:
        p = O.place_model(rot, shift, (sx, sy, sz), ws)
        cur, src, dst = O.blockify(demo["world"], 15, [p])
        bricks = np.zeros((512, 16, 16, 16), np.uint8); bricks[:15] = demo["bricks"]
        O.copy_bricks(bricks, src, dst)
        before = bricks.copy()
        O.map_model(cur, bricks, model, p, 15)
        diff = bricks != before
        assert not diff[:15].any()
        assert np.isin(bricks[diff], np.unique(model[model != 0])).all()
        assert diff.any()                                                   # something was mapped
        # idempotent: mapping the same model again changes nothing
        again = bricks.copy()
        O.map_model(cur, again, model, p, 15)
        assert np.array_equal(again, bricks)
