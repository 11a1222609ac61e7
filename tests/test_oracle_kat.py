"""CPU tests: the oracle against known-answer values derived from the shader text
 and against the committed golden vectors of tests/golden/."""
import os

import numpy as np
import pytest

from conftest import oracle_scene

GOLD = os.path.join(os.path.dirname(__file__), "golden", "oracle_demo_golden.npz")


def test_hash33_kats(oracle_mod):
    O = oracle_mod
    kats = {
        (1, 0, 0): (0x807DBCB5, 0x32EF6E3C, 0x5765CCB7),
        (2, 0, 0): (0x00FB796A, 0xD2DEDC78, 0xAECB996E),
        (63, 0, 0): (0x9EF1708B, 0x1C131967, 0x7DB3EEBD),
        (64, 0, 0): (0x1F6F2D40, 0x28149501, 0xBCDAEEAA),
        (1, 1, 1): (0x84FA8D3B,) * 3,
        (2, 1, 1): (0x637D05D4, 0xC95A4424, 0x098AEB55),
        (12352, 7, 7): (0xBE648074, 0xD33071D9, 0x7E556973),
        (0, 0, 0): (0, 0, 0),
    }
    for k, v in kats.items():
        u, f = O.hash33(*k)
        assert u == v
        # scale is exactly 2^-32 on the RNE float of the integer (radiance.comp:58)
        assert f == tuple(float(np.float32(np.float32(x) * np.float32(2.0 ** -32))) for x in v)


def test_hash33_matches_numpy_restatement(oracle_mod):
    rng = np.random.default_rng(7)
    for _ in range(200):
        x = [int(v) for v in rng.integers(0, 2 ** 32, 3)]
        a = np.array(x, dtype=np.uint64)
        for _r in range(3):
            a = (((a >> np.uint64(8)) ^ a[[1, 2, 0]]) * np.uint64(1103515245)) & np.uint64(0xFFFFFFFF)
        assert oracle_mod.hash33(*x)[0] == tuple(int(v) for v in a)


def test_direction_table(oracle_mod):
    t = oracle_mod.direction_table(64)
    spots = {0: (0.17608480, 0.0, 0.984375), 1: (-0.22311068, -0.20438774, 0.953125),
             31: (0.54077083, 0.84102482, 0.015625), 63: (0.16210027, -0.06877034, -0.984375)}
    for i, v in spots.items():
        assert t[i, 2] == np.float32(v[2])                       # z exact
        assert np.allclose(t[i, :2], v[:2], rtol=0, atol=2e-7)   # x,y within 1 ulp
    n = np.linalg.norm(t.astype(np.float64), axis=1)
    assert np.allclose(n, 1.0, atol=1e-6)
    # cosTheta = 1 - (2i+1)/64 is exact in f32
    assert np.array_equal(t[:, 2], (1.0 - (2.0 * np.arange(64) + 1.0) / 64.0).astype(np.float32))


def test_sun_dir(oracle_mod):
    s = oracle_mod.sun_dir()
    assert np.allclose(s, (0.43685203, 0.43685203, -0.78633365), atol=1e-7)


def test_pack_unpack(oracle_mod):
    O = oracle_mod
    assert O.pack((0, 0, 0)) == 3 << 30
    assert O.pack((1, 1, 1)) == 0xFFFFFFFF
    assert O.pack((2.0, -1.0, float("nan"))) == (3 << 30) | 1023
    assert O.pack((0.5, 0.25, 0.75)) == (3 << 30) | 512 | (256 << 10) | (767 << 20)  # 511.5->512 (even), 255.75->256, 767.25->767
    for w in (0, 0xFFFFFFFF, 0x12345678, 3 << 30):
        rgb = O.unpack(w)
        assert O.pack(rgb) == (w & 0x3FFFFFFF) | (3 << 30)


def test_demo_request_list(oracle_mod, demo):
    req = oracle_mod.build_requests(demo["world"])
    assert len(req) == 2689
    lin = req[:, 0].astype(np.int64) + 48 * (req[:, 1] + 48 * req[:, 2].astype(np.int64))
    assert np.all(np.diff(lin) > 0)               # z,y,x order
    # the demo's per-frame water specials (examples/demo.cpp:208-214): 140 pushed, 66 not yet selected
    sp = [(x, y, z, 0) for x in range(5, 12) for y in range(6, 16) for z in (2, 1)]
    req2 = oracle_mod.build_requests(demo["world"], np.array(sp, np.int16))
    assert len(req2) == 2689 + 66
    assert np.array_equal(req2[:2689], req)


def test_request_rule_vs_numpy(oracle_mod, small_synth):
    w = small_synth["world"].astype(np.int64)
    p = np.pad(w, 1, mode="edge")
    s = sum(p[1 + dz:p.shape[0] - 1 + dz, 1 + dy:p.shape[1] - 1 + dy, 1 + dx:p.shape[2] - 1 + dx]
            for dz in (-1, 0, 1) for dy in (-1, 0, 1) for dx in (-1, 0, 1))
    z, y, x = np.nonzero(s > 0)
    req = oracle_mod.build_requests(small_synth["world"])
    assert np.array_equal(req[:, :3], np.stack([x, y, z], 1))


def test_get_voxel_semantics(oracle_mod, demo):
    O = oracle_mod
    sc = oracle_scene(O, demo)
    w, b = demo["world"], demo["bricks"]
    rng = np.random.default_rng(3)
    for _ in range(500):
        x, y, z = (int(v) for v in rng.integers(0, (48 * 16, 48 * 16, 16 * 16)))
        bid = int(w[z // 16, y // 16, x // 16])
        assert O.get_voxel(sc, x, y, z) == (int(b[bid, x % 16, y % 16, z % 16]), bid)
    # out of range block -> id 0 -> brick 0 (empty); negative coords: trunc-div, floor-mod
    assert O.get_voxel(sc, 48 * 16, 0, 0) == (0, 0)
    assert O.get_voxel(sc, 0, 0, 16 * 16 + 5) == (0, 0)
    bid = int(w[0, 0, 0])
    assert O.get_voxel(sc, -5, 0, 0) == (int(b[bid, 11, 0, 0]), bid)     # -5/16 = 0, -5 mod 16 = 11
    assert O.get_voxel(sc, -16, 0, 0) == (0, 0)                          # block -1 -> out of range


def test_gather_is_trilinear_repeat(oracle_mod, demo):
    O = oracle_mod
    sc = oracle_scene(O, demo)
    rng = np.random.default_rng(5)
    field = rng.integers(0, 2 ** 32, sc.cells, dtype=np.uint64).astype(np.uint32)
    f3 = field.reshape(16, 48, 48)
    # integer coordinates return the texel itself
    for (x, y, z) in [(0, 0, 0), (47, 47, 15), (5, 7, 3)]:
        assert np.array_equal(O.gather(sc, field, (x, y, z)), O.unpack(f3[z, y, x]))
    # REPEAT: -1 wraps to W-1, W wraps to 0
    assert np.array_equal(O.gather(sc, field, (-1.0, 0.0, 0.0)), O.unpack(f3[0, 0, 47]))
    assert np.array_equal(O.gather(sc, field, (48.0, 48.0, 16.0)), O.unpack(f3[0, 0, 0]))
    # midpoint along x between texel 47 and texel 0 (wrap)
    a, b = O.unpack(f3[2, 3, 47]).astype(np.float32), O.unpack(f3[2, 3, 0]).astype(np.float32)
    exp = a + np.float32(0.5) * (b - a)
    assert np.array_equal(O.gather(sc, field, (47.5, 3.0, 2.0)), exp)


def test_shift_field(oracle_mod):
    dims = (6, 5, 4)
    f = np.arange(6 * 5 * 4, dtype=np.uint32)
    out = oracle_mod.shift_field(f, dims, 1, -2, 0).reshape(4, 5, 6)
    f3 = f.reshape(4, 5, 6)
    exp = f3.copy()
    exp[:, 2:, :5] = f3[:, :3, 1:]
    assert np.array_equal(out, exp)
    assert np.array_equal(oracle_mod.shift_field(f, dims, 6, 0, 0), f)   # zero-volume overlap: no-op


def test_jacobi_and_untouched_texels(oracle_mod, demo):
    O = oracle_mod
    sc = oracle_scene(O, demo)
    rng = np.random.default_rng(11)
    f0 = rng.integers(0, 2 ** 32, sc.cells, dtype=np.uint64).astype(np.uint32)
    req = O.build_requests(demo["world"])[100:164]
    r = O.update(sc, f0, req, 3)
    lin = req[:, 0].astype(np.int64) + 48 * (req[:, 1] + 48 * req[:, 2].astype(np.int64))
    mask = np.ones(sc.cells, bool); mask[lin] = False
    assert np.array_equal(r["field"][mask], f0[mask])
    assert np.all((r["field"][lin] >> 30) == 3)
    # order independence (Jacobi): reversing the list permutes pre_blend only through the RNG seeds,
    # so running each request alone at its own workgroup index must reproduce the joint result
    # for workgroup 0 exactly.
    r0 = O.update(sc, f0, req[:1], 3)
    assert np.array_equal(r0["pre_blend"][0], r["pre_blend"][0])
    assert r0["field"][lin[0]] == r["field"][lin[0]]


def test_size_shift_dispatch(oracle_mod, demo):
    O = oracle_mod
    sc = oracle_scene(O, demo)
    f0 = np.zeros(sc.cells, np.uint32)
    req = O.build_requests(demo["world"])
    r = O.update(sc, f0, req, 5, size=3, shift=2)
    assert r["totals"][2] == len(req) // 3
    # workgroup w handles request 3w+2 with gid = w*64+i: equals a size-1 run over the strided list
    r1 = O.update(sc, f0, req[2::3][:len(req) // 3], 5)
    assert np.array_equal(r["pre_blend"], r1["pre_blend"])
    assert np.array_equal(r["field"], r1["field"])


def test_golden_demo_vectors(oracle_mod, demo):
    """Committed golden outputs (tests/golden/make_oracle_golden.py): guards the oracle itself
    against regressions and pins the exact numbers the GPU path must reproduce."""
    g = np.load(GOLD)
    O = oracle_mod
    sc = oracle_scene(O, demo)
    req = O.build_requests(demo["world"])
    f = np.zeros(sc.cells, np.uint32)
    for frame in range(int(g["frames"])):
        r = O.update(sc, f, req, 1 + frame)
        f = r["field"]
        assert r["totals"] == tuple(int(v) for v in g["totals"][frame])
        assert np.array_equal(r["pre_blend"][g["probe_idx"]], g["pre_blend"][frame])
    assert np.array_equal(f, g["field"])
