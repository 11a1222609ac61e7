"""The CUDA kernels' per-thread arithmetic, compiled for the host, against the oracle -- no GPU needed.

lum_b200/csrc/radiance_device.cuh (ray setup, RNG, the march variants -- general, fast, fused, single-step --, pending-sample resolution,
refinement, material fetch, field gather, shading, unorm10 pack/unpack, glossy_fragment and map_model_invocation,
the bodies of the glossy and map kernels) is compiled with g++ through
tests/device_host/cuda_host_shim.hpp and run one "thread" after another by tests/device_host/device_on_host.cpp,
which restates only the workgroup reduction and the blend around it.  This is test infrastructure: it checks the
device SOURCE where there is no device (a change to the march that breaks parity shows up here before any GPU
time is spent); the `-m gpu` tests remain the parity tests of the kernels proper.  Bar: bit-exact, like on the GPU.
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "device_host", "device_on_host.cpp")
DEPS = [SRC, os.path.join(HERE, "device_host", "cuda_host_shim.hpp"),
        os.path.join(HERE, "..", "lum_b200", "csrc", "radiance_device.cuh")]
OUT = os.path.join(HERE, "_build", "libdevice_on_host.so")
CUDA_INC = os.environ.get("CUDA_HOME", "/usr/local/cuda") + "/include"

MODES = {"general": 0, "fast": 1, "fused": 2, "single": 3}


def build_host_lib(out, defines=()):
    if not os.path.exists(os.path.join(CUDA_INC, "cuda_runtime.h")):
        pytest.skip("CUDA headers not found")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in DEPS + [__file__]):
        # -ffp-contract=off = nvcc -fmad=false: one rounding per written operation (DESIGN.md decree 4)
        r = subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math",
                            "-fopenmp", "-Wno-unknown-pragmas", *[f"-D{d}" for d in defines], "-I", CUDA_INC, "-o", out, SRC],
                           capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
    lib = C.CDLL(out)
    lib.device_on_host_update.restype = C.c_int
    return lib


@pytest.fixture(scope="module")
def host_lib():
    return build_host_lib(OUT)


def run(lib, O, scn, f0, req, seed, mode, rays=64, size=1, shift=0, reactiveness=0.01):
    world = np.ascontiguousarray(scn["world"], np.int16)
    bricks = np.ascontiguousarray(scn["bricks"], np.uint8)
    mats = np.ascontiguousarray(scn["materials"], np.float32)
    wz, wy, wx = world.shape
    req = np.ascontiguousarray(req, np.int16).reshape(-1, 4)
    nwg = len(req) // size
    f0 = np.ascontiguousarray(f0, np.uint32).reshape(-1)
    f1 = np.empty_like(f0)
    pre = np.zeros((nwg, 3), np.float32)
    rs = np.zeros(nwg * rays, np.uint32)
    rh = np.zeros(nwg * rays, np.uint8)
    table = np.ascontiguousarray(O.direction_table(rays), np.float32)
    sun = np.ascontiguousarray(O.sun_dir(), np.float32)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    rc = lib.device_on_host_update(C.c_int(wx), C.c_int(wy), C.c_int(wz), C.c_int(bricks.shape[0]),
                                   C.c_int(int(not bricks[0].any())), C.c_int(rays), C.c_float(reactiveness),
                                   p(world), p(bricks), p(mats), p(f0), p(f1), p(req), C.c_longlong(len(req)),
                                   C.c_int(seed), C.c_int(size), C.c_int(shift), p(table), p(sun), C.c_int(mode),
                                   p(pre), p(rs), p(rh))
    assert rc == 0
    return dict(field=f1, pre_blend=pre, ray_samples=rs, ray_hits=rh)


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def assert_same(got, ref):
    assert np.array_equal(got["ray_samples"], ref["ray_samples"])
    assert np.array_equal(got["ray_hits"], ref["ray_hits"])
    nan = np.isnan(ref["pre_blend"])
    assert np.array_equal(np.isnan(got["pre_blend"]), nan)
    assert np.array_equal(got["pre_blend"][~nan], ref["pre_blend"][~nan])          # value-equal (+0 == -0)
    nz = ~nan & (ref["pre_blend"] != 0)
    assert np.array_equal(bits(got["pre_blend"])[nz], bits(ref["pre_blend"])[nz])    # and bit-equal where non-zero
    assert np.array_equal(got["field"], ref["field"])


def scene_of(O, scn, rays=64, reactiveness=0.01):
    return O.Scene(scn["world"], scn["bricks"], scn["materials"], rays_per_probe=rays, reactiveness=reactiveness)


@pytest.mark.parametrize("mode", list(MODES))
def test_demo_world_full_recompute(host_lib, oracle_mod, demo, mode):
    """BASELINE configs[1]: all 2,689 probes x 64 rays of the demo world on a random field, each march variant."""
    O = oracle_mod
    sc = scene_of(O, demo)
    req = O.build_requests(demo["world"])
    assert len(req) == 2689
    f0 = np.random.default_rng(3).integers(0, 2 ** 32, sc.cells, dtype=np.uint64).astype(np.uint32)
    ref = O.update(sc, f0, req, 5, want_rays=True)
    assert_same(run(host_lib, O, demo, f0, req, 5, MODES[mode]), ref)


@pytest.mark.parametrize("rays", [16, 32, 64, 128, 256])
def test_synthetic_all_ray_counts(host_lib, oracle_mod, small_synth, rays):
    """Dense random world (the C4 generator at 24x20x12), every supported rays-per-probe, fast and fused march."""
    O = oracle_mod
    sc = scene_of(O, small_synth, rays=rays)
    req = O.build_requests(small_synth["world"])[::7]
    f0 = np.random.default_rng(rays).integers(0, 2 ** 32, sc.cells, dtype=np.uint64).astype(np.uint32)
    ref = O.update(sc, f0, req, 11, want_rays=True)
    for mode in ("fast", "fused", "single"):
        assert_same(run(host_lib, O, small_synth, f0, req, 11, MODES[mode], rays=rays), ref)


def test_synthetic_64x64x16_full_update(host_lib, oracle_mod):
    """BASELINE configs[4]'s smallest world, whole (65,456 probes, 4.2 M rays), fused and single-step march."""
    from lum_b200 import scenes
    O = oracle_mod
    syn = scenes.synthetic_scene(64, 64, 16)
    sc = scene_of(O, syn)
    req = O.build_requests(syn["world"])
    f0 = np.random.default_rng(0).integers(0, 2 ** 32, sc.cells, dtype=np.uint64).astype(np.uint32)
    ref = O.update(sc, f0, req, 7, want_rays=True)
    for mode in ("fused", "single"):
        assert_same(run(host_lib, O, syn, f0, req, 7, MODES[mode]), ref)


@pytest.mark.parametrize("case", range(24))
def test_random_small_worlds(host_lib, oracle_mod, case):
    """The fuzz of tests/test_parity_gpu.py::test_random_small_worlds on the host build: odd world sizes, sparse to
    solid bricks, out-of-world probes, ragged workgroup counts, random fields and seeds.  Scenes the device would
    send down the general march (brick 0 loaded) run mode 0 only."""
    O = oracle_mod
    rng = np.random.default_rng(1000 + case)
    wx, wy, wz = (int(v) for v in rng.integers(1, [9, 8, 5]))
    P = int(rng.integers(2, 41))
    bricks = (rng.random((P, 16, 16, 16)) < rng.choice([0.002, 0.05, 0.3, 0.9])) * rng.integers(1, 256, (P, 16, 16, 16))
    bricks = bricks.astype(np.uint8)
    if case % 3:
        bricks[0] = 0
    world = ((rng.random((wz, wy, wx)) < rng.choice([0.1, 0.5, 1.0])) * rng.integers(1, P, (wz, wy, wx))).astype(np.int16)
    mats = rng.random((256, 6)).astype(np.float32)
    mats[:, 4] *= rng.choice([0.0, 1.0, 18.0])
    scn = dict(world=world, bricks=bricks, materials=mats)
    sc = scene_of(O, scn)
    inside = np.array([(x, y, z, 0) for z in range(wz) for y in range(wy) for x in range(wx)], np.int16)
    outside = np.array([(-1, 0, 0, 0), (wx, 0, 0, 0), (0, -2, wz, 0), (wx + 3, wy + 1, -1, 0)], np.int16)
    req = np.concatenate([inside, outside])
    req = req[rng.permutation(len(req))][: max(1, int(rng.integers(1, len(req) + 1)))]
    f0 = rng.integers(0, 2 ** 32, sc.cells, dtype=np.uint64).astype(np.uint32)
    size = min(int(rng.integers(1, 4)), len(req))
    shift = int(rng.integers(0, size))
    seed = int(rng.integers(-2 ** 31, 2 ** 31))
    ref = O.update(sc, f0, req, seed, size=size, shift=shift, want_rays=True)
    modes = ("general", "fast", "fused", "single") if not bricks[0].any() else ("general",)
    for mode in modes:
        assert_same(run(host_lib, O, scn, f0, req, seed, MODES[mode], size=size, shift=shift), ref)


# ---- the glossy pass: glossy_fragment() is the body of glossy_kernel ------------------------------------------
def run_glossy(lib, scn, field, view, mat_norm, depth):
    world = np.ascontiguousarray(scn["world"], np.int16)
    bricks = np.ascontiguousarray(scn["bricks"], np.uint8)
    mats = np.ascontiguousarray(scn["materials"], np.float32)
    wz, wy, wx = world.shape
    field = np.ascontiguousarray(field, np.uint32).reshape(-1)
    mn = np.ascontiguousarray(mat_norm, np.uint8)
    dp = np.ascontiguousarray(depth, np.float32)
    h, w = mn.shape[:2]
    v = np.concatenate([np.asarray(view[k], np.float32).reshape(-1) for k in
                        ("campos", "camdir", "horizline_scaled", "vertiline_scaled", "global_light_dir", "frame_size")])
    v = np.ascontiguousarray(v, np.float32)
    assert v.size == 17
    frame = np.zeros((h, w, 4), np.float32)
    cov = np.zeros((h, w), np.uint8)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    lib.device_on_host_glossy.restype = C.c_int
    rc = lib.device_on_host_glossy(C.c_int(wx), C.c_int(wy), C.c_int(wz), C.c_int(bricks.shape[0]),
                                   C.c_int(int(not bricks[0].any())), p(world), p(bricks), p(mats), p(field), p(v),
                                   C.c_int(w), C.c_int(h), p(mn), p(dp), p(frame), p(cov))
    assert rc == 0
    return frame, cov


def assert_frames_equal(got, ref):
    assert np.array_equal(got, ref)                     # -0.0 == +0.0 allowed, everything else bit for bit
    nz = ref != 0
    assert np.array_equal(bits(got)[nz], bits(ref)[nz])


@pytest.mark.parametrize("tag", ["demo", "demo_all", "synth"])
def test_glossy_fragment_matches_oracle(host_lib, oracle_mod, tag):
    """The cases of tests/test_glossy_gpu.py on the host build (default and opt-in march)."""
    from glossy_cases import make_case
    O = oracle_mod
    c = make_case(O, tag)
    ref = O.glossy(c["scene"], c["field"], c["view"], c["mat_norm"], c["depth"])
    got, cov = run_glossy(host_lib, c["scn"], c["field"], c["view"], c["mat_norm"], c["depth"])
    assert np.array_equal(cov, ref["covered"]) and cov.any()
    assert_frames_equal(got, ref["frame"])


@pytest.mark.parametrize("seed", range(4))
def test_glossy_fragment_random_gbuffers(host_lib, oracle_mod, seed):
    """Arbitrary G-buffers (origins inside solids and outside the world, unnormalised normals); seed 3 loads a
    brick 0 (general march)."""
    from glossy_cases import make_case
    O = oracle_mod
    c = make_case(O, "synth")
    scn = {k: v.copy() for k, v in c["scn"].items()}
    rng = np.random.default_rng(seed)
    if seed == 3:
        scn["bricks"][0] = (rng.random((16, 16, 16)) < 0.01) * 5
    sc = O.Scene(scn["world"], scn["bricks"], scn["materials"])
    h, w = 37 + seed, 61 - seed
    mn = rng.integers(0, 256, (h, w, 4), dtype=np.uint8)
    dp = rng.random((h, w), dtype=np.float32) * np.float32(0.6)
    view = O.camera_view((rng.random(3) * 300 - 50), (0.3 + seed, 0.9 - seed, -0.6), w, h, pixels_in_voxel=0.2)
    ref = O.glossy(sc, c["field"], view, mn, dp)
    for lib in (host_lib,):
        got, cov = run_glossy(lib, scn, c["field"], view, mn, dp)
        assert np.array_equal(cov, ref["covered"]) and cov.any()
        assert_frames_equal(got, ref["frame"])


# ---- map.comp: map_model_invocation() is the body of map_model_kernel -----------------------------------------
def _quat_yaw(a):
    import math
    return (np.float32(math.cos(a / 2)), np.float32(0), np.float32(0), np.float32(math.sin(a / 2)))


def _quat_axis(axis, a):
    import math
    ax = np.asarray(axis, np.float64); ax = ax / np.linalg.norm(ax)
    s = math.sin(a / 2)
    return tuple(np.float32(v) for v in (math.cos(a / 2), ax[0] * s, ax[1] * s, ax[2] * s))


MAP_PLACEMENTS = [           # the cases of tests/test_feeders_gpu.py::test_map_model_matches_oracle
    ("identity", (1, 0, 0, 0), (13.1 * 16, 14.1 * 16, 3.1 * 16)),
    ("yaw", _quat_yaw(0.7), (300.0, 310.0, 40.0)),
    ("tilted", _quat_axis((1, 2, 3), 1.1), (400.5, 200.25, 100.75)),
    ("low corner, partly outside", _quat_yaw(-2.1), (12.0, 17.0, 30.0)),
    ("high corner, partly outside", _quat_axis((0, 1, 0), 0.3), (750.0, 755.0, 245.0)),
    ("fully outside", (1, 0, 0, 0), (-400.0, -400.0, -400.0)),
]


@pytest.mark.parametrize("name,rot,shift", MAP_PLACEMENTS, ids=[p[0] for p in MAP_PLACEMENTS])
def test_map_model_invocation_matches_oracle(host_lib, oracle_mod, demo, name, rot, shift):
    O = oracle_mod
    host_lib.device_on_host_map_model.restype = C.c_int
    p_ = lambda a: a.ctypes.data_as(C.c_void_p)
    for mname in ("model_tank_body", "model_robot"):
        model = np.ascontiguousarray(demo[mname], np.uint8)
        sz, sy, sx = model.shape
        pl = O.place_model(rot, shift, (sx, sy, sz), (48, 48, 16))
        cur, src, dst = O.blockify(demo["world"], 15, [pl])
        bricks = np.zeros((512, 16, 16, 16), np.uint8); bricks[:15] = demo["bricks"]
        O.copy_bricks(bricks, src, dst)
        got = bricks.copy()
        O.map_model(cur, bricks, model, pl, 15)
        cur = np.ascontiguousarray(cur, np.int16)
        inv = np.array(pl.inverse_trans, np.float32)
        sh = np.array(list(pl.map_shift), np.int32)
        ar = np.array(list(pl.map_area), np.int32)
        rc = host_lib.device_on_host_map_model(C.c_int(48), C.c_int(48), C.c_int(16), C.c_int(512), p_(cur), p_(got),
                                               p_(model), C.c_int(sx), C.c_int(sy), C.c_int(sz), p_(inv), p_(sh), p_(ar),
                                               C.c_int(15))
        assert rc == 0
        assert np.array_equal(got, bricks), mname
