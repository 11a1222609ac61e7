"""The CUDA kernels themselves on a host SIMT emulator, against the oracle -- no GPU needed.

tests/device_host/simt_emu.hpp runs the __global__ functions of lum_b200/csrc/radiance_kernels.cuh with one OS thread per
CUDA thread (warp collectives as rendezvous, __syncthreads / the named barrier as counting barriers, __shared__ as a
static of the one resident CTA).  tests/device_host/kernels_on_host.cpp launches them the way lumcu.cu's do_update()
does: validate_world_kernel, the trace kernel (pooled: ray pool, lane refill, parking, tail passes, reduction; or one
workgroup per CTA), commit_kernel.  Where tests/test_device_source_on_host.py checks the per-thread arithmetic, this
checks the ORCHESTRATION -- which ray lands in which slot, what a parked ray carries, who reduces what -- before any GPU
time is spent.  Test infrastructure only; the emulator is slow (a barrier per collective), so the scenes are small.
Bar: bit-exact, like on the GPU.
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from test_device_source_on_host import CUDA_INC, HERE, assert_same, scene_of

SRC = os.path.join(HERE, "device_host", "kernels_on_host.cpp")
DEPS = [SRC, __file__] + [os.path.join(HERE, "device_host", f) for f in ("simt_emu.hpp", "cuda_host_shim.hpp")] + \
       [os.path.join(HERE, "..", "lum_b200", "csrc", f) for f in ("radiance_device.cuh", "radiance_kernels.cuh")]
OUT = os.path.join(HERE, "_build", "libkernels_on_host.so")

POOL, WG = 1, 0

pytestmark = pytest.mark.timeout(900)      # pytest-timeout: a deadlocked emulator must fail, not hang the suite


def build(out, defines=()):
    if not os.path.exists(os.path.join(CUDA_INC, "cuda_runtime.h")):
        pytest.skip("CUDA headers not found")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in DEPS):
        r = subprocess.run(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-pthread", "-ffp-contract=off",
                            "-fno-fast-math", "-Wno-unknown-pragmas", *[f"-D{d}" for d in defines], "-I", CUDA_INC,
                            "-o", out, SRC], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
    lib = C.CDLL(out)
    lib.kernels_on_host_update.restype = C.c_int
    return lib


@pytest.fixture(scope="module")
def emu():
    return build(OUT)


def run(lib, O, scn, f0, req, seed, variant, rays=64, size=1, shift=0, reactiveness=0.01, slack=0):
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
    tot = np.zeros(2, np.uint64)
    table = np.ascontiguousarray(O.direction_table(rays), np.float32)
    sun = np.ascontiguousarray(O.sun_dir(), np.float32)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    rc = lib.kernels_on_host_update(C.c_int(wx), C.c_int(wy), C.c_int(wz), C.c_int(bricks.shape[0]),
                                    C.c_int(int(not bricks[0].any())), C.c_int(rays), C.c_float(reactiveness),
                                    p(world), p(bricks), p(mats), p(f0), p(f1), p(req), C.c_longlong(len(req)),
                                    C.c_int(seed), C.c_int(size), C.c_int(shift), p(table), p(sun), C.c_int(variant),
                                    C.c_int(0), C.c_int(1), p(pre), p(rs), p(rh), p(tot), C.c_longlong(slack))
    assert rc == 0
    return dict(field=f1, pre_blend=pre, ray_samples=rs, ray_hits=rh, totals=(int(tot[0]), int(tot[1])))


def check(lib, O, scn, req, seed, variant, rays=64, size=1, shift=0, field_seed=3, slack=0):
    sc = scene_of(O, scn, rays=rays)
    f0 = np.random.default_rng(field_seed).integers(0, 2 ** 32, sc.cells, dtype=np.uint64).astype(np.uint32)
    ref = O.update(sc, f0, req, seed, size=size, shift=shift, want_rays=True)
    got = run(lib, O, scn, f0, req, seed, variant, rays=rays, size=size, shift=shift, slack=slack)
    assert_same(got, ref)
    assert got["totals"] == ref["totals"][:2]


@pytest.mark.parametrize("variant", [POOL, WG], ids=["pooled", "workgroup"])
def test_demo_world_kernels(emu, oracle_mod, demo, variant):
    """61 probes of the demo world (a ragged last batch for the pooled kernel: 10 full CTAs + 1 probe)."""
    O = oracle_mod
    check(emu, O, demo, O.build_requests(demo["world"])[1000:1061], 5, variant)


@pytest.mark.parametrize("variant", [POOL, WG], ids=["pooled", "workgroup"])
def test_dispatch_cut_down_to_device_side_count(emu, oracle_mod, demo, variant):
    """lumcu_build_requests without a host sync: the launch is sized for an upper bound (here 100 requests more than the
    list holds) and trace and commit kernels stop at the count they read from device memory."""
    O = oracle_mod
    req = O.build_requests(demo["world"])[1000:1061]
    check(emu, O, demo, req, 5, variant, slack=100)
    check(emu, O, demo, req[:60], 6, variant, size=3, shift=1, slack=7)


@pytest.mark.parametrize("rays", [16, 64, 256])
def test_dense_world_pooled_kernel_ray_counts(emu, oracle_mod, small_synth, rays):
    """The C4 generator's dense world: long and short rays side by side, so lanes refill, warps park and the tail
    warps run; 16 (24 probes per CTA), 64 and 256 rays per probe (256-thread CTAs)."""
    O = oracle_mod
    n = {16: 50, 64: 20, 256: 7}[rays]
    check(emu, O, small_synth, O.build_requests(small_synth["world"])[::97][:n], 11, POOL, rays=rays)


def test_dense_world_workgroup_kernel(emu, oracle_mod, small_synth):
    O = oracle_mod
    check(emu, O, small_synth, O.build_requests(small_synth["world"])[::131][:24], 2, WG)
    check(emu, O, small_synth, O.build_requests(small_synth["world"])[::131][:24], 2, WG, rays=16)


@pytest.mark.parametrize("case", [0, 1, 2, 7, 9, 15])
def test_fuzz_worlds_kernels(emu, oracle_mod, case):
    """Cases of the GPU fuzz (tests/test_parity_gpu.py::test_random_small_worlds): odd sizes, out-of-world probes,
    magicSize dispatch, a loaded brick 0 (cases 0, 9, 15) and an id outside the palette (7, 15): the general march,
    selected by the host flag and by validate_world_kernel's device flag."""
    O = oracle_mod
    rng = np.random.default_rng(1000 + case)
    wx, wy, wz = (int(v) for v in rng.integers(1, [9, 8, 5]))
    P = int(rng.integers(2, 41))
    bricks = ((rng.random((P, 16, 16, 16)) < rng.choice([0.002, 0.05, 0.3, 0.9])) * rng.integers(1, 256, (P, 16, 16, 16))).astype(np.uint8)
    if case % 3:
        bricks[0] = 0
    world = ((rng.random((wz, wy, wx)) < rng.choice([0.1, 0.5, 1.0])) * rng.integers(1, P, (wz, wy, wx))).astype(np.int16)
    if case % 8 == 7:
        world.reshape(-1)[0] = P + 3
    mats = rng.random((256, 6)).astype(np.float32)
    mats[:, 4] *= rng.choice([0.0, 1.0, 18.0])
    scn = dict(world=world, bricks=bricks, materials=mats)
    inside = np.array([(x, y, z, 0) for z in range(wz) for y in range(wy) for x in range(wx)], np.int16)
    outside = np.array([(-1, 0, 0, 0), (wx, 0, 0, 0), (0, -2, wz, 0), (wx + 3, wy + 1, -1, 0)], np.int16)
    req = np.concatenate([inside, outside])
    req = req[rng.permutation(len(req))][: max(1, min(40, int(rng.integers(1, len(req) + 1))))]
    size = min(int(rng.integers(1, 4)), len(req))
    shift = int(rng.integers(0, size))
    seed = int(rng.integers(-2 ** 31, 2 ** 31))
    for variant in (POOL, WG):
        check(emu, O, scn, req, seed, variant, size=size, shift=shift, field_seed=case)


def test_request_builder_kernels(emu, oracle_mod, demo):
    """SURVEY a13 on the emulator: build_fused_kernel (rule, look-back scan over 36 tiles, ordered scatter) and
    build_specials_kernel against the oracle's list (demo world: 2,689 requests in z,y,x order; specials appended in
    submission order unless already selected)."""
    O = oracle_mod
    emu.kernels_on_host_build_requests.restype = C.c_longlong
    world = np.ascontiguousarray(demo["world"], np.int16)
    wz, wy, wx = world.shape
    p = lambda a: a.ctypes.data_as(C.c_void_p) if a is not None else None
    rng = np.random.default_rng(0)
    specials = np.array([(int(rng.integers(0, wx)), int(rng.integers(0, wy)), int(rng.integers(0, wz)), 0) for _ in range(40)]
                        + [(3, 3, 15, 0), (3, 3, 15, 0)], np.int16)          # duplicates among specials are kept
    for sp in (None, specials):
        nsp = 0 if sp is None else len(sp)
        out = np.zeros((world.size + nsp, 4), np.int16)
        nbuilt = C.c_longlong(0)
        n = emu.kernels_on_host_build_requests(C.c_int(wx), C.c_int(wy), C.c_int(wz), p(world), p(sp), C.c_int(nsp),
                                               p(out), C.c_longlong(len(out)), C.byref(nbuilt))
        ref = O.build_requests(world, sp)
        assert nbuilt.value == 2689
        assert n == len(ref)
        assert np.array_equal(out[:n], ref)


@pytest.mark.parametrize("dims,density", [((37, 11, 9), 0.02), ((64, 33, 5), 0.004), ((4, 4, 1), 0.5), ((1, 1, 1), 1.0),
                                          ((130, 3, 7), 0.3), ((256, 12, 3), 0.0)])
def test_request_builder_shapes(emu, oracle_mod, dims, density):
    """Rows that are / are not a multiple of four cells long (the 8-byte row loads vs the cell-by-cell rule), a
    ragged last tile, worlds smaller than a tile, negative ids (they lower the neighbourhood sum), an empty world."""
    O = oracle_mod
    emu.kernels_on_host_build_requests.restype = C.c_longlong
    wx, wy, wz = dims
    rng = np.random.default_rng(wx * 1000 + wy)
    world = np.where(rng.random((wz, wy, wx)) < density, rng.integers(-2, 30, (wz, wy, wx)), 0).astype(np.int16)
    sp = np.array([(int(rng.integers(0, wx)), int(rng.integers(0, wy)), int(rng.integers(0, wz)), 0) for _ in range(9)], np.int16)
    out = np.zeros((world.size + len(sp), 4), np.int16)
    nbuilt = C.c_longlong(0)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    n = emu.kernels_on_host_build_requests(C.c_int(wx), C.c_int(wy), C.c_int(wz), p(world), p(sp), C.c_int(len(sp)),
                                           p(out), C.c_longlong(len(out)), C.byref(nbuilt))
    ref = O.build_requests(world, sp)
    assert nbuilt.value == len(O.build_requests(world))
    assert n == len(ref) and np.array_equal(out[:n], ref)


@pytest.mark.parametrize("dims", [(16, 5, 3), (13, 4, 2), (4, 3, 2), (3, 2, 2), (64, 2, 1)])
@pytest.mark.parametrize("slack", [0, 9], ids=["count-known", "count-on-device"])
def test_commit_kernels_on_crafted_lists(emu, dims, slack):
    """The store side alone: commit_vec4_kernel -- one 16-byte store where four consecutive requests are four
    x-adjacent texels at an aligned cell, element by element otherwise -- against commit_kernel and a numpy scatter, on
    lists made to hit every branch: full rows in z,y,x order (rows of 16 / 64 texels: aligned quads; 13, 3: quads that
    straddle rows or start misaligned), every other texel, runs that start at x = 1, probes outside the world on every
    side, a ragged tail of 1..3 requests, and a launch sized for more requests than the device-side count holds."""
    wx, wy, wz = dims
    rng = np.random.default_rng(wx * 100 + wy * 10 + wz + slack)
    full = np.array([(x, y, z, 0) for z in range(wz) for y in range(wy) for x in range(wx)], np.int16)
    outside = np.array([(-1, 0, 0, 0), (wx, 0, 0, 0), (0, wy, 0, 0), (0, 0, -1, 0), (wx - 1, wy - 1, wz, 0), (-4, -4, -4, 0),
                        (-1, 0, 0, 0)], np.int16)
    lists = {
        "rows": full,
        "every other texel": full[::2],
        "runs from x = 1": full[full[:, 0] >= 1],
        "outside probes in between": np.concatenate([full[:5], outside, full[5:]]),
        "ragged tails": np.concatenate([full[: 4 * (len(full) // 8)], full[-3:]]),
        "w component set": np.concatenate([full[:8] + np.array([0, 0, 0, 3], np.int16), full[8:]]),
        "shuffled": full[rng.permutation(len(full))],
    }
    emu.kernels_on_host_commit.restype = C.c_int
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    for name, req in lists.items():
        req = np.ascontiguousarray(req, np.int16)
        staged = rng.integers(1, 2 ** 32, len(req), dtype=np.uint64).astype(np.uint32)
        f0 = rng.integers(0, 2 ** 32, wx * wy * wz, dtype=np.uint64).astype(np.uint32)
        ref = f0.copy()
        for (x, y, z, _), v in zip(req.tolist(), staged.tolist()):            # no texel appears twice in these lists
            if 0 <= x < wx and 0 <= y < wy and 0 <= z < wz:
                ref[x + wx * (y + wy * z)] = v
        for vec in (1, 0):
            got = f0.copy()
            assert emu.kernels_on_host_commit(C.c_int(wx), C.c_int(wy), C.c_int(wz), p(req), C.c_longlong(len(req)), p(staged),
                                              p(got), C.c_int(vec), C.c_longlong(slack)) == 0
            assert np.array_equal(got, ref), (name, "vec4" if vec else "element-wise")


@pytest.mark.parametrize("d", [(1, 0, 0), (0, -2, 1), (-3, 4, -1), (0, 0, 0), (48, 0, 0)])
def test_shift_kernel(emu, oracle_mod, d):
    O = oracle_mod
    wx, wy, wz = 12, 10, 6
    f = np.random.default_rng(1).integers(0, 2 ** 32, wx * wy * wz, dtype=np.uint64).astype(np.uint32)
    ref = O.shift_field(f.copy(), (wx, wy, wz), *d)
    got = f.copy()
    emu.kernels_on_host_shift(C.c_int(wx), C.c_int(wy), C.c_int(wz), got.ctypes.data_as(C.c_void_p), *[C.c_int(v) for v in d])
    assert np.array_equal(got, np.asarray(ref).reshape(-1))


def test_emulated_kernels_are_race_free_under_tsan(defines=()):
    """Host-side racecheck: tests/device_host/tsan_main.cpp runs the emulated pooled and per-workgroup kernels (dense
    random world: refill, parking, tail passes) under ThreadSanitizer.  Every CUDA thread is an OS thread and every
    barrier / warp collective a real synchronisation, so a missing __syncthreads / __syncwarp in the kernels is a
    data race TSAN reports (removing the barrier after pass 0 of the march yields 31 reports)."""
    if not os.path.exists(os.path.join(CUDA_INC, "cuda_runtime.h")):
        pytest.skip("CUDA headers not found")
    src = os.path.join(HERE, "device_host", "tsan_main.cpp")
    exe = os.path.join(HERE, "_build", "tsan_" + ("optin" if defines else "default"))
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    r = subprocess.run(["g++", "-fsanitize=thread", "-O1", "-g", "-std=c++17", "-pthread", "-ffp-contract=off",
                        "-fno-fast-math", "-Wno-unknown-pragmas", *[f"-D{d}" for d in defines], "-I", CUDA_INC,
                        "-o", exe, src], capture_output=True, text=True)
    if r.returncode != 0 and "tsan" in r.stderr.lower():
        pytest.skip("ThreadSanitizer runtime not available")
    assert r.returncode == 0, r.stderr
    run_ = subprocess.run([exe], capture_output=True, text=True, timeout=600,
                          env=dict(os.environ, TSAN_OPTIONS="exitcode=66"))
    if "FATAL: ThreadSanitizer" in run_.stderr:          # e.g. an unsupported memory layout inside a sandbox
        pytest.skip("ThreadSanitizer cannot run here: " + run_.stderr.splitlines()[0])
    assert "WARNING: ThreadSanitizer" not in run_.stderr, run_.stderr[:4000]
    assert run_.returncode == 0, run_.stderr[:2000]
    assert run_.stdout.count("S=") == 4


def test_copy_bricks_kernel(emu, oracle_mod, demo):
    """exec_copies: static bricks copied into dynamic ones, source 0 zeroes the destination."""
    O = oracle_mod
    rng = np.random.default_rng(0)
    bricks = np.concatenate([demo["bricks"], rng.integers(0, 256, (49, 16, 16, 16)).astype(np.uint8)])   # stale dynamic bricks
    src = rng.integers(0, 15, 40).astype(np.int32)
    dst = np.arange(20, 60, dtype=np.int32)
    ref = bricks.copy()
    O.copy_bricks(ref, src, dst)
    got = np.ascontiguousarray(bricks.copy())
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    assert emu.kernels_on_host_copy_bricks(p(got), C.c_int(len(got)), p(src), p(dst), C.c_int(len(src))) == 0
    assert np.array_equal(got, ref)
    assert (got[dst[src == 0]] == 0).all()


def test_slab_bound_kernel(emu, oracle_mod, demo):
    """The z-slab boundaries of the multi-GPU slab sharding (lumcu_slab_range) against a binary search."""
    O = oracle_mod
    req = np.ascontiguousarray(O.build_requests(demo["world"]), np.int16)
    z = req[:, 2].astype(np.int64)
    out = np.zeros(2, np.int64)
    for size, shift in ((1, 0), (3, 1)):
        zs = z[shift::size][: len(req) // size]
        for zb, ze in ((0, 16), (4, 8), (8, 12), (15, 16), (-5, 3), (16, 40)):
            emu.kernels_on_host_slab_range(req.ctypes.data_as(C.c_void_p), C.c_longlong(len(zs)), C.c_int(size), C.c_int(shift),
                                           C.c_int(zb), C.c_int(ze), out.ctypes.data_as(C.c_void_p))
            assert (int(out[0]), int(out[1])) == (int(np.searchsorted(zs, zb, "left")), int(np.searchsorted(zs, ze, "left")))


@pytest.mark.parametrize("slack", [0, 31], ids=["count-known", "count-on-device"])
@pytest.mark.parametrize("push", [0, 1], ids=["pull", "push"])
@pytest.mark.parametrize("chunk", [12, 0], ids=["interleaved", "contiguous"])
@pytest.mark.parametrize("nranks,variant", [(2, POOL), (3, POOL), (2, WG)])
def test_interleaved_shards_with_peer_commit(emu, oracle_mod, small_synth, nranks, variant, chunk, push, slack):
    """SURVEY 8e on the emulator: three frames dealt to `nranks` ranks (contiguous shards or interleaved chunks, il_map /
    shard_window in the kernels), arrive_kernel, then the transfer -- pull: gather_commit_kernel reads every texel from its
    owner's staged[] (alternating halves); push: the trace stored into every rank's inbox and commit_kernel scatters locally
    -- behind the flag wait: all replicas equal the single-device field."""
    O = oracle_mod
    sc = scene_of(O, small_synth)
    req = np.ascontiguousarray(O.build_requests(small_synth["world"])[::53][:45], np.int16)      # 45 probes, chunks of 12
    req_buf = np.zeros((len(req) + slack, 4), np.int16); req_buf[:len(req)] = req                 # the list buffer holds the bound
    f0 = np.random.default_rng(8).integers(0, 2 ** 32, sc.cells, dtype=np.uint64).astype(np.uint32)
    ref = f0
    for f in range(3):
        ref = O.update(sc, ref, req, 4 + f)["field"]
    world = np.ascontiguousarray(small_synth["world"], np.int16)
    bricks = np.ascontiguousarray(small_synth["bricks"], np.uint8)
    mats = np.ascontiguousarray(small_synth["materials"], np.float32)
    wz, wy, wx = world.shape
    replicas = np.ascontiguousarray(np.tile(f0, (nranks, 1)))
    table = np.ascontiguousarray(O.direction_table(64), np.float32)
    sun = np.ascontiguousarray(O.sun_dir(), np.float32)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    emu.kernels_on_host_update_sharded.restype = C.c_int
    rc = emu.kernels_on_host_update_sharded(C.c_int(wx), C.c_int(wy), C.c_int(wz), C.c_int(bricks.shape[0]), C.c_int(1),
                                            C.c_int(64), C.c_float(0.01), p(world), p(bricks), p(mats), p(replicas),
                                            C.c_int(nranks), C.c_int(chunk), p(req_buf), C.c_longlong(len(req)), C.c_int(4),
                                            C.c_int(1), C.c_int(0), p(table), p(sun), C.c_int(variant), C.c_int(3), C.c_int(push), C.c_longlong(slack))
    assert rc == 0
    for r in range(nranks):
        assert np.array_equal(replicas[r], ref), r
