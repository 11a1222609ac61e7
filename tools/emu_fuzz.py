#!/usr/bin/env python3
"""The kernels on the host SIMT emulator (tests/device_host/simt_emu.hpp) against the oracle on MANY random small worlds --
tests/test_kernels_on_host.py::test_fuzz_worlds_kernels beyond the six cases the suite runs: pooled and per-workgroup
kernel, odd world sizes, out-of-world probes, magicSize dispatches, a loaded brick 0 and ids outside the palette (the
general march through the host flag and through validate_world_kernel's device flag).  No GPU; ~1 s per case.
--shards: random multi-rank frames (ranks, shard modes, pull / push exchange, device-side counts) on emulated ranks.
--builder: the request-list builder kernels on random worlds.
usage: emu_fuzz.py [first_case] [count] [--shards | --builder]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, ROOT)
import test_kernels_on_host as T
import oracle as O

import ctypes as C
import numpy as np
from lum_b200 import scenes

argv = [a for a in sys.argv[1:] if not a.startswith("--")]
first = int(argv[0]) if len(argv) > 0 else 16
count = int(argv[1]) if len(argv) > 1 else 50
lib = T.build(T.OUT)
t0 = time.time()
if "--shards" in sys.argv:
    # random multi-rank frames -- 2..5 emulated ranks, interleaved chunks of 6..48 workgroups or contiguous shards,
    # pull (gather_commit_kernel) or push (inbox + commit_kernel), the count known or on the device, magicSize dispatches,
    # 1..3 chained frames -- every replica against the single-device oracle frames
    small = scenes.synthetic_scene(24, 20, 12, palette=48, seed=1234)
    sc = T.scene_of(O, small)
    world = np.ascontiguousarray(small["world"], np.int16)
    bricks = np.ascontiguousarray(small["bricks"], np.uint8)
    mats = np.ascontiguousarray(small["materials"], np.float32)
    wz, wy, wx = world.shape
    table = np.ascontiguousarray(O.direction_table(64), np.float32)
    sun = np.ascontiguousarray(O.sun_dir(), np.float32)
    all_req = O.build_requests(small["world"])
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    lib.kernels_on_host_update_sharded.restype = C.c_int
    for case in range(first, first + count):
        rng = np.random.default_rng(840000 + case)
        nranks = int(rng.integers(2, 6))
        variant = int(rng.integers(0, 2))
        chunk = int(rng.choice([0, 6, 12, 18, 24, 48]))
        push = int(rng.integers(0, 2))
        slack = int(rng.choice([0, 0, 5, 31]))
        size = int(rng.integers(1, 4))
        shift = int(rng.integers(0, size))
        frames = int(rng.integers(1, 4))
        n = int(rng.integers(size, 70))
        req = np.ascontiguousarray(all_req[rng.permutation(len(all_req))[:n]], np.int16)
        req_buf = np.zeros((len(req) + slack, 4), np.int16); req_buf[:len(req)] = req
        f0 = rng.integers(0, 2 ** 32, sc.cells, dtype=np.uint64).astype(np.uint32)
        ref = f0
        for f in range(frames):
            ref = O.update(sc, ref, req, 4 + f, size=size, shift=shift)["field"]
        replicas = np.ascontiguousarray(np.tile(f0, (nranks, 1)))
        rc = lib.kernels_on_host_update_sharded(C.c_int(wx), C.c_int(wy), C.c_int(wz), C.c_int(bricks.shape[0]), C.c_int(1), C.c_int(64),
                                                C.c_float(0.01), p(world), p(bricks), p(mats), p(replicas), C.c_int(nranks), C.c_int(chunk),
                                                p(req_buf), C.c_longlong(len(req)), C.c_int(4), C.c_int(size), C.c_int(shift), p(table), p(sun),
                                                C.c_int(variant), C.c_int(frames), C.c_int(push), C.c_longlong(slack))
        assert rc == 0, case
        for r in range(nranks):
            assert np.array_equal(replicas[r], ref), (case, r, nranks, variant, chunk, push, slack, size, shift, frames, n)
    print(f"shard cases {first}..{first + count - 1}: every replica equal to the single-device frames ({time.time() - t0:.0f} s)")
elif "--builder" in sys.argv:
    # build_fused_kernel + build_specials_kernel on random worlds (rows that are / are not a multiple of four cells
    # long, ragged last tiles, worlds smaller than a tile, negative ids, random specials incl. duplicates) against the oracle's list
    lib.kernels_on_host_build_requests.restype = C.c_longlong
    p = lambda a: a.ctypes.data_as(C.c_void_p) if a is not None else None
    total = 0
    for case in range(first, first + count):
        rng = np.random.default_rng(850000 + case)
        wx, wy, wz = int(rng.integers(1, 140)), int(rng.integers(1, 40)), int(rng.integers(1, 12))
        dens = float(rng.choice([0.0, 0.003, 0.02, 0.3, 1.0]))
        world = np.where(rng.random((wz, wy, wx)) < dens, rng.integers(-2, 30, (wz, wy, wx)), 0).astype(np.int16)
        nsp = int(rng.integers(0, 30))
        sp = np.stack([rng.integers(0, wx, nsp), rng.integers(0, wy, nsp), rng.integers(0, wz, nsp), np.zeros(nsp, np.int64)], 1).astype(np.int16) if nsp else None
        out = np.zeros((world.size + nsp, 4), np.int16)
        nbuilt = C.c_longlong(0)
        n = lib.kernels_on_host_build_requests(C.c_int(wx), C.c_int(wy), C.c_int(wz), p(world), p(sp), C.c_int(nsp), p(out),
                                               C.c_longlong(len(out)), C.byref(nbuilt))
        ref = O.build_requests(world, sp)
        assert nbuilt.value == len(O.build_requests(world)) and n == len(ref) and np.array_equal(out[:n], ref), (case, wx, wy, wz, dens, nsp)
        total += n
    print(f"builder cases {first}..{first + count - 1}: {total} requests, lists equal to the oracle's ({time.time() - t0:.0f} s)")
else:
    for case in range(first, first + count):
        T.test_fuzz_worlds_kernels(lib, O, case)
    print(f"cases {first}..{first + count - 1}: both kernels bit-exact on the emulator ({time.time() - t0:.0f} s)")
