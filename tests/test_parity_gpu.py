"""GPU parity tests: the CUDA path, called through the C-ABI (include/lumcu.h), against the CPU
oracle on the same inputs.  Bar: per-probe pre-blend mean radiance and the
packed field are BIT-EXACT (integer/byte work and f32 with one rounding per operation on both
sides); the stated north_star tolerance 1e-5*max(1,|x|) is therefore met with margin 0."""
import os

import numpy as np
import pytest

from conftest import make_stage, oracle_scene
from lum_b200 import DEBUG_PRE_BLEND, DEBUG_RAYS, scenes

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(__file__), "golden", "oracle_demo_golden.npz")
REF_GOLD = os.path.join(os.path.dirname(__file__), "golden", "ref_glsl_golden.npz")
REF_REL_TOL = 1e-6      # see tests/test_ref_glsl.py: texture-coordinate un-normalisation in the shim


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def assert_update_equal(st, O, sc, f0, req, seed, size=1, shift=0, rays=True):
    """Both kernel variants (1 = pooled production kernel, 0 = one workgroup per CTA), instrumented
    and plain, against the oracle: sample counts, hits, pre-blend radiance and field bit for bit."""
    st.set_requests(req)
    nwg = len(req) // size
    ref = O.update(sc, f0, req, seed, size=size, shift=shift, want_rays=rays)
    for variant in (1, 0):
        st.set_kernel_variant(variant)
        st.upload_field(f0)
        st.set_debug(DEBUG_PRE_BLEND | (DEBUG_RAYS if rays else 0))
        st.update(seed, size, shift)
        dbg = st.download_debug(nwg, rays=rays)
        got = st.download_field()
        if rays:
            assert np.array_equal(dbg["ray_samples"], ref["ray_samples"]), variant
            assert np.array_equal(dbg["ray_hits"], ref["ray_hits"]), variant
        assert dbg["totals"] == ref["totals"], variant
        # -0.0 == +0.0 is allowed; everything else must match bit for bit
        assert np.array_equal(dbg["pre_blend"], ref["pre_blend"]), variant
        nz = ref["pre_blend"] != 0
        assert np.array_equal(bits(dbg["pre_blend"])[nz], bits(ref["pre_blend"])[nz]), variant
        assert np.array_equal(got, ref["field"]), variant
        # the non-instrumented kernel gives the same field
        st.upload_field(f0)
        st.set_debug(0)
        st.update(seed, size, shift)
        assert np.array_equal(st.download_field(), ref["field"]), variant
    st.set_kernel_variant(1)
    return ref


def test_direction_table_matches_oracle(oracle_mod, demo):
    st = make_stage(demo)
    assert np.array_equal(bits(st.direction_table()), bits(oracle_mod.direction_table(64)))


def test_unpack_all_codes(oracle_mod, demo):
    """unorm10 -> float on the device (reciprocal + FMA correction) == IEEE c/1023 for all codes."""
    st = make_stage(demo)
    exp = (np.arange(1024, dtype=np.float32) / np.float32(1023.0)).astype(np.float32)
    assert np.array_equal(bits(st.debug_unpack_table()), bits(exp))
    for c in (0, 1, 511, 512, 1022, 1023):
        assert oracle_mod.unpack(c)[0] == exp[c]


def test_demo_single_update_config1(oracle_mod, demo):
    """BASELINE config 1/2: demo world, all 2,689 probes x 64 rays, time_seed 1, zero field."""
    O = oracle_mod
    st = make_stage(demo)
    sc = oracle_scene(O, demo)
    req = O.build_requests(demo["world"])
    ref = assert_update_equal(st, O, sc, np.zeros(sc.cells, np.uint32), req, 1)
    assert ref["totals"][2] == 2689


def test_demo_golden_vectors(demo):
    """Committed golden vectors (no oracle code involved at run time)."""
    g = np.load(GOLD)
    st = make_stage(demo)
    n = st.build_requests()
    assert n == 2689
    st.set_debug(DEBUG_PRE_BLEND)
    for frame in range(int(g["frames"])):
        st.update(1 + frame)
        d = st.download_debug(n)
        assert d["totals"] == tuple(int(v) for v in g["totals"][frame])
        assert np.array_equal(d["pre_blend"][g["probe_idx"]], g["pre_blend"][frame])
    assert np.array_equal(st.download_field(), g["field"])


@pytest.mark.parametrize("tag", ["demo", "demo3", "synth"])
def test_reference_shader_golden_vectors(demo, tag):
    """CUDA against the outputs of the reference's OWN shaders/radiance.comp (executed on the CPU in
    the build container, tests/golden/make_ref_glsl_golden.py); no oracle code runs here.
    Packed field words equal; pre-blend radiance within REF_REL_TOL*max(1,|x|)."""
    g = np.load(REF_GOLD)
    scn = demo if tag.startswith("demo") else scenes.synthetic_scene(24, 24, 8, palette=64)
    st = make_stage(scn)
    n = st.build_requests()
    req = g[f"{tag}_requests"]
    assert n == len(req) and np.array_equal(st.download_requests(), req)
    size = int(g[f"{tag}_size"])
    for variant in (1, 0):
        st.set_kernel_variant(variant)
        for frame in (int(f) for f in g[f"{tag}_frames"]):
            st.upload_field(g[f"{tag}_before_{frame}"])
            st.set_debug(DEBUG_PRE_BLEND)
            st.update(frame, size, frame % size)
            pre = st.download_debug(n // size)["pre_blend"]
            want = g[f"{tag}_pre_{frame}"]
            assert np.all(np.abs(pre - want) <= REF_REL_TOL * np.maximum(1.0, np.abs(want))), (tag, frame)
            assert np.array_equal(st.download_field(), g[f"{tag}_after_{frame}"]), (tag, frame)
            if frame == 0:      # zero field: nothing is left to tolerance
                assert np.array_equal(pre, want), (tag, frame)
    st.set_kernel_variant(1)


@pytest.mark.parametrize("seed", [0, 2, 77, 2 ** 31 - 1, -5])
def test_demo_seeds_and_random_field(oracle_mod, demo, seed):
    O = oracle_mod
    st = make_stage(demo)
    sc = oracle_scene(O, demo)
    rng = np.random.default_rng(seed & 0xFFFF)
    f0 = rng.integers(0, 2 ** 32, sc.cells, dtype=np.uint64).astype(np.uint32)
    req = O.build_requests(demo["world"])
    assert_update_equal(st, O, sc, f0, req, seed)


def test_demo_warm_multiframe(oracle_mod, demo):
    """config 3 shape (static part): 16 accumulating frames, field fed back each frame."""
    O = oracle_mod
    st = make_stage(demo)
    sc = oracle_scene(O, demo)
    req = O.build_requests(demo["world"])
    st.set_requests(req)
    f = np.zeros(sc.cells, np.uint32)
    for frame in range(16):
        st.update(frame)
        f = O.update(sc, f, req, frame)["field"]
    assert np.array_equal(st.download_field(), f)


@pytest.mark.parametrize("size,shift", [(2, 0), (2, 1), (3, 2), (10, 7)])
def test_magic_size_dispatch(oracle_mod, demo, size, shift):
    """magicSize sub-sampling (internal_render.cpp:1190-1206): N/size workgroups, request w*size+shift."""
    O = oracle_mod
    st = make_stage(demo)
    sc = oracle_scene(O, demo)
    req = O.build_requests(demo["world"])
    rng = np.random.default_rng(size * 16 + shift)
    f0 = rng.integers(0, 2 ** 32, sc.cells, dtype=np.uint64).astype(np.uint32)
    assert_update_equal(st, O, sc, f0, req, 9, size=size, shift=shift)


def test_specials_and_request_builder(oracle_mod, demo):
    O = oracle_mod
    st = make_stage(demo)
    sp = np.array([(x, y, z, 0) for x in range(5, 12) for y in range(6, 16) for z in (2, 1)], np.int16)
    n = st.build_requests(sp)
    ref = O.build_requests(demo["world"], sp)
    assert n == len(ref) == 2689 + 66
    assert np.array_equal(st.download_requests(), ref)
    assert st.build_requests() == 2689
    assert np.array_equal(st.download_requests(), ref[:2689])


@pytest.mark.parametrize("rays", [16, 32, 64, 128, 256])
def test_synthetic_small_all_ray_counts(oracle_mod, small_synth, rays):
    O = oracle_mod
    st = make_stage(small_synth, rays=rays)
    sc = oracle_scene(O, small_synth, rays=rays)
    assert np.array_equal(bits(st.direction_table()), bits(O.direction_table(rays)))
    req = O.build_requests(small_synth["world"])
    rng = np.random.default_rng(rays)
    f0 = rng.integers(0, 2 ** 32, sc.cells, dtype=np.uint64).astype(np.uint32)
    assert_update_equal(st, O, sc, f0, req, 4)


def test_synthetic_builder_matches_oracle(oracle_mod):
    sc = scenes.synthetic_scene(67, 33, 9, palette=16, seed=99, p_cell=0.02)
    st = make_stage(sc)
    n = st.build_requests()
    ref = oracle_mod.build_requests(sc["world"])
    assert n == len(ref)
    assert np.array_equal(st.download_requests(), ref)


def test_edge_cases(oracle_mod, demo):
    O = oracle_mod
    sc = oracle_scene(O, demo)
    st = make_stage(demo)
    f0 = np.full(sc.cells, 0xFFFFFFFF, np.uint32)
    # empty request list: field untouched
    st.upload_field(f0)
    st.set_requests(np.zeros((0, 4), np.int16))
    st.update(1)
    assert np.array_equal(st.download_field(), f0)
    # probes on every world corner / face (rays leave through low and high faces, REPEAT gather wraps)
    req = np.array([(x, y, z, 0) for x in (0, 47) for y in (0, 47) for z in (0, 15)] +
                   [(0, 20, 15, 0), (47, 20, 8, 0), (20, 0, 15, 0), (20, 47, 15, 0)], np.int16)
    assert_update_equal(st, O, sc, f0, req, 6)
    # a probe listed alone runs as workgroup 0 whatever its position
    assert_update_equal(st, O, sc, f0, req[5:6], 6)


def test_empty_world_all_sky(oracle_mod):
    O = oracle_mod
    scn = dict(world=np.zeros((4, 6, 5), np.int16), bricks=np.zeros((2, 16, 16, 16), np.uint8),
               materials=np.zeros((256, 6), np.float32))
    st = make_stage(scn)
    sc = oracle_scene(O, scn)
    assert st.build_requests() == 0
    req = np.array([(x, y, z, 0) for z in range(4) for y in range(6) for x in range(5)], np.int16)
    ref = assert_update_equal(st, O, sc, np.zeros(sc.cells, np.uint32), req, 1)
    assert ref["totals"][1] == 0            # no hits: every ray is sky


def test_solid_world_and_saturation(oracle_mod):
    """Every voxel solid and strongly emissive: immediate hits, texels saturate at 1023."""
    O = oracle_mod
    mats = np.zeros((256, 6), np.float32); mats[:, :3] = 1.0; mats[:, 4] = 18.0
    scn = dict(world=np.ones((3, 4, 5), np.int16), bricks=np.full((2, 16, 16, 16), 7, np.uint8), materials=mats)
    st = make_stage(scn, reactiveness=1.0)
    sc = oracle_scene(O, scn, reactiveness=1.0)
    req = O.build_requests(scn["world"])
    assert len(req) == 60
    ref = assert_update_equal(st, O, sc, np.zeros(sc.cells, np.uint32), req, 1)
    assert np.all(ref["field"] == 0xFFFFFFFF)


def test_boundary_shell_hits_after_leaving(oracle_mod):
    """The sample the reference takes at pos in (-1,0) aliases onto voxel row 0 of the edge block
    (SURVEY a6).  Edge bricks here are solid only in their outermost voxel rows, so the rays that leave
    through a low face hit exactly through that aliased sample (the fast march's "pending" lanes)."""
    O = oracle_mod
    rng = np.random.default_rng(5)
    shell = np.zeros((16, 16, 16), np.uint8)        # [x][y][z]
    shell[0, :, :] = 3; shell[:, 0, :] = 4; shell[:, :, 0] = 5
    mats = rng.random((256, 6)).astype(np.float32)
    bricks = np.stack([np.zeros_like(shell), shell, np.zeros_like(shell)])
    world = np.zeros((3, 4, 5), np.int16)            # [z][y][x]
    world[0, :, :] = 1; world[:, 0, :] = 1; world[:, :, 0] = 1
    world[2, 3, 4] = 2                               # a non-empty block whose brick is all empty voxels
    scn = dict(world=world, bricks=bricks, materials=mats)
    st = make_stage(scn)
    sc = oracle_scene(O, scn)
    req = np.array([(x, y, z, 0) for z in range(3) for y in range(4) for x in range(5)], np.int16)
    f0 = rng.integers(0, 2 ** 32, sc.cells, dtype=np.uint64).astype(np.uint32)
    for seed in (0, 3, 11):
        ref = assert_update_equal(st, O, sc, f0, req, seed)
    assert 0 < ref["totals"][1] < len(req) * 64


@pytest.mark.parametrize("case", ["ids_outside_palette", "brick0_loaded"])
def test_general_march_path(oracle_mod, small_synth, case):
    """The pooled kernel's fast march requires brick 0 empty and every world id inside the palette; both are
    detected (host flag / validate_world_kernel) and such scenes take the general march."""
    O = oracle_mod
    scn = {k: v.copy() for k, v in small_synth.items()}
    rng = np.random.default_rng(17)
    if case == "ids_outside_palette":
        flat = scn["world"].reshape(-1)
        idx = rng.choice(flat.size, 40, replace=False)
        flat[idx[:20]] = -7                          # negative and beyond-capacity ids: "non-empty, no voxels"
        flat[idx[20:]] = scn["bricks"].shape[0] + 5
    else:
        scn["bricks"][0] = (rng.random((16, 16, 16)) < 0.02) * 9
    st = make_stage(scn)
    sc = oracle_scene(O, scn)
    req = O.build_requests(scn["world"])
    if len(req) == 0 or case == "brick0_loaded":
        wz, wy, wx = scn["world"].shape
        req = np.array([(x, y, z, 0) for z in range(wz) for y in range(wy) for x in range(wx)], np.int16)
    f0 = rng.integers(0, 2 ** 32, sc.cells, dtype=np.uint64).astype(np.uint32)
    assert_update_equal(st, O, sc, f0, req, 9)
    # back to a valid world on the same stage: the flag is recomputed per upload
    st.upload_world(small_synth["world"]); st.upload_bricks(0, small_synth["bricks"])
    sc2 = oracle_scene(O, small_synth)
    assert_update_equal(st, O, sc2, f0, O.build_requests(small_synth["world"]), 10)


def test_ranges_and_slabs_compose(oracle_mod, demo):
    """Probe sharding: disjoint workgroup ranges / z-slabs reproduce the full update exactly."""
    O = oracle_mod
    sc = oracle_scene(O, demo)
    sp = np.array([(7, 9, 2, 0), (40, 40, 14, 0), (6, 8, 1, 0)], np.int16)   # unsorted tail
    req = O.build_requests(demo["world"], sp)
    rng = np.random.default_rng(21)
    f0 = rng.integers(0, 2 ** 32, sc.cells, dtype=np.uint64).astype(np.uint32)
    ref = O.update(sc, f0, req, 12)["field"]
    st = make_stage(demo)
    st.set_requests(req)
    # (a) workgroup ranges on ONE field are NOT Jacobi across calls, so compose on copies
    n = len(req)
    cuts = [0, 500, 501, 1700, n]
    out = f0.copy()
    for b, e in zip(cuts[:-1], cuts[1:]):
        st.upload_field(f0)
        st.update_range(12, 1, 0, b, e)
        got = st.download_field()
        ch = got != f0
        out[ch] = got[ch]
    assert np.array_equal(out, ref)
    # (b) z-slabs: each "rank" owns a contiguous z range of the field
    out = np.empty_like(f0).reshape(16, 48, 48)
    for r in range(4):
        st.upload_field(f0)
        st.update_slab(12, 1, 0, 4 * r, 4 * r + 4)
        out[4 * r:4 * r + 4] = st.download_field().reshape(16, 48, 48)[4 * r:4 * r + 4]
    assert np.array_equal(out.reshape(-1), ref)
    zs = req[:, 2]
    prefix = n if np.all(np.diff(zs) >= 0) else int(np.argmax(np.diff(zs) < 0)) + 1
    assert st.slab_range(1, 0, 0, 16) == (0, prefix)
    assert st.slab_range(1, 0, 4, 8) == (int(np.searchsorted(zs[:prefix], 4)), int(np.searchsorted(zs[:prefix], 8)))


@pytest.mark.parametrize("mode", ["contiguous", "interleaved"])
@pytest.mark.parametrize("nranks", [2, 3, 8])
def test_interleaved_shards_compose(oracle_mod, nranks, mode):
    """lumcu_update_shard: chunks of 3072 workgroups dealt round-robin to the ranks.  Each shard, run from the
    same starting field, must change exactly its own probes' texels to the oracle's values; together they make
    the full update.  The world is large enough for several chunks per rank and a ragged last chunk, and for
    both kernels."""
    O = oracle_mod
    scn = scenes.synthetic_scene(40, 36, 16, palette=48, seed=21)
    st = make_stage(scn)
    sc = oracle_scene(O, scn)
    n = st.build_requests(np.array([(1, 2, 3, 0), (30, 30, 15, 0)], np.int16))
    req = st.download_requests()
    assert n > 6 * 3072 and n % 3072 != 0
    rng = np.random.default_rng(nranks)
    f0 = rng.integers(0, 2 ** 32, sc.cells, dtype=np.uint64).astype(np.uint32)
    ref = O.update(sc, f0, req, 5)["field"]
    cell = req[:, 0].astype(np.int64) + 40 * (req[:, 1].astype(np.int64) + 36 * req[:, 2].astype(np.int64))
    from lum_b200 import distributed as D
    import lum_b200
    st.set_shard_mode(lum_b200.SHARD_INTERLEAVED if mode == "interleaved" else lum_b200.SHARD_CONTIGUOUS)
    owner = np.full(n, -1, np.int64)
    for r in range(nranks):
        owner[D.shard_workgroups(n, r, nranks, st.shard_chunk(1, nranks))] = r
    assert (owner >= 0).all()
    # >= 64 chunks per rank, from the built list's upper bound (23,040 cells + 2 specials), not the count asked for above
    assert st.shard_chunk(1, nranks) == (0 if mode == "contiguous" else {2: 96, 3: 96, 8: 48}[nranks])
    assert mode == "contiguous" or st.shard_chunk(1, nranks) == D.interleave_chunk(st.cells + 2, nranks)   # the Python mirror
    for variant in (1, 0):                     # pooled and per-workgroup kernel: both map launch index -> workgroup
        st.set_kernel_variant(variant)
        union = f0.copy()
        for r in range(nranks):
            st.upload_field(f0)
            st.update_shard(5, 1, 0, r, nranks)
            got = st.download_field()
            mine = np.zeros(sc.cells, bool); mine[cell[owner == r]] = True
            assert np.array_equal(got[mine], ref[mine]), (variant, r)
            assert np.array_equal(got[~mine], f0[~mine]), (variant, r)
            union[mine] = got[mine]
        assert np.array_equal(union, ref), variant


def test_shift_matches_oracle(oracle_mod, demo):
    O = oracle_mod
    st = make_stage(demo)
    rng = np.random.default_rng(2)
    f0 = rng.integers(0, 2 ** 32, 48 * 48 * 16, dtype=np.uint64).astype(np.uint32)
    for d in [(1, 0, 0), (-3, 2, 1), (0, 0, -15), (47, -47, 0), (48, 0, 0), (0, 0, 0)]:
        st.upload_field(f0)
        st.shift(*d)
        assert np.array_equal(st.download_field(), O.shift_field(f0, (48, 48, 16), *d)), d


def test_round_trip_properties_full_size():
    """Size-independent properties at BASELINE config-4 size (256x256x64, no oracle involved):
    determinism (same inputs -> same field), every requested texel gets alpha=3, untouched
    texels keep their words, slabs compose to the full update."""
    scn = scenes.synthetic_scene(256, 256, 64)
    st = make_stage(scn)
    n = st.build_requests()
    assert n > 4_000_000
    st.update(1)
    f1 = st.download_field()
    st.upload_field(np.zeros(st.cells, np.uint32))
    st.update(1)
    assert np.array_equal(st.download_field(), f1)
    req = st.download_requests()
    lin = req[:, 0].astype(np.int64) + 256 * (req[:, 1].astype(np.int64) + 256 * req[:, 2].astype(np.int64))
    assert np.all(f1[lin] >> 30 == 3)
    mask = np.ones(st.cells, bool); mask[lin] = False
    assert np.all(f1[mask] == 0)
    out = np.empty_like(f1).reshape(64, 256, 256)
    for r in range(2):
        st.upload_field(np.zeros(st.cells, np.uint32))
        st.update_slab(1, 1, 0, 32 * r, 32 * r + 32)
        out[32 * r:32 * r + 32] = st.download_field().reshape(64, 256, 256)[32 * r:32 * r + 32]
    assert np.array_equal(out.reshape(-1), f1)


def test_full_size_config4_matches_oracle(oracle_mod):
    """The benchmark's own workload, whole: all 4,191,908 probes x 64 rays of the synthetic 256x256x64 world
    (BASELINE configs[3]) on a random starting field -- packed field, per-probe pre-blend radiance and the S/H
    totals of the roofline numerator against the oracle, bit for bit (the oracle needs a few seconds on the box's
    host cores for the 268 M rays)."""
    O = oracle_mod
    scn = scenes.synthetic_scene(256, 256, 64)
    st = make_stage(scn)
    sc = oracle_scene(O, scn)
    n = st.build_requests()
    req = st.download_requests()
    assert n == len(req) == 4_191_908
    rng = np.random.default_rng(4)
    f0 = rng.integers(0, 2 ** 32, sc.cells, dtype=np.uint64).astype(np.uint32)
    st.upload_field(f0)
    st.set_debug(DEBUG_PRE_BLEND)
    st.update(7)
    dbg = st.download_debug(n)
    got = st.download_field()
    st.set_debug(0)
    ref = O.update(sc, f0, req, 7)
    assert dbg["totals"] == ref["totals"]
    assert np.array_equal(got, ref["field"])
    # A ray whose direction has an exactly zero component (fib + 0.1*jitter cancels: about one ray in 2^28, i.e. one
    # or two probes of this update) gets a NaN refined fraction from 0*inf in radiance.comp:257-285, hence NaN
    # radiance on BOTH sides (the texel then packs as 0, decree 7).  NaN payloads are not part of the contract.
    nan = np.isnan(ref["pre_blend"])
    assert np.array_equal(np.isnan(dbg["pre_blend"]), nan) and nan.sum() <= 3 * 4
    ok = ~nan & (ref["pre_blend"] != 0)
    assert np.array_equal(dbg["pre_blend"][~nan], ref["pre_blend"][~nan])
    assert np.array_equal(bits(dbg["pre_blend"])[ok], bits(ref["pre_blend"])[ok])
    # the non-instrumented production kernel gives the same field
    st.upload_field(f0)
    st.update(7)
    assert np.array_equal(st.download_field(), ref["field"])
    st.close()


@pytest.mark.parametrize("case", range(32))
def test_random_small_worlds(oracle_mod, case):
    """Fuzz: odd world sizes (down to a single block), sparse to solid bricks, palettes of 1..40 bricks, requests
    that include duplicates' neighbours, out-of-world probes and a ragged workgroup count, random fields and seeds --
    every kernel variant against the oracle, bit for bit."""
    O = oracle_mod
    rng = np.random.default_rng(1000 + case)
    wx, wy, wz = (int(v) for v in rng.integers(1, [9, 8, 5]))
    P = int(rng.integers(2, 41))
    bricks = (rng.random((P, 16, 16, 16)) < rng.choice([0.002, 0.05, 0.3, 0.9])) * rng.integers(1, 256, (P, 16, 16, 16))
    bricks = bricks.astype(np.uint8)
    if case % 3:
        bricks[0] = 0                                   # two thirds of the cases keep brick 0 empty (fast march)
    world = (rng.random((wz, wy, wx)) < rng.choice([0.1, 0.5, 1.0])) * rng.integers(1, P, (wz, wy, wx))
    world = world.astype(np.int16)
    if case % 8 == 7:
        world.reshape(-1)[0] = P + 3                    # an id outside the palette (general march via the device flag)
    mats = rng.random((256, 6)).astype(np.float32)
    mats[:, 4] *= rng.choice([0.0, 1.0, 18.0])
    scn = dict(world=world, bricks=bricks, materials=mats)
    st = make_stage(scn)
    sc = oracle_scene(O, scn)
    inside = np.array([(x, y, z, 0) for z in range(wz) for y in range(wy) for x in range(wx)], np.int16)
    outside = np.array([(-1, 0, 0, 0), (wx, 0, 0, 0), (0, -2, wz, 0), (wx + 3, wy + 1, -1, 0)], np.int16)
    req = np.concatenate([inside, outside])
    req = req[rng.permutation(len(req))][: max(1, int(rng.integers(1, len(req) + 1)))]
    f0 = rng.integers(0, 2 ** 32, sc.cells, dtype=np.uint64).astype(np.uint32)
    size = int(rng.integers(1, 4))
    size = min(size, len(req))
    assert_update_equal(st, O, sc, f0, req, int(rng.integers(-2 ** 31, 2 ** 31)), size=size, shift=int(rng.integers(0, size)))
    st.close()


def test_slab_update_is_jacobi_across_its_two_passes(oracle_mod, demo):
    """lumcu_update_slab traces the slab's sorted range and then the unsorted special tail: the tail's rays must see
    the field as it was BEFORE the frame even where they land on texels of the sorted range (the first
    pass used to commit before the second one traced).  Specials sit in air right above the ground probes of their own
    slab, the field is random, one slab covers the whole world -- so the tail's gathers hit texels the first pass
    rewrites."""
    O = oracle_mod
    sc = oracle_scene(O, demo)
    rule = O.build_requests(demo["world"])
    sel = np.zeros((16, 48, 48), bool)
    sel[rule[:, 2], rule[:, 1], rule[:, 0]] = True
    # unselected cells directly above a selected one: their probe rays start next to updated texels
    cand = np.argwhere(~sel[1:] & sel[:-1])[:40]
    sp = np.array([(x, y, z + 1, 0) for z, y, x in cand], np.int16)[::-1].copy()      # descending: an unsorted tail
    assert len(sp) >= 20
    req = O.build_requests(demo["world"], sp)
    assert len(req) == len(rule) + len(sp)
    f0 = np.random.default_rng(77).integers(0, 2 ** 32, sc.cells, dtype=np.uint64).astype(np.uint32)
    ref = O.update(sc, f0, req, 3)["field"]
    st = make_stage(demo)
    st.build_requests(sp)
    assert np.array_equal(st.download_requests(), req)
    st.upload_field(f0)
    st.update_slab(3, 1, 0, 0, 16)
    assert np.array_equal(st.download_field(), ref)
    # unbounded slab arguments mean the whole list (the predecessor of the first workgroup is z = -32768)
    assert st.slab_range(1, 0, -40000, 40000) == (0, len(rule))
    st.upload_field(f0)
    st.update_slab(3, 1, 0, -2 ** 31, 2 ** 31 - 1)
    assert np.array_equal(st.download_field(), ref)
    st.close()


@pytest.mark.parametrize("with_specials", [False, True], ids=["ahead", "in-place"])
@pytest.mark.parametrize("variant", [0, 1, 2])      # 2 = the library chooses, from the last count the host has seen
def test_async_build_and_device_side_count(oracle_mod, demo, variant, with_specials):
    """lumcu_build_requests(out_count = NULL): nothing waits for the GPU, the launches are sized for an upper bound of
    the request count (14x the real one on the demo world) and the kernels stop at the count they read on the device.
    Without specials the list is built AHEAD -- other list buffer, prep stream, beside the frame in flight, with the
    world's validation -- so the frames here alternate between two different worlds (uploaded from pinned memory
    without a wait, into alternating world buffers): a stale list, count, world or validation flag would show.  Five
    frames queued back to back, result read back through the asynchronous download + fence: equal to the oracle."""
    import torch
    O = oracle_mod
    w0 = np.ascontiguousarray(demo["world"], np.int16)
    w1 = w0.copy()
    w1[8:, :, :] = 0                      # the upper half cleared ...
    w1[9, 20:30, 20:30] = 3               # ... and a floating slab of stone: other requests, other count
    worlds = [w0, w1]
    sp = np.array([(7, 9, 2, 0), (40, 40, 14, 0)], np.int16) if with_specials else None
    f0 = np.random.default_rng(5).integers(0, 2 ** 32, w0.size, dtype=np.uint64).astype(np.uint32)
    ref, counts = f0, []
    for f in range(5):
        w = worlds[f & 1]
        sc = O.Scene(w, demo["bricks"], demo["materials"])
        req = O.build_requests(w, sp)
        counts.append(len(req))
        ref = O.update(sc, ref, req, 40 + f)["field"]
    assert counts[0] != counts[1]
    st = make_stage(demo)
    st.set_kernel_variant(variant)
    pins = [torch.empty(w0.size, dtype=torch.int16, pin_memory=True) for _ in range(2)]
    for p, w in zip(pins, worlds):
        p.numpy()[:] = w.reshape(-1)
    pin_f = torch.empty(w0.size, dtype=torch.int32, pin_memory=True)
    st.upload_field(f0)
    for f in range(5):
        st.upload_world_async(pins[f & 1].numpy())
        assert st.build_requests(sp, sync=False) is None
        st.update(40 + f)
    out = pin_f.numpy().view(np.uint32)
    st.download_field_async(out)
    st.record_fence(1)
    st.wait_fence(1)
    assert np.array_equal(out, ref)
    assert st.request_count() == counts[4]                     # the host learns the count when it asks
    assert np.array_equal(st.download_requests(), O.build_requests(worlds[0], sp))
    # a synchronous build after builds ahead (shares the builder's scratch), then ahead again
    assert st.build_requests(sp) == counts[0]
    st.upload_world_async(pins[1].numpy())
    st.build_requests(sync=False)
    st.update(99)
    sc1 = O.Scene(w1, demo["bricks"], demo["materials"])
    ref = O.update(sc1, ref, O.build_requests(w1), 99)["field"]
    assert np.array_equal(st.download_field(), ref)
    st.wait_fence(3)                                           # never recorded: returns at once
    st.close()
    del pins, pin_f


def test_brick_copy_list_must_not_alias(demo):
    import lum_b200
    st = make_stage(demo, max_requests=0)
    with pytest.raises(lum_b200.LumcuError):
        st.copy_bricks([1, 2], [9, 9])                         # one destination twice
    st.close()
    st = lum_b200.RadianceStage((48, 48, 16), 32)
    st.upload_world(demo["world"]); st.upload_bricks(0, demo["bricks"]); st.upload_materials(demo["materials"])
    with pytest.raises(lum_b200.LumcuError):
        st.copy_bricks([1, 20], [20, 21])                      # a destination that is also a source
    st.copy_bricks([1, 2], [20, 21])
    got = st.download_bricks(20, 2)
    assert np.array_equal(got, demo["bricks"][1:3])
    st.close()
