"""The oracle's HOST-side functions pinned against the REFERENCE's own C++.

oracle/_ref/libref_host.so is LumInternalRenderer::update_radiance's list rule (internal_render.cpp:1132-1169),
get_shift (:989-1013), start_blockify / blockify_mesh (:977-987, :1029-1076), map_mesh's arithmetic (:1501-1510)
and table3d (internal_render.hpp:121-198), cut out of the reference tree verbatim and compiled against its vendored
glm (oracle/Makefile, oracle/host_prep.py, oracle/ref_host.cpp).  tests/golden/ref_host_golden.npz holds its outputs
(tests/golden/make_ref_host_golden.py).  Here oracle.build_requests / place_model / blockify -- what the CUDA
request builder, the facade's blockify and lumcu_map_model's arguments are compared with bit for bit -- are checked
against those vectors everywhere, and against live runs of the reference code where the tree is present.
Bar: equal.  Integers, and floats that both sides compute with the same operations in the same order."""
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(__file__), "golden", "ref_host_golden.npz")
PFIELDS_F = ("aabb_min", "aabb_max")
PFIELDS_I = ("block_min", "block_max", "map_shift", "map_area")


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def placement_dict(p):
    d = {f: np.array(list(getattr(p, f)), np.float32) for f in PFIELDS_F}
    d.update({f: np.array(list(getattr(p, f)), np.int32) for f in PFIELDS_I})
    d["inverse"] = np.array(list(p.inverse_trans), np.float32)
    return d


def test_request_list_matches_reference_golden(oracle_mod, demo):
    O = oracle_mod
    g = np.load(GOLD)
    assert len(g["demo_requests"]) == 2689                                     # SURVEY 8c KAT
    assert np.array_equal(O.build_requests(demo["world"]), g["demo_requests"])
    got = O.build_requests(demo["world"], g["demo_specials"])
    assert np.array_equal(got, g["demo_requests_sp"])
    assert len(got) == 2689 + 16          # of the demo's 66 water specials (+ repeats) only the unselected ones are appended, repeats twice
    assert np.array_equal(O.build_requests(g["synth_world"], g["synth_specials"]), g["synth_requests"])


def test_placement_matches_reference_golden(oracle_mod):
    O = oracle_mod
    g = np.load(GOLD)
    for i in range(int(g["n_place"])):
        v = g[f"place_{i}_in"]
        p = placement_dict(O.place_model(v[0:4], v[4:7], v[7:10].astype(int), (48, 48, 16)))
        for f in PFIELDS_F:
            assert np.array_equal(bits(p[f]), bits(g[f"place_{i}_{f}"])), (i, f)
        for f in ("map_shift", "map_area"):
            assert np.array_equal(p[f], g[f"place_{i}_{f}"]), (i, f)
        assert np.array_equal(bits(p["inverse"]), bits(g[f"place_{i}_inverse"])), i
        # the reference's block range shows as the extent of the cells it allocated on an empty world; a model whose
        # clamped border is one cell thick at a world edge still allocates that edge cell, as the oracle's range says
        assert np.array_equal(p["block_min"], g[f"place_{i}_block_min"]), i
        assert np.array_equal(p["block_max"], g[f"place_{i}_block_max"]), i


def test_blockify_matches_reference_golden(oracle_mod, demo):
    O = oracle_mod
    g = np.load(GOLD)
    for i in range(int(g["n_blockify"])):
        ms = g[f"blockify_{i}_models"]
        pls = [O.place_model(m[0:4], m[4:7], m[7:10].astype(int), (48, 48, 16)) for m in ms]
        cur, src, dst = O.blockify(demo["world"], 15, pls)
        assert np.array_equal(cur, g[f"blockify_{i}_world"]), i
        assert np.array_equal(src, g[f"blockify_{i}_src"]) and np.array_equal(dst, g[f"blockify_{i}_dst"]), i
        assert int(g[f"blockify_{i}_queued"]) == int(np.count_nonzero(src))       # the reference queues no copy for empty sources (:1051)


# ---- live against the reference code (build container only) ------------------------------------------------------
def _live(O):
    O.build()
    if not O.ref_host_available():
        pytest.skip("oracle/_ref/libref_host.so not built (no reference tree here)")


@pytest.mark.parametrize("seed", range(6))
def test_request_list_live(oracle_mod, seed):
    O = oracle_mod
    _live(O)
    rng = np.random.default_rng(seed)
    wx, wy, wz = (int(v) for v in rng.integers(1, 40, 3))
    dens = [0.0, 0.01, 0.1, 0.6, 0.003, 1.0][seed]
    w = np.where(rng.random((wz, wy, wx)) < dens, rng.integers(-3, 40, (wz, wy, wx)), 0).astype(np.int16)   # negative ids lower the sum (:1153)
    nsp = int(rng.integers(0, 80))
    sp = np.stack([rng.integers(0, wx, nsp), rng.integers(0, wy, nsp), rng.integers(0, wz, nsp), rng.integers(0, 3, nsp)], 1).astype(np.int16)
    assert np.array_equal(O.build_requests(w, sp if nsp else None), O.ref_host_build_requests(w, sp if nsp else None))


@pytest.mark.parametrize("seed", range(8))
def test_placement_and_blockify_live(oracle_mod, demo, seed):
    O = oracle_mod
    _live(O)
    rng = np.random.default_rng(100 + seed)
    models = []
    for k in range(int(rng.integers(1, 5))):
        ax = rng.normal(size=3); ax /= np.linalg.norm(ax)
        a = rng.uniform(-3.2, 3.2)
        rot = (np.float32(np.cos(a / 2)), *(np.float32(np.sin(a / 2)) * ax.astype(np.float32)))
        size = [int(v) for v in rng.integers(1, 64, 3)]
        shift = [float(np.float32(v)) for v in rng.uniform(-60, 16 * 48 + 40, 3)]
        models.append((rot, shift, size))
    pls = []
    for rot, shift, size in models:
        a, b = placement_dict(O.place_model(rot, shift, size, (48, 48, 16))), placement_dict(O.ref_host_place_model(rot, shift, size, (48, 48, 16)))
        for f in PFIELDS_F + ("inverse",):
            assert np.array_equal(bits(a[f]), bits(b[f])), f
        for f in PFIELDS_I:
            assert np.array_equal(a[f], b[f]), f
        pls.append(O.place_model(rot, shift, size, (48, 48, 16)))
    cur, src, dst = O.blockify(demo["world"], 15, pls)
    rcur, rsrc, rdst, queued = O.ref_host_blockify(demo["world"], 15, models)
    assert np.array_equal(cur, rcur) and np.array_equal(src, rsrc) and np.array_equal(dst, rdst)
    assert queued == int(np.count_nonzero(src))


def test_host_side_fuzz_against_reference_cpp(oracle_mod):
    """tools/ref_host_fuzz.py: the request rule, placement, blockify and map.comp over random inputs against the reference's
    own C++ / shader text (round 2 ran 80,000 cases: 336 M requests, 200 K placements, 852 M mapped voxels, all equal);
    a slice of it here."""
    import subprocess, sys
    O = oracle_mod
    _live(O)
    if not O.ref_map_available():
        pytest.skip("oracle/_ref/libref_map_glsl.so not built (no reference tree here)")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    p = subprocess.run([sys.executable, os.path.join(root, "tools", "ref_host_fuzz.py"), "0", "25"], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0 and "all equal" in p.stdout, p.stdout[-500:] + p.stderr[-2000:]

