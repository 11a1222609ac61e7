"""The glossy pass (shaders/glossy.frag) on the CPU side: the C restatement
oracle.glossy (oracle/lum_oracle.c: lum_oracle_glossy) against the REFERENCE's own shader text --
golden vectors made by executing glossy.frag (tests/golden/make_ref_glossy_golden.py) and, where
oracle/_ref is built, live runs on further inputs.

Bar: coverage masks and per-fragment world-fetch counts EQUAL (integer work); fragment colour within
REL_TOL*max(1,|x|) -- the same single source of tolerance as for radiance.comp (the shim's textureLod
un-normalises the coordinate in f32; see tests/test_ref_glsl.py); bit-equal on a zero field."""
import os

import numpy as np
import pytest

from glossy_cases import CASES, make_case

GOLD = os.path.join(os.path.dirname(__file__), "golden", "ref_glossy_golden.npz")
REL_TOL = 1e-6


def close(a, b):
    return np.all(np.abs(a - b) <= REL_TOL * np.maximum(1.0, np.abs(b)))


@pytest.mark.parametrize("tag", CASES)
def test_oracle_matches_reference_shader_golden(oracle_mod, tag):
    O = oracle_mod
    g = np.load(GOLD)
    c = make_case(O, tag)
    r = O.glossy(c["scene"], c["field"], c["view"], c["mat_norm"], c["depth"])
    assert np.array_equal(r["covered"], g[f"{tag}_covered"])
    assert r["covered"].any()
    assert close(r["frame"], g[f"{tag}_frame"])
    # texelFetch(blocks) per invocation: 2 dead at the origin (glossy.frag:131-132), 2 per coarse
    # iteration (:149-150), 1 per refinement sample (:203,216)
    cov = r["covered"] > 0
    mine = np.where(cov, 2 + r["coarse"].astype(np.int64) + r["samples"].astype(np.int64), 0)
    assert np.array_equal(mine, g[f"{tag}_fetches"].astype(np.int64))
    if tag == "demo":      # zero field: nothing is left to tolerance
        assert np.array_equal(r["frame"].view(np.uint32), g[f"{tag}_frame"].view(np.uint32))


def test_uncovered_fragments_keep_the_frame(oracle_mod):
    O = oracle_mod
    c = make_case(O, "demo")
    f0 = np.full((c["height"], c["width"], 4), 7.25, np.float32)
    r = O.glossy(c["scene"], c["field"], c["view"], c["mat_norm"], c["depth"], frame=f0)
    cov = r["covered"] > 0
    assert 0 < cov.sum() < cov.size
    assert np.all(r["frame"][~cov] == 7.25)
    assert np.all(r["frame"][cov][:, 3] <= 1.0)


needs_ref = pytest.mark.skipif(not __import__("oracle").ref_glossy_available(),
                               reason="oracle/_ref not built (needs /root/reference)")


@needs_ref
@pytest.mark.parametrize("seed", [0, 1, 2])
def test_oracle_matches_live_reference_shader_random_gbuffer(oracle_mod, seed):
    """G-buffers that no rasteriser would produce: origins inside solids, outside the world, arbitrary
    normals (so the mirror direction is arbitrary too, including components of either sign)."""
    O = oracle_mod
    c = make_case(O, "synth")
    rng = np.random.default_rng(seed)
    h, w = 40, 56
    mn = rng.integers(0, 256, (h, w, 4), dtype=np.uint8)
    dp = rng.random((h, w), dtype=np.float32) * np.float32(0.5)
    view = O.camera_view((rng.random(3) * 300 - 50), (0.3 + seed, 0.9, -0.6), w, h, pixels_in_voxel=0.2)
    a = O.glossy(c["scene"], c["field"], view, mn, dp)
    b = O.ref_glsl_glossy(c["scene"], c["field"], view, mn, dp)
    assert np.array_equal(a["covered"], b["covered"])
    cov = a["covered"] > 0
    assert close(a["frame"], b["frame"])
    assert np.array_equal(np.where(cov, 2 + a["coarse"].astype(np.int64) + a["samples"], 0), b["fetches"])


@needs_ref
def test_glossy_fuzz_against_reference_shader():
    """tools/ref_glossy_fuzz.py: random worlds, cameras, fields and G-buffers through the reference's glossy.frag text and
    through oracle.glossy (round 2 ran 400,000 cases, 38 M mirror rays: coverage and fetch counts equal, colour within
    4.3e-7); a slice of it here."""
    import os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    p = subprocess.run([sys.executable, os.path.join(root, "tools", "ref_glossy_fuzz.py"), "0", "150"], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0 and "fetch counts equal" in p.stdout, p.stdout[-500:] + p.stderr[-2000:]


@needs_ref
def test_marcher_of_glossy_is_the_marcher_of_radiance():
    """CastRay_fast is duplicated in the reference (radiance.comp:183-291, glossy.frag:119-226); the oracle
    and the CUDA stage use ONE implementation for both, which is only right while the two texts agree."""
    import re
    root = "/root/reference/shaders/"

    def body(path):
        src = open(root + path, encoding="utf-8", errors="replace").read().replace("\r\n", "\n")
        src = re.sub(r"//[^\n]*", "", src)                 # comments (some hold unbalanced braces)
        src = re.sub(r"\[\[\w+\]\]", "", src)              # loop attributes
        i = src.index("bool CastRay_fast(")
        depth, j = 0, src.index("{", i)
        k = j
        while True:
            depth += (src[k] == "{") - (src[k] == "}")
            k += 1
            if depth == 0:
                break
        return re.sub(r"\s+", "", src[j:k])
    assert body("radiance.comp") == body("glossy.frag")
