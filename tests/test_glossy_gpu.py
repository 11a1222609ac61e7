"""GPU parity of the glossy pass: lumcu_glossy through the C-ABI against the CPU oracle
(bit for bit: the same f32 operations in the same order on both sides) and against the golden vectors
produced by the reference's own glossy.frag (tests/golden/ref_glossy_golden.npz; 1e-6 relative, see
tests/test_glossy.py for the one source of that tolerance)."""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import make_stage
from glossy_cases import CASES, make_case

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(__file__), "golden", "ref_glossy_golden.npz")
REL_TOL = 1e-6


def assert_bits_equal(got, ref):
    """Bit for bit, except that -0.0 == +0.0 is allowed."""
    assert np.array_equal(got, ref)
    nz = ref != 0
    assert np.array_equal(got.view(np.uint32)[nz], ref.view(np.uint32)[nz])


@pytest.mark.parametrize("tag", CASES)
def test_glossy_matches_oracle_and_reference_golden(oracle_mod, tag):
    O = oracle_mod
    c = make_case(O, tag)
    st = make_stage(c["scn"])
    st.upload_field(c["field"])
    f0 = np.full((c["height"], c["width"], 4), -3.5, np.float32)
    got, cov = st.glossy(c["view"], c["mat_norm"], c["depth"], frame=f0)
    ref = O.glossy(c["scene"], c["field"], c["view"], c["mat_norm"], c["depth"], frame=f0)
    assert np.array_equal(cov, ref["covered"])
    assert_bits_equal(got, ref["frame"])
    g = np.load(GOLD)
    m = cov > 0
    assert np.array_equal(cov, g[f"{tag}_covered"])
    assert np.all(np.abs(got[m] - g[f"{tag}_frame"][m]) <= REL_TOL * np.maximum(1.0, np.abs(g[f"{tag}_frame"][m])))
    assert np.all(got[~m] == -3.5)          # fragments the stencil rule rejects keep the frame
    st.close()


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_glossy_random_gbuffers(oracle_mod, seed):
    """Arbitrary G-buffers: origins inside solids and outside the world, arbitrary (unnormalised) normals,
    odd frame sizes (partial tiles).  seed 3 uses a scene with a loaded brick 0 (general march)."""
    O = oracle_mod
    c = make_case(O, "synth")
    scn = {k: v.copy() for k, v in c["scn"].items()}
    rng = np.random.default_rng(seed)
    if seed == 3:
        scn["bricks"][0] = (rng.random((16, 16, 16)) < 0.01) * 5
    sc = O.Scene(scn["world"], scn["bricks"], scn["materials"])
    st = make_stage(scn)
    st.upload_field(c["field"])
    h, w = 37 + seed, 61 - seed
    mn = rng.integers(0, 256, (h, w, 4), dtype=np.uint8)
    dp = rng.random((h, w), dtype=np.float32) * np.float32(0.6)
    view = O.camera_view((rng.random(3) * 300 - 50), (0.3 + seed, 0.9 - seed, -0.6), w, h, pixels_in_voxel=0.2)
    got, cov = st.glossy(view, mn, dp)
    ref = O.glossy(sc, c["field"], view, mn, dp)
    assert np.array_equal(cov, ref["covered"])
    assert cov.any()
    assert_bits_equal(got, ref["frame"])
    st.close()


def test_glossy_sees_the_updated_field(oracle_mod):
    """update -> glossy on the same stream: the pass reads the committed field of the frame."""
    O = oracle_mod
    c = make_case(O, "demo_all")
    st = make_stage(c["scn"])
    st.upload_field(c["field"])
    n = st.build_requests()
    st.update(3)
    got, _ = st.glossy(c["view"], c["mat_norm"], c["depth"])
    req = O.build_requests(c["scn"]["world"])
    assert n == len(req)
    f1 = O.update(c["scene"], c["field"], req, 3)["field"]
    assert np.array_equal(st.download_field(), f1)
    ref = O.glossy(c["scene"], f1, c["view"], c["mat_norm"], c["depth"])
    assert_bits_equal(got, ref["frame"])
    st.close()


def test_glossy_device_entry_and_errors(oracle_mod):
    """lumcu_glossy_device with buffers that already live on the GPU (torch tensors), and argument checks."""
    import torch
    O = oracle_mod
    c = make_case(O, "synth")
    st = make_stage(c["scn"])
    st.upload_field(c["field"])
    h, w = c["height"], c["width"]
    dev = torch.device("cuda:0")
    mn = torch.from_numpy(c["mat_norm"]).to(dev)
    dp = torch.from_numpy(c["depth"]).to(dev)
    fr = torch.zeros((h, w, 4), dtype=torch.float32, device=dev)
    cv = torch.zeros((h, w), dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()
    v = st.glossy_view(c["view"])
    rc = st.lib.lumcu_glossy_device(st.ctx, C.byref(v), w, h, mn.data_ptr(), dp.data_ptr(), fr.data_ptr(), cv.data_ptr())
    assert rc == 0, st.lib.lumcu_last_error()
    st.wait_idle()
    ref = O.glossy(c["scene"], c["field"], c["view"], c["mat_norm"], c["depth"])
    assert np.array_equal(cv.cpu().numpy(), ref["covered"])
    assert_bits_equal(fr.cpu().numpy(), ref["frame"])
    assert st.lib.lumcu_glossy_device(st.ctx, C.byref(v), 0, h, mn.data_ptr(), dp.data_ptr(), fr.data_ptr(), 0) == -1
    assert st.lib.lumcu_glossy_device(st.ctx, C.byref(v), w, h, 0, dp.data_ptr(), fr.data_ptr(), 0) == -1
    assert st.lib.lumcu_glossy(st.ctx, None, w, h, None, None, None, None) == -1
    st.close()
