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


def test_glossy_through_the_lum_api_mirror(oracle_mod, demo):
    """lum_b200_glossy (include/clum.h): the pass driven from the Lum API mirror, uniform block filled from the
    renderer's camera the way start_raygen does (internal_render.cpp:1580-1592), after three submitted frames."""
    O = oracle_mod
    from lum_b200 import clum
    mats = demo["materials"].copy()
    mats[:, 5] *= 0.5                                     # every material glossy
    r = clum.Renderer((48, 48, 16), 15)
    r.load_world(demo["world"])
    for i in range(1, 15):
        r.load_block(i, demo["bricks"][i])
    r.set_material_palette(mats)
    r.upload_block_palette(); r.upload_material_palette()
    r.set_radiance_controller(False, 1)
    r.set_camera((60.0, 0.0, 194.0), (0.61, 1.0, -0.8))   # not normalised on purpose: the reference stores it as given
    for _ in range(3):
        r.start_frame(); r.prepare_frame(); r.submit_frame()
    field = r.download_field()
    w, h = 1920 // 16, 1080 // 16                         # the mirror's camera has the reference's 1920x1080 / 5 view
    d = np.array([0.61, 1.0, -0.8])
    plane = np.array([d[0], d[1], 0.0]); plane /= np.linalg.norm(plane)
    hz = np.cross(plane, [0, 0, 1.0]); hz /= np.linalg.norm(hz)
    vt = np.cross(d, hz); vt /= np.linalg.norm(vt)
    view = dict(campos=np.array([60.0, 0.0, 194.0], np.float32), camdir=d.astype(np.float32),
                horizline_scaled=(hz * (1920 / 5.0) / 2.0).astype(np.float32),
                vertiline_scaled=(vt * (1080 / 5.0) / 2.0).astype(np.float32),
                global_light_dir=(np.array([0.5, 0.5, -0.9]) / np.linalg.norm([0.5, 0.5, -0.9])).astype(np.float32),
                frame_size=np.array([w, h], np.float32))
    sc = O.Scene(demo["world"], np.concatenate([demo["bricks"], np.zeros((256 - 15, 16, 16, 16), np.uint8)]), mats)
    mn, dp = O.gbuffer(sc, view, w, h)
    got, cov = r.glossy(mn, dp)
    ref = O.glossy(sc, field, view, mn, dp)
    assert cov.any() and np.array_equal(cov, ref["covered"])
    assert_bits_equal(got, ref["frame"])
    r.close()
