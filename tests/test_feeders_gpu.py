"""GPU parity tests of the feeders and of the Lum API mirror: the CUDA
path called through the C-ABI (lumcu_copy_bricks / lumcu_map_model) and through Lum's own C API
(include/clum.h), against the CPU oracle on the same inputs.  Bar: bit-exact bricks, world ids and
radiance-field words."""
import math

import numpy as np
import pytest

from conftest import make_stage
from voxutil import write_scene, write_vox

pytestmark = pytest.mark.gpu

WS = (48, 48, 16)
STATIC = 15


def quat_yaw(a):
    return (np.float32(math.cos(a / 2)), np.float32(0), np.float32(0), np.float32(math.sin(a / 2)))


def quat_axis(axis, a):
    ax = np.asarray(axis, np.float64); ax = ax / np.linalg.norm(ax)
    s = math.sin(a / 2)
    return tuple(np.float32(v) for v in (math.cos(a / 2), ax[0] * s, ax[1] * s, ax[2] * s))


# ---- C-ABI level ---------------------------------------------------------------------------
def test_copy_bricks_matches_oracle(oracle_mod, demo):
    O = oracle_mod
    st = make_stage(dict(world=demo["world"], bricks=np.concatenate([demo["bricks"], np.zeros((49, 16, 16, 16), np.uint8)]),
                         materials=demo["materials"]))
    rng = np.random.default_rng(0)
    src = rng.integers(0, 15, 40).astype(np.int32)
    dst = np.arange(20, 60, dtype=np.int32)
    junk = rng.integers(0, 256, (49, 16, 16, 16)).astype(np.uint8)          # stale dynamic bricks
    st.upload_bricks(15, junk)
    ref = np.concatenate([demo["bricks"], junk]).copy()
    O.copy_bricks(ref, src, dst)
    st.copy_bricks(src, dst)
    assert np.array_equal(st.download_bricks(0, 64), ref)
    assert (ref[dst[src == 0]] == 0).all()                                   # src 0 zeroes the brick


PLACEMENTS = [
    ("identity", (1, 0, 0, 0), (13.1 * 16, 14.1 * 16, 3.1 * 16)),
    ("yaw", quat_yaw(0.7), (300.0, 310.0, 40.0)),
    ("tilted", quat_axis((1, 2, 3), 1.1), (400.5, 200.25, 100.75)),
    ("low corner, partly outside", quat_yaw(-2.1), (12.0, 17.0, 30.0)),
    ("high corner, partly outside", quat_axis((0, 1, 0), 0.3), (750.0, 755.0, 245.0)),
    ("fully outside", (1, 0, 0, 0), (-400.0, -400.0, -400.0)),
]


@pytest.mark.parametrize("name,rot,shift", PLACEMENTS, ids=[p[0] for p in PLACEMENTS])
def test_map_model_matches_oracle(oracle_mod, demo, name, rot, shift):
    O = oracle_mod
    for mname in ("model_tank_body", "model_robot"):
        model = demo[mname]
        sz, sy, sx = model.shape
        p = O.place_model(rot, shift, (sx, sy, sz), WS)
        cur, src, dst = O.blockify(demo["world"], STATIC, [p])
        bricks = np.zeros((512, 16, 16, 16), np.uint8); bricks[:15] = demo["bricks"]
        st = make_stage(dict(world=cur, bricks=bricks, materials=demo["materials"]))
        O.copy_bricks(bricks, src, dst)
        O.map_model(cur, bricks, model, p, STATIC)
        h = st.upload_model(model)
        st.copy_bricks(src, dst)
        st.map_model(h, np.array(p.inverse_trans, np.float32), list(p.map_shift), list(p.map_area), STATIC)
        assert np.array_equal(st.download_bricks(0, 512), bricks), mname
        assert np.array_equal(st.download_world(), cur)
        st.free_model(h)
        st.close()


def test_map_two_overlapping_models_later_wins(oracle_mod, demo):
    O = oracle_mod
    a, b = demo["model_tank_body"], demo["model_tank_head"]
    pa = O.place_model((1, 0, 0, 0), (209.6, 225.6, 49.6), a.shape[::-1], WS)
    pb = O.place_model(quat_yaw(0.5), (212.0, 228.0, 52.0), b.shape[::-1], WS)
    cur, src, dst = O.blockify(demo["world"], STATIC, [pa, pb])
    bricks = np.zeros((256, 16, 16, 16), np.uint8); bricks[:15] = demo["bricks"]
    st = make_stage(dict(world=cur, bricks=bricks, materials=demo["materials"]))
    O.copy_bricks(bricks, src, dst)
    O.map_model(cur, bricks, a, pa, STATIC)
    O.map_model(cur, bricks, b, pb, STATIC)
    ha, hb = st.upload_model(a), st.upload_model(b)
    st.copy_bricks(src, dst)
    st.map_model(ha, np.array(pa.inverse_trans, np.float32), list(pa.map_shift), list(pa.map_area), STATIC)
    st.map_model(hb, np.array(pb.inverse_trans, np.float32), list(pb.map_shift), list(pb.map_area), STATIC)
    assert np.array_equal(st.download_bricks(0, 256), bricks)


def test_radiance_update_sees_mapped_model(oracle_mod, demo):
    """The hot path downstream of the feeders: update after blockify + copy + map, CUDA vs oracle."""
    O = oracle_mod
    model = demo["model_tank"]
    p = O.place_model(quat_yaw(0.3), (150.0, 300.0, 80.0), model.shape[::-1], WS)
    cur, src, dst = O.blockify(demo["world"], STATIC, [p])
    bricks = np.zeros((256, 16, 16, 16), np.uint8); bricks[:15] = demo["bricks"]
    st = make_stage(dict(world=cur, bricks=bricks, materials=demo["materials"]))
    O.copy_bricks(bricks, src, dst)
    O.map_model(cur, bricks, model, p, STATIC)
    h = st.upload_model(model)
    st.copy_bricks(src, dst)
    st.map_model(h, np.array(p.inverse_trans, np.float32), list(p.map_shift), list(p.map_area), STATIC)
    n = st.build_requests()
    req = O.build_requests(cur)
    assert n == len(req) and n > 2689                                        # the tank adds probes
    assert np.array_equal(st.download_requests(), req)
    sc = O.Scene(cur, bricks, demo["materials"])
    ref = O.update(sc, np.zeros(sc.cells, np.uint32), req, 5)
    st.update(5)
    assert np.array_equal(st.download_field(), ref["field"])


# ---- Lum C API level (include/clum.h) ----------------------------------------------------------
def camera_matrix(pos=(60, 0, 194), direction=(0.61, 1.0, -0.8)):
    """Camera::updateCamera (internal_render.cpp:923-934) in float64."""
    eye = np.array(pos, np.float64)
    f = np.array(direction, np.float64); f /= np.linalg.norm(f)
    s = np.cross(f, [0, 0, 1.0]); s /= np.linalg.norm(s)
    u = np.cross(s, f)
    view = np.eye(4)
    view[0, :3], view[1, :3], view[2, :3] = s, u, -f
    view[0, 3], view[1, 3], view[2, 3] = -s @ eye, -u @ eye, f @ eye
    w, h = 1920 / 5.0, 1080 / 5.0
    l, r, b, t, n, fa = -w / 2, w / 2, h / 2, -h / 2, -0.0, 2000.0
    proj = np.diag([2 / (r - l), 2 / (t - b), -2 / (fa - n), 1.0])
    proj[0, 3], proj[1, 3], proj[2, 3] = -(r + l) / (r - l), -(t + b) / (t - b), -(fa + n) / (fa - n)
    return proj @ view


def draw_order(shifts, cam):
    """prepareFrame (renderer.cpp:395-409): far-to-near by -(cam * shift).z, stable."""
    d = [np.float32(-(cam @ np.array([*np.float32(s).astype(np.float64), 1.0]))[2]) for s in shifts]
    return sorted(range(len(shifts)), key=lambda i: -d[i])


def make_renderer(demo, controller=False, magic=1):
    from lum_b200 import clum
    r = clum.Renderer(WS, STATIC)
    r.load_world(demo["world"])
    for i in range(1, 15):
        r.load_block(i, demo["bricks"][i])
    r.set_material_palette(demo["materials"])
    r.upload_block_palette()
    r.upload_material_palette()
    r.set_radiance_controller(controller, magic)
    return r


def oracle_frame(O, demo, field, frame, models, specials, size=1):
    """One frame of Lum::Renderer::submitFrame with the snapshot semantics of DESIGN.md decree 9:
    blockify -> brick copies -> map (draw order) -> request list -> update."""
    pls = [O.place_model(rot, shift, m.shape[::-1], WS) for m, rot, shift in models]
    cur, src, dst = O.blockify(demo["world"], STATIC, pls)
    bricks = np.zeros((256, 16, 16, 16), np.uint8); bricks[:15] = demo["bricks"]
    O.copy_bricks(bricks, src, dst)
    for (m, _, _), p in zip(models, pls):
        O.map_model(cur, bricks, m, p, STATIC)
    req = O.build_requests(cur, specials)
    sc = O.Scene(cur, bricks, demo["materials"])
    out = O.update(sc, field, req, frame, size=size, shift=frame % size)
    return out["field"], cur, bricks, len(req)


def test_api_static_frames_config1(oracle_mod, demo):
    """BASELINE config 1/2 through Lum's own API: demo world, every probe every frame."""
    O = oracle_mod
    r = make_renderer(demo)
    field = np.zeros(r.cells, np.uint32)
    for f in range(4):
        r.start_frame(); r.draw_world(); r.prepare_frame(); r.submit_frame()
        field, cur, _, n = oracle_frame(O, demo, field, f, [], None)
        assert r.request_count() == n == 2689
        assert r.frame_index() == f + 1
        assert np.array_equal(r.download_field(), field), f
    ts = r.timestamps()
    assert [t[0] for t in ts] == ["update_radiance", "update_radiance"] and ts[1][1] > ts[0][1]
    r.close()


def test_api_special_updates_and_world_edits(oracle_mod, demo):
    O = oracle_mod
    r = make_renderer(demo)
    d = dict(demo); d["world"] = demo["world"].copy()
    field = np.zeros(r.cells, np.uint32)
    sp = [(xx, yy, zz, 0) for xx in range(5, 12) for yy in range(6, 16) for zz in (2, 1)]   # examples/demo.cpp:208-214
    for f in range(3):
        if f == 1:
            r.set_world_block(40, 40, 12, 8); d["world"][12, 40, 40] = 8     # a lamp in the air
            assert r.get_world_block(40, 40, 12) == 8
        r.start_frame()
        for s in sp:
            r.request_radiance_update(*s[:3])
        r.prepare_frame(); r.submit_frame()
        field, cur, _, n = oracle_frame(O, d, field, f, [], np.array(sp, np.int16))
        assert r.request_count() == n
        assert np.array_equal(r.download_world(), cur)
        assert np.array_equal(r.download_field(), field), f
    r.close()


@pytest.mark.parametrize("pair", [("tank_body", "tank_head"), ("robot", "tank")])
def test_api_config3_sixteen_frames_moving_models(oracle_mod, demo, pair):
    """BASELINE config 3: 16-frame temporal accumulation with two mapped models moving through the
    demo world (SURVEY.md 8d "C3 inputs"), driven through lum_draw_model / lum_submit_frame."""
    O = oracle_mod
    r = make_renderer(demo)
    ma, mb = demo[f"model_{pair[0]}"], demo[f"model_{pair[1]}"]
    ha, hb = r.load_mesh_data(ma), r.load_mesh_data(mb)
    assert r.model_size(ha) == ma.shape[::-1]
    cam = camera_matrix()
    field = np.zeros(r.cells, np.uint32)
    for f in range(16):
        base = np.float32([13.1 * 16, 14.1 * 16, 3.1 * 16]) + np.float32(f) * np.float32([0.75, 0.5, 0.0])
        ta = (quat_yaw(0.05 * f), tuple(base))
        tb = (quat_yaw(-0.11 * f), tuple(base + np.float32([2.0, 3.0, ma.shape[0]])))
        r.start_frame(); r.draw_world()
        r.draw_model(ha, *ta); r.draw_model(hb, *tb)
        r.prepare_frame(); r.submit_frame()
        models = [(ma, *ta), (mb, *tb)]
        order = draw_order([ta[1], tb[1]], cam)
        field, cur, bricks, n = oracle_frame(O, demo, field, f, [models[i] for i in order], None)
        assert np.array_equal(r.download_world(), cur), f
        ndyn = int(cur.max()) + 1
        assert ndyn > STATIC
        assert np.array_equal(r.download_block_palette(0, ndyn), bricks[:ndyn]), f
        assert r.request_count() == n
        assert np.array_equal(r.download_field(), field), f
    # models gone: the world returns to the static one
    r.start_frame(); r.prepare_frame(); r.submit_frame()
    assert np.array_equal(r.download_world(), demo["world"])
    r.free_mesh(ha); r.free_mesh(hb)
    r.close()


def test_api_probe_update_controller(oracle_mod, demo):
    """internal_render.cpp:1190-1206: magicSize starts at 2, follows the dispatch time reported for the previous
    frame (injected here: the trajectory must not depend on how fast the GPU under test is -- under
    compute-sanitizer a demo update takes far longer than 2.2 ms), the dispatch covers N/magicSize requests
    starting at iFrame % magicSize."""
    O = oracle_mod
    from lum_b200 import clum
    r = make_renderer(demo, controller=True, magic=0)
    assert r.magic_size() == 2
    field = np.zeros(r.cells, np.uint32)
    # reported time per frame -> magicSize the NEXT update_radiance uses: < 1.8 ms decrements, > 2.2 ms increments,
    # clamped to [1, 10]; frame 0 sees timeTakenByRadiance = 0
    reported = [0.5, 0.5, 3.0, 3.0, 2.0, 2.0, 1.0]
    expect = [1, 1, 1, 2, 3, 3, 3]
    for f, (ms, size) in enumerate(zip(reported, expect)):
        r.set_radiance_time_override(ms)
        r.start_frame(); r.prepare_frame(); r.submit_frame()
        assert r.magic_size() == size, (f, r.magic_size())
        field, *_ = oracle_frame(O, demo, field, f, [], None, size=size)
        assert np.array_equal(r.download_field(), field), f
    r.set_radiance_time_override(50.0)
    for f in range(len(expect), len(expect) + 9):
        r.start_frame(); r.prepare_frame(); r.submit_frame()
    assert r.magic_size() == 10                                 # clamped (:1200-1201)
    r.close()
    # fixed magicSize 3: every third request, phase = frame % 3
    r = make_renderer(demo, controller=False, magic=3)
    field = np.zeros(r.cells, np.uint32)
    for f in range(4):
        r.start_frame(); r.prepare_frame(); r.submit_frame()
        field, *_ = oracle_frame(O, demo, field, f, [], None, size=3)
        assert np.array_equal(r.download_field(), field), f
    r.close()


def test_api_file_loaders_and_shift(oracle_mod, demo, tmp_path):
    """loadWorld(file) / loadBlock(file) / loadModel(file) on files written by the test, then
    shift_radiance through the API."""
    O = oracle_mod
    from lum_b200 import clum
    scene = str(tmp_path / "scene")
    world = demo["world"].copy(); world[10, 3, 3] = 99                       # out-of-palette id: clamped to 15
    write_scene(scene, world)
    r = clum.Renderer(WS, STATIC)
    r.load_world_file(scene)
    assert r.get_world_block(3, 3, 10) == STATIC
    r.set_world_block(3, 3, 10, 0)
    for i in range(1, 15):
        p = str(tmp_path / f"b{i}.vox")
        write_vox(p, demo["bricks"][i].transpose(2, 1, 0))                   # file order: x fastest
        r.load_block_file(i, p)
    mp = str(tmp_path / "m.vox")
    pal = np.random.default_rng(9).integers(0, 256, (256, 4)).astype(np.uint8)
    write_vox(mp, demo["model_tank_head"], palette_rgba=pal, matl={5: {"_type": "_emit", "_emit": 1, "_flux": 2}})
    h = r.load_mesh_file(mp, True)                                           # first model defines the palette
    mats = r.get_material_palette()
    assert mats[5, 4] == 10.0 and mats[6, 5] == 1.0
    assert np.array_equal(mats, clum.read_vox(mp)[1])
    r.set_material_palette(demo["materials"])
    r.upload_block_palette(); r.upload_material_palette()
    r.set_radiance_controller(False, 1)
    r.start_frame(); r.prepare_frame(); r.submit_frame()
    f0, *_ = oracle_frame(O, demo, np.zeros(r.cells, np.uint32), 0, [], None)
    assert np.array_equal(r.download_field(), f0)
    r.set_radiance_shift(2, -1, 1)
    r.start_frame(); r.prepare_frame(); r.submit_frame()
    shifted = O.shift_field(f0, WS, 2, -1, 1)
    f1, *_ = oracle_frame(O, demo, shifted, 1, [], None)
    assert np.array_equal(r.download_field(), f1)
    r.free_mesh(h)
    r.close()


def test_pinned_host_buffers_upload_directly(oracle_mod, small_synth):
    """A page-locked caller buffer is copied without the staging memcpy (lumcu.cu staged_upload);
    same device contents, and the buffer may be overwritten as soon as the call returns."""
    import torch
    O = oracle_mod
    sc = small_synth
    big = np.tile(sc["world"], (1, 4, 4))                                   # > 64 KiB: takes the direct path
    st = make_stage(dict(world=big, bricks=sc["bricks"], materials=sc["materials"]))
    pin = torch.empty(big.size, dtype=torch.int16, pin_memory=True)
    pin.numpy()[:] = big.reshape(-1)
    st.upload_world(pin.numpy())
    pin.numpy()[:] = 0                                                      # consumed before return
    assert np.array_equal(st.download_world(), big)
    n = st.build_requests()
    req = O.build_requests(big)
    assert n == len(req)
    scn = O.Scene(big, sc["bricks"], sc["materials"])
    ref = O.update(scn, np.zeros(scn.cells, np.uint32), req, 3)
    st.update(3)
    out = torch.empty(big.size, dtype=torch.int32, pin_memory=True)
    got = st.download_field(out.numpy().view(np.uint32))
    assert np.array_equal(got, ref["field"])
