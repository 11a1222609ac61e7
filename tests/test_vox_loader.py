"""The product's own .vox reader (lum_b200/csrc/host/vox.cpp, exported as lum_b200_read_vox) against
(a) the reference's decoder where it is available (oracle/_ref, build container only), (b) the
committed fixture that decoder produced, (c) hand-computed expectations on synthetic files."""
import os
import struct
import subprocess

import numpy as np
import pytest

from lum_b200 import clum
from voxutil import write_vox

REF_ASSETS = "/root/reference/assets"
REF_TOOL = os.path.join(os.path.dirname(__file__), "..", "oracle", "_ref", "ref_vox_dump")
BLOCKS = ["dirt", "grass", "grassNdirt", "stone_dirt", "bush", "leaves", "iron", "lamp", "stone_brick",
          "stone_brick_cracked", "stone_pack", "bark", "wood", "planks"]
MODELS = ["tank_body", "tank_head", "tank_lf_rb_leg", "tank_rf_lb_leg", "robot", "tank"]


def ref_dump(path, tmp):
    out = os.path.join(tmp, "dump.bin")
    subprocess.run([REF_TOOL, path, out], check=True)
    raw = open(out, "rb").read()
    sx, sy, sz = struct.unpack("<3i", raw[:12])
    n = sx * sy * sz
    return (np.frombuffer(raw[12:12 + n], np.uint8).reshape(sz, sy, sx),
            np.frombuffer(raw[12 + n:12 + n + 256 * 24], np.float32).reshape(256, 6))


def expected_materials(pal_by_index, matl):
    m = np.zeros((256, 6), np.float32)
    m[:, :4] = (pal_by_index.astype(np.float64) / 256.0).astype(np.float32)
    for i in range(256):
        d = matl.get(i, {})
        t = d.get("_type", "_diffuse")
        if t == "_diffuse":
            m[i, 5] = 1.0
        elif t == "_emit":
            m[i, 4] = np.float32(np.float32(d.get("_emit", 0)) * (2.0 + np.float64(np.float32(d.get("_flux", 0))) * 4.0))
            m[i, 5] = 0.5
        elif t == "_metal":
            m[i, 5] = np.float32((np.float64(np.float32(d.get("_rough", 0))) + (1.0 - np.float64(np.float32(d.get("_metal", 0))))) / 2.0)
    return m


@pytest.mark.skipif(not os.path.isdir(REF_ASSETS), reason="reference assets only exist in the build container")
def test_reference_assets_match_fixture(demo):
    for i, name in enumerate(BLOCKS, start=1):
        v, _ = clum.read_vox(f"{REF_ASSETS}/{name}.vox")
        assert np.array_equal(v.transpose(2, 1, 0), demo["bricks"][i]), name      # voxels[x][y][z]
    for name in MODELS:
        v, m = clum.read_vox(f"{REF_ASSETS}/{name}.vox")
        assert np.array_equal(v, demo[f"model_{name}"]), name
        if name == "tank_body":
            assert np.array_equal(m.view(np.uint32), demo["materials"].view(np.uint32))


@pytest.mark.skipif(not (os.path.isdir(REF_ASSETS) and os.path.exists(REF_TOOL)), reason="needs oracle/_ref (build container)")
def test_every_reference_asset_matches_reference_decoder(tmp_path):
    names = sorted(f for f in os.listdir(REF_ASSETS) if f.endswith(".vox"))
    assert len(names) >= 20
    for f in names:
        rv, rm = ref_dump(f"{REF_ASSETS}/{f}", str(tmp_path))
        v, m = clum.read_vox(f"{REF_ASSETS}/{f}")
        assert np.array_equal(v, rv), f
        assert np.array_equal(m.view(np.uint32), rm.view(np.uint32)), f


def synthetic(seed):
    rng = np.random.default_rng(seed)
    vox = (rng.random((5, 7, 6)) < 0.4) * rng.integers(1, 256, (5, 7, 6))
    pal = rng.integers(0, 256, (256, 4)).astype(np.uint8)
    matl = {3: {"_type": "_emit", "_emit": 0.75, "_flux": 3, "_ldr": 0.1},
            7: {"_type": "_metal", "_rough": 0.3, "_metal": 0.8},
            9: {"_type": "_glass", "_rough": 0.1},
            256: {"_type": "_emit", "_emit": 1, "_flux": 4}}        # id 256 is material 0
    return vox.astype(np.uint8), pal, matl


def test_synthetic_rgba_matl(tmp_path):
    vox, pal, matl = synthetic(1)
    p = str(tmp_path / "a.vox")
    write_vox(p, vox, palette_rgba=pal, matl=matl)
    v, m = clum.read_vox(p)
    assert np.array_equal(v, vox)
    by_index = np.roll(pal, 1, axis=0)           # file entry k is colour index k+1
    by_index[0, 3] = 0                           # index 0 = empty: alpha 0
    mm = dict(matl); mm[0] = mm.pop(256)
    exp = expected_materials(by_index, mm)
    assert np.array_equal(m.view(np.uint32), exp.view(np.uint32))
    assert m[3, 4] == np.float32(0.75 * 14.0) and m[0, 4] == 18.0 and m[9, 5] == 0.0


def test_synthetic_imap_remaps_voxels_palette_and_materials(tmp_path):
    vox, pal, matl = synthetic(2)
    rng = np.random.default_rng(5)
    imap = rng.permutation(256).astype(np.uint8)
    imap[imap == 0], imap[255] = imap[255], 0    # the usual convention: display 255 <-> stored 0
    p = str(tmp_path / "b.vox")
    write_vox(p, vox, palette_rgba=pal, imap=imap, matl=matl)
    v, m = clum.read_vox(p)
    inv = np.zeros(256, np.uint8); inv[imap] = np.arange(256, dtype=np.uint8)
    assert np.array_equal(v, (1 + inv[vox].astype(np.int32)).astype(np.uint8))
    assert np.array_equal(v == 0, vox == 0)      # empty stays empty under that convention
    disp = pal[(imap.astype(np.int32) + 255) & 255]
    by_index = np.roll(disp, 1, axis=0); by_index[0, 3] = 0
    mm = dict(matl); mm[0] = mm.pop(256)
    remapped = {i: mm[int(imap[(i + 255) & 255])] for i in range(256) if int(imap[(i + 255) & 255]) in mm}
    exp = expected_materials(by_index, remapped)
    assert np.array_equal(m.view(np.uint32), exp.view(np.uint32))


def test_default_palette_when_no_rgba_chunk(tmp_path):
    vox = np.zeros((2, 2, 2), np.uint8); vox[0, 0, 0] = 1
    p = str(tmp_path / "c.vox")
    write_vox(p, vox)
    _, m = clum.read_vox(p)
    c = lambda *b: [np.float32(x / 256.0) for x in b]
    assert list(m[1, :4]) == c(255, 255, 255, 255)       # colour index 1 = white
    assert list(m[2, :4]) == c(255, 255, 204, 255)
    assert list(m[216, :4]) == c(238, 0, 0, 255)         # start of the red ramp
    assert list(m[255, :4]) == c(17, 17, 17, 255)
    assert list(m[0, :4]) == c(0, 0, 0, 0)
    if os.path.exists(REF_TOOL):
        rv, rm = ref_dump(p, str(tmp_path))
        assert np.array_equal(m.view(np.uint32), rm.view(np.uint32))


@pytest.mark.skipif(not os.path.exists(REF_TOOL), reason="needs oracle/_ref (build container)")
def test_synthetic_files_match_reference_decoder(tmp_path):
    for seed in (11, 12):
        vox, pal, matl = synthetic(seed)
        imap = np.random.default_rng(seed).permutation(256).astype(np.uint8)
        p = str(tmp_path / f"s{seed}.vox")
        write_vox(p, vox, palette_rgba=pal, imap=imap if seed == 12 else None, matl=matl, version=200)
        rv, rm = ref_dump(p, str(tmp_path))
        v, m = clum.read_vox(p)
        assert np.array_equal(v, rv) and np.array_equal(m.view(np.uint32), rm.view(np.uint32))


def test_malformed_files_are_rejected(tmp_path):
    p = str(tmp_path / "bad.vox")
    open(p, "wb").write(b"NOPE" + b"\0" * 64)
    with pytest.raises(ValueError):
        clum.read_vox(p)
    open(p, "wb").write(b"VOX " + struct.pack("<I", 150))      # header only: no model
    with pytest.raises(ValueError):
        clum.read_vox(p)
    with pytest.raises(ValueError):
        clum.read_vox(str(tmp_path / "missing.vox"))


def _raw_vox(chunks):
    body = b"".join(cid + struct.pack("<II", len(c), 0) + c for cid, c in chunks)
    return b"VOX " + struct.pack("<I", 150) + b"MAIN" + struct.pack("<II", 0, len(body)) + body


def test_hostile_size_chunk_is_rejected(tmp_path):
    """three unchecked 32-bit sizes whose product wraps to a small allocation (2^22 * 2^21 * 2^21 = 0
    mod 2^64) let a crafted file write out of bounds; XYZI coordinates are bytes, so no valid file exceeds 256."""
    p = str(tmp_path / "wrap.vox")
    open(p, "wb").write(_raw_vox([(b"SIZE", struct.pack("<III", 1 << 22, 1 << 21, 1 << 21)),
                                  (b"XYZI", struct.pack("<I", 1) + bytes([200, 200, 200, 7]))]))
    with pytest.raises(ValueError):
        clum.read_vox(p)
    open(p, "wb").write(_raw_vox([(b"SIZE", struct.pack("<III", 257, 4, 4)),
                                  (b"XYZI", struct.pack("<I", 1) + bytes([1, 1, 1, 7]))]))
    with pytest.raises(ValueError):
        clum.read_vox(p)
    open(p, "wb").write(_raw_vox([(b"SIZE", struct.pack("<III", 256, 2, 2)),          # the limit itself is fine
                                  (b"XYZI", struct.pack("<I", 1) + bytes([255, 1, 1, 7]))]))
    assert clum.read_vox(p)[0].shape == (2, 2, 256)


def test_imap_that_is_not_a_permutation(tmp_path):
    """Stored indices no IMAP entry names map to display index 0 (+1), not to whatever the stack held."""
    p = str(tmp_path / "imap.vox")
    v = np.zeros((2, 2, 2), np.uint8); v[0, 0, 0] = 9; v[1, 1, 1] = 200
    imap = np.full(256, 9, np.uint8)                      # every display index -> stored 9; 200 is never named
    write_vox(p, v, imap=imap)
    a, b = clum.read_vox(p)[0], clum.read_vox(p)[0]
    assert np.array_equal(a, b)                           # deterministic
    assert a[1, 1, 1] == 1                                # 1 + inv[200], inv zero-initialised
    assert a[0, 0, 0] == 0                                # 1 + inv[9] = 1 + 255 (the last display index naming 9), as a byte
