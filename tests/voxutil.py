"""Test helper: write MagicaVoxel .vox files (SIZE/XYZI/RGBA/IMAP/MATL) from arrays, so the loader
tests do not depend on the reference's asset files (absent on the GPU box)."""
import struct

import numpy as np


def _chunk(cid, body, children=b""):
    return cid + struct.pack("<II", len(body), len(children)) + body + children


def _dict(d):
    out = struct.pack("<i", len(d))
    for k, v in d.items():
        kb, vb = k.encode(), str(v).encode()
        out += struct.pack("<i", len(kb)) + kb + struct.pack("<i", len(vb)) + vb
    return out


def write_vox(path, voxels_zyx, palette_rgba=None, imap=None, matl=None, version=150):
    """voxels_zyx: uint8 [sz,sy,sx] of STORED colour indices; palette_rgba: uint8 [256,4] in file
    order (entry k = colour index k+1); imap: uint8 [256]; matl: {id: {"_type": "_emit", ...}}."""
    v = np.asarray(voxels_zyx, np.uint8)
    sz, sy, sx = v.shape
    z, y, x = np.nonzero(v)
    xyzi = np.stack([x, y, z, v[z, y, x]], 1).astype(np.uint8)
    ch = _chunk(b"SIZE", struct.pack("<III", sx, sy, sz))
    ch += _chunk(b"XYZI", struct.pack("<I", len(xyzi)) + xyzi.tobytes())
    if palette_rgba is not None:
        ch += _chunk(b"RGBA", np.asarray(palette_rgba, np.uint8).reshape(256, 4).tobytes())
    if imap is not None:
        ch += _chunk(b"IMAP", np.asarray(imap, np.uint8).reshape(256).tobytes())
    for mid, d in (matl or {}).items():
        ch += _chunk(b"MATL", struct.pack("<i", mid) + _dict(d))
    ch += _chunk(b"NOTE", b"\x00\x00\x00\x00")                  # an unknown chunk the reader must skip
    with open(path, "wb") as f:
        f.write(b"VOX " + struct.pack("<I", version) + _chunk(b"MAIN", b"", ch))


def write_scene(path, world_zyx):
    """Lum's scene file (load_stuff.cpp:187-231): int32 sx,sy,sz then int16 ids, x fastest."""
    w = np.ascontiguousarray(world_zyx, np.int16)
    wz, wy, wx = w.shape
    with open(path, "wb") as f:
        f.write(struct.pack("<3i", wx, wy, wz) + w.tobytes())
