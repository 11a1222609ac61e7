"""ctypes wrapper around the CPU oracle (oracle/lum_oracle.c).

TEST INFRASTRUCTURE ONLY -- pinned against the reference's own shaders run on the CPU (see
oracle/lum_oracle.h; host-side feeders remain unpinned).  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this
module.  The product package (lum_b200) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liblum_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    """Compile the oracle (and oracle/_ref when the reference tree is present)."""
    # make decides what is stale; the _ref targets are attempted only where /root/reference exists
    args = ["make", "-C", _HERE] + (["-B"] if force else []) + ["all"]
    r = subprocess.run(args, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + r.stdout + r.stderr)
    assert os.path.exists(_LIB_PATH)
    return _LIB_PATH


class _Scene(C.Structure):
    _fields_ = [("wx", C.c_int), ("wy", C.c_int), ("wz", C.c_int),
                ("palette_capacity", C.c_int), ("rays_per_probe", C.c_int),
                ("reactiveness", C.c_float),
                ("world", C.c_void_p), ("bricks", C.c_void_p), ("materials", C.c_void_p)]


class Ray(C.Structure):
    _fields_ = [("light", C.c_float * 3), ("origin", C.c_float * 3), ("dir", C.c_float * 3),
                ("fraction", C.c_float), ("normal", C.c_float * 3), ("hit", C.c_int32),
                ("coarse_samples", C.c_int32), ("refine_samples", C.c_int32), ("voxel", C.c_int32)]


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.lum_oracle_build_requests.restype = C.c_int64
        _lib.lum_oracle_update.restype = C.c_int
        _lib.lum_oracle_update_ex.restype = C.c_int
        _lib.lum_oracle_pack_rgb10a2.restype = C.c_uint32
        _lib.lum_oracle_max_threads.restype = C.c_int
        _lib.lum_oracle_blockify.restype = C.c_int32
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class Scene:
    """world int16 [wz,wy,wx]; bricks uint8 [P,16,16,16] (x,y,z); materials float32 [256,6]."""

    def __init__(self, world, bricks, materials, rays_per_probe=64, reactiveness=0.01):
        self.world = np.ascontiguousarray(world, dtype=np.int16)
        self.bricks = np.ascontiguousarray(bricks, dtype=np.uint8)
        self.materials = np.ascontiguousarray(materials, dtype=np.float32)
        assert self.world.ndim == 3 and self.bricks.shape[1:] == (16, 16, 16)
        assert self.materials.shape == (256, 6)
        self.wz, self.wy, self.wx = self.world.shape
        self.rays = int(rays_per_probe)
        self.c = _Scene(self.wx, self.wy, self.wz, self.bricks.shape[0], self.rays,
                        np.float32(reactiveness), _p(self.world), _p(self.bricks), _p(self.materials))

    @property
    def cells(self):
        return self.wx * self.wy * self.wz


def hash33(x, y, z):
    i = (C.c_uint32 * 3)(x & 0xFFFFFFFF, y & 0xFFFFFFFF, z & 0xFFFFFFFF)
    u = (C.c_uint32 * 3)()
    f = (C.c_float * 3)()
    lib().lum_oracle_hash33(i, u, f)
    return tuple(u), tuple(f)


def direction_table(n=64):
    out = np.zeros((n, 3), np.float32)
    lib().lum_oracle_direction_table(C.c_int(n), _p(out))
    return out


def sun_dir():
    out = np.zeros(3, np.float32)
    lib().lum_oracle_sun_dir(_p(out))
    return out


def pack(rgb):
    a = (C.c_float * 3)(*[float(v) for v in rgb])
    return int(lib().lum_oracle_pack_rgb10a2(a))


def unpack(word):
    out = (C.c_float * 3)()
    lib().lum_oracle_unpack_rgb10a2(C.c_uint32(int(word)), out)
    return np.array(list(out), np.float32)


def get_voxel(scene: Scene, x, y, z):
    out = (C.c_int * 2)()
    lib().lum_oracle_get_voxel(C.byref(scene.c), C.c_int(x), C.c_int(y), C.c_int(z), out)
    return out[0], out[1]


def gather(scene: Scene, field, bp):
    field = np.ascontiguousarray(field, np.uint32)
    a = (C.c_float * 3)(*[float(v) for v in bp])
    out = (C.c_float * 3)()
    lib().lum_oracle_gather(C.byref(scene.c), _p(field), a, out)
    return np.array(list(out), np.float32)


def build_requests(world, special=None):
    world = np.ascontiguousarray(world, np.int16)
    wz, wy, wx = world.shape
    sp = None if special is None else np.ascontiguousarray(special, np.int16).reshape(-1, 4)
    nsp = 0 if sp is None else sp.shape[0]
    cap = world.size + nsp
    out = np.zeros((cap, 4), np.int16)
    n = lib().lum_oracle_build_requests(C.c_int(wx), C.c_int(wy), C.c_int(wz), _p(world), _p(sp),
                                        C.c_int(nsp), _p(out), C.c_int64(cap))
    return out[:n].copy()


def trace(scene: Scene, field, probe, gid_x, ray_index, time_seed, table=None):
    field = np.ascontiguousarray(field, np.uint32)
    if table is None:
        table = direction_table(scene.rays)
    pr = (C.c_int16 * 3)(*[int(v) for v in probe[:3]])
    r = Ray()
    lib().lum_oracle_trace(C.byref(scene.c), _p(field), _p(table), pr, C.c_uint32(gid_x),
                           C.c_int(ray_index), C.c_int(time_seed), C.byref(r))
    return r


def update(scene: Scene, field_old, requests, time_seed, size=1, shift=0, nthreads=0,
           want_rays=False, wg_range=None):
    """Returns dict(field, pre_blend, totals=(S,H,N), threads[, ray_samples, ray_hits, ray_coarse])."""
    field_old = np.ascontiguousarray(field_old, np.uint32).reshape(-1)
    assert field_old.size == scene.cells
    req = np.ascontiguousarray(requests, np.int16).reshape(-1, 4)
    nreq = req.shape[0]
    nwg = nreq // size if size > 0 else 0
    field_new = np.empty_like(field_old)
    pre = np.zeros((nwg, 3), np.float32)
    rs = np.zeros(nwg * scene.rays, np.uint32) if want_rays else None
    rh = np.zeros(nwg * scene.rays, np.uint8) if want_rays else None
    rcs = np.zeros(nwg * scene.rays, np.uint16) if want_rays else None
    tot = (C.c_uint64 * 3)()
    b, e = (0, -1) if wg_range is None else wg_range
    used = lib().lum_oracle_update_ex(C.byref(scene.c), _p(field_old), _p(field_new), _p(req),
                                      C.c_int64(nreq), C.c_int(time_seed), C.c_int(size), C.c_int(shift),
                                      _p(pre), _p(rs), _p(rh), _p(rcs), tot, C.c_int(nthreads),
                                      C.c_int64(b), C.c_int64(e))
    res = dict(field=field_new, pre_blend=pre, totals=(int(tot[0]), int(tot[1]), int(tot[2])),
               threads=used)
    if want_rays:
        res["ray_samples"] = rs
        res["ray_hits"] = rh
        res["ray_coarse"] = rcs
    return res


# ---- glossy.frag ---------------------------------------------------------------
class GlossyView(C.Structure):
    _fields_ = [("campos", C.c_float * 3), ("camdir", C.c_float * 3), ("horizline_scaled", C.c_float * 3),
                ("vertiline_scaled", C.c_float * 3), ("global_light_dir", C.c_float * 3),
                ("frame_size", C.c_float * 2)]


def glossy_view(view):
    v = GlossyView()
    for k in ("campos", "camdir", "horizline_scaled", "vertiline_scaled", "global_light_dir"):
        getattr(v, k)[:] = [float(np.float32(x)) for x in view[k]]
    v.frame_size[:] = [float(np.float32(x)) for x in view["frame_size"]]
    return v


def camera_view(campos, camdir, width, height, pixels_in_voxel=5.0, light_dir=(0.5, 0.5, -0.9)):
    """The uniform-block fields as Camera::updateCamera + start_raygen fill them (internal_types.hpp:54-66,
    internal_render.cpp:921-934,1580-1592): double precision on the host, rounded to f32 for the UBO."""
    d = np.asarray(camdir, np.float64); d = d / np.linalg.norm(d)
    plane = np.array([d[0], d[1], 0.0]); plane = plane / np.linalg.norm(plane)
    horiz = np.cross(plane, [0.0, 0.0, 1.0]); horiz = horiz / np.linalg.norm(horiz)
    verti = np.cross(d, horiz); verti = verti / np.linalg.norm(verti)
    view_size = np.array([width, height], np.float64) / pixels_in_voxel
    ld = np.asarray(light_dir, np.float64); ld = ld / np.linalg.norm(ld)
    return dict(campos=np.asarray(campos, np.float64).astype(np.float32), camdir=d.astype(np.float32),
                horizline_scaled=(horiz * view_size[0] / 2.0).astype(np.float32),
                vertiline_scaled=(verti * view_size[1] / 2.0).astype(np.float32),
                global_light_dir=ld.astype(np.float32), frame_size=np.array([width, height], np.float32))


def gbuffer(scene: Scene, view, width, height):
    """Fabricated G-buffer (test input): mat_norm uint8 [h,w,4], depth float32 [h,w]."""
    mn = np.zeros((height, width, 4), np.uint8)
    dp = np.zeros((height, width), np.float32)
    v = glossy_view(view)
    lib().lum_oracle_gbuffer(C.byref(scene.c), C.byref(v), C.c_int(width), C.c_int(height), _p(mn), _p(dp))
    return mn, dp


def glossy(scene: Scene, field, view, mat_norm, depth, frame=None, nthreads=0):
    """Returns dict(frame float32 [h,w,4], covered uint8 [h,w], samples uint32 [h,w], coarse uint32 [h,w])."""
    field = np.ascontiguousarray(field, np.uint32).reshape(-1)
    assert field.size == scene.cells
    mn = np.ascontiguousarray(mat_norm, np.uint8)
    h, w = mn.shape[:2]
    dp = np.ascontiguousarray(depth, np.float32)
    fr = np.zeros((h, w, 4), np.float32) if frame is None else np.array(frame, np.float32, copy=True)
    cov = np.zeros((h, w), np.uint8)
    ns = np.zeros((h, w), np.uint32)
    nc = np.zeros((h, w), np.uint32)
    v = glossy_view(view)
    lib().lum_oracle_glossy(C.byref(scene.c), _p(field), C.byref(v), C.c_int(w), C.c_int(h), _p(mn), _p(dp),
                            _p(fr), _p(cov), _p(ns), _p(nc), C.c_int(nthreads))
    return dict(frame=fr, covered=cov, samples=ns, coarse=nc)


def shift_field(field, dims, dx, dy, dz):
    wx, wy, wz = dims
    field = np.ascontiguousarray(field, np.uint32).reshape(-1)
    out = np.empty_like(field)
    lib().lum_oracle_shift_field(C.c_int(wx), C.c_int(wy), C.c_int(wz), _p(field), _p(out),
                                 C.c_int(dx), C.c_int(dy), C.c_int(dz))
    return out


def max_threads():
    return int(lib().lum_oracle_max_threads())


# ---- feeders (lum_oracle_feeders.c) ---------------------------------------------------------
class Placement(C.Structure):
    _fields_ = [("aabb_min", C.c_float * 3), ("aabb_max", C.c_float * 3),
                ("block_min", C.c_int32 * 3), ("block_max", C.c_int32 * 3),
                ("map_shift", C.c_int32 * 3), ("map_area", C.c_int32 * 3),
                ("inverse_trans", C.c_float * 16)]


def place_model(rot_wxyz, shift, size, world_size):
    """rot (w,x,y,z), shift (voxels), size (sx,sy,sz), world_size (blocks) -> Placement."""
    r = (C.c_float * 4)(*[float(v) for v in rot_wxyz])
    s = (C.c_float * 3)(*[float(v) for v in shift])
    z = (C.c_int32 * 3)(*[int(v) for v in size])
    w = (C.c_int32 * 3)(*[int(v) for v in world_size])
    out = Placement()
    lib().lum_oracle_place_model(r, s, z, w, C.byref(out))
    return out


def blockify(origin_world, static_count, placements):
    """-> (current_world, copy_src, copy_dst): start_blockify + blockify_mesh in draw order."""
    ow = np.ascontiguousarray(origin_world, np.int16)
    wz, wy, wx = ow.shape
    cur = np.empty_like(ow)
    arr = (Placement * max(1, len(placements)))(*placements)
    cap = ow.size
    src = np.zeros(cap, np.int32)
    dst = np.zeros(cap, np.int32)
    n = lib().lum_oracle_blockify(C.c_int(wx), C.c_int(wy), C.c_int(wz), _p(ow), _p(cur), C.c_int(static_count),
                                  arr, C.c_int(len(placements)), _p(src), _p(dst), C.c_int32(cap))
    return cur, src[:n].copy(), dst[:n].copy()


def copy_bricks(bricks, src, dst):
    """In place on `bricks` (uint8 [P,16,16,16], C-contiguous)."""
    assert bricks.flags.c_contiguous and bricks.dtype == np.uint8
    s = np.ascontiguousarray(src, np.int32)
    d = np.ascontiguousarray(dst, np.int32)
    lib().lum_oracle_copy_bricks(_p(bricks), C.c_int(bricks.shape[0]), _p(s), _p(d), C.c_int32(len(s)))


def map_model(world, bricks, model, placement, static_count):
    """map.comp for one model; `bricks` modified in place.  model: uint8 [sz,sy,sx] (x fastest)."""
    w = np.ascontiguousarray(world, np.int16)
    wz, wy, wx = w.shape
    m = np.ascontiguousarray(model, np.uint8)
    sz, sy, sx = m.shape
    assert bricks.flags.c_contiguous and bricks.dtype == np.uint8
    lib().lum_oracle_map_model(C.c_int(wx), C.c_int(wy), C.c_int(wz), _p(w), _p(bricks), C.c_int(bricks.shape[0]),
                               _p(m), C.c_int(sx), C.c_int(sy), C.c_int(sz), placement.inverse_trans,
                               placement.map_shift, placement.map_area, C.c_int(static_count))


# ---- oracle/_ref: the reference's own radiance.comp executed on the CPU -----------------------
# (oracle/ref_radiance_glsl.cpp; built only where /root/reference exists, travels as a built .so)
_REF_GLSL_PATH = os.path.join(_HERE, "_ref", "libref_radiance_glsl.so")
_ref_glsl = None


def ref_glsl_available() -> bool:
    return os.path.exists(_REF_GLSL_PATH)


def ref_glsl_update(scene: Scene, field_old, requests, time_seed, size=1, shift=0, want_rays=False):
    """One dispatch of the reference's shaders/radiance.comp (its own text, compiled through
    oracle/glsl_shim.hpp).  Same inputs as update(); request coordinates must fit the reference's
    i8vec4 (<= 127).  Returns dict(field, pre_blend[, ray_light, ray_fetches]) where
    ray_fetches = texelFetch(blocks) calls per invocation = 2 + 2*coarse + refinement samples."""
    global _ref_glsl
    if _ref_glsl is None:
        _ref_glsl = C.CDLL(_REF_GLSL_PATH)
        _ref_glsl.ref_glsl_radiance_update.restype = C.c_int
    assert scene.rays == 64, "the reference defines RAYS_PER_PROBE = 64 only"
    field_old = np.ascontiguousarray(field_old, np.uint32).reshape(-1)
    assert field_old.size == scene.cells
    req16 = np.ascontiguousarray(requests, np.int16).reshape(-1, 4)
    assert req16.size == 0 or (req16.max() <= 127 and req16.min() >= -128)
    req = np.ascontiguousarray(req16.astype(np.int8))
    nwg = req.shape[0] // size
    field_new = np.empty_like(field_old)
    pre = np.zeros((nwg, 3), np.float32)
    rl = np.zeros((nwg * 64, 3), np.float32) if want_rays else None
    rf = np.zeros(nwg * 64, np.uint32) if want_rays else None
    rc = _ref_glsl.ref_glsl_radiance_update(
        C.c_int(scene.wx), C.c_int(scene.wy), C.c_int(scene.wz), _p(scene.world), _p(scene.bricks),
        C.c_int(scene.bricks.shape[0]), _p(scene.materials), _p(field_old), _p(field_new), _p(req),
        C.c_long(req.shape[0]), C.c_int(time_seed), C.c_int(size), C.c_int(shift), _p(pre), _p(rl), _p(rf))
    assert rc == 0, rc
    res = dict(field=field_new, pre_blend=pre)
    if want_rays:
        res["ray_light"] = rl
        res["ray_fetches"] = rf
    return res


_REF_MAP_PATH = os.path.join(_HERE, "_ref", "libref_map_glsl.so")
_ref_map = None


def ref_map_available() -> bool:
    return os.path.exists(_REF_MAP_PATH)


def ref_glsl_map_model(world, bricks, model, placement):
    """One map_mesh dispatch of the reference's own shaders/map.comp (oracle/ref_map_glsl.cpp);
    same arguments as map_model().  The shader hard-codes STATIC_BLOCK_COUNT = 15 (map.comp:22)."""
    global _ref_map
    if _ref_map is None:
        _ref_map = C.CDLL(_REF_MAP_PATH)
        _ref_map.ref_glsl_map_model.restype = C.c_int
    w = np.ascontiguousarray(world, np.int16)
    wz, wy, wx = w.shape
    m = np.ascontiguousarray(model, np.uint8)
    sz, sy, sx = m.shape
    assert bricks.flags.c_contiguous and bricks.dtype == np.uint8
    rc = _ref_map.ref_glsl_map_model(C.c_int(wx), C.c_int(wy), C.c_int(wz), _p(w), _p(bricks), C.c_int(bricks.shape[0]),
                                     _p(m), C.c_int(sx), C.c_int(sy), C.c_int(sz), placement.inverse_trans,
                                     placement.map_shift, placement.map_area)
    assert rc == 0, rc


_REF_GLOSSY_PATH = os.path.join(_HERE, "_ref", "libref_glossy_glsl.so")
_ref_glossy = None


def ref_glossy_available() -> bool:
    return os.path.exists(_REF_GLOSSY_PATH)


def ref_glsl_glossy(scene: Scene, field, view, mat_norm, depth, frame=None):
    """The reference's own shaders/glossy.frag (oracle/ref_glossy_glsl.cpp), one invocation per covered
    fragment; same arguments and result as glossy(), plus fetches = texelFetch(blocks) calls per fragment
    (= 2 + 2*coarse iterations + refinement samples)."""
    global _ref_glossy
    if _ref_glossy is None:
        _ref_glossy = C.CDLL(_REF_GLOSSY_PATH)
        _ref_glossy.ref_glsl_glossy.restype = C.c_int
    field = np.ascontiguousarray(field, np.uint32).reshape(-1)
    assert field.size == scene.cells
    mn = np.ascontiguousarray(mat_norm, np.uint8)
    h, w = mn.shape[:2]
    dp = np.ascontiguousarray(depth, np.float32)
    fr = np.zeros((h, w, 4), np.float32) if frame is None else np.array(frame, np.float32, copy=True)
    cov = np.zeros((h, w), np.uint8)
    nf = np.zeros((h, w), np.uint32)
    v = np.concatenate([np.asarray(view[k], np.float32).reshape(-1) for k in
                        ("campos", "camdir", "horizline_scaled", "vertiline_scaled", "global_light_dir", "frame_size")])
    v = np.ascontiguousarray(v, np.float32)
    assert v.size == 17
    rc = _ref_glossy.ref_glsl_glossy(C.c_int(scene.wx), C.c_int(scene.wy), C.c_int(scene.wz), _p(scene.world),
                                     _p(scene.bricks), C.c_int(scene.bricks.shape[0]), _p(scene.materials), _p(field),
                                     _p(v), C.c_int(w), C.c_int(h), _p(mn), _p(dp), _p(fr), _p(cov), _p(nf))
    assert rc == 0, rc
    return dict(frame=fr, covered=cov, fetches=nf)


# ---- oracle/_ref: the reference's HOST-side C++ of the path, compiled in place (ref_host.cpp) ----
_REF_HOST_PATH = os.path.join(_HERE, "_ref", "libref_host.so")
_ref_host = None


def ref_host_available() -> bool:
    return os.path.exists(_REF_HOST_PATH)


def _ref_host_lib():
    global _ref_host
    if _ref_host is None:
        _ref_host = C.CDLL(_REF_HOST_PATH)
        _ref_host.ref_host_place_model.restype = C.c_int
        _ref_host.ref_host_blockify.restype = C.c_int
        _ref_host.ref_host_build_requests.restype = C.c_longlong
    return _ref_host


def ref_host_place_model(rot_wxyz, shift, size, world_size):
    """The reference's get_shift + blockify_mesh border + map_mesh arithmetic (internal_render.cpp:989-1013,
    1030-1037, 1501-1510) -> a Placement like place_model()'s.  block_min/max come back as the extent of the cells
    blockify_mesh actually allocated on an empty world."""
    r = (C.c_float * 4)(*[float(v) for v in rot_wxyz])
    s = (C.c_float * 3)(*[float(v) for v in shift])
    z = (C.c_int32 * 3)(*[int(v) for v in size])
    w = (C.c_int32 * 3)(*[int(v) for v in world_size])
    out = Placement()
    _ref_host_lib().ref_host_place_model(r, s, z, w, out.aabb_min, out.aabb_max, out.block_min, out.block_max,
                                         out.map_shift, out.map_area, out.inverse_trans)
    return out


def ref_host_blockify(origin_world, static_count, models):
    """The reference's start_blockify + blockify_mesh (internal_render.cpp:977-987, 1029-1076) per model in draw
    order.  models: [(rot_wxyz, shift, size)].  -> (current_world, copy_src, copy_dst, queued_image_copies)."""
    ow = np.ascontiguousarray(origin_world, np.int16)
    wz, wy, wx = ow.shape
    cur = np.empty_like(ow)
    rot = np.ascontiguousarray([m[0] for m in models], np.float32).reshape(-1)
    sh = np.ascontiguousarray([m[1] for m in models], np.float32).reshape(-1)
    sz = np.ascontiguousarray([m[2] for m in models], np.int32).reshape(-1)
    cap = ow.size
    src = np.zeros(cap, np.int32); dst = np.zeros(cap, np.int32)
    queued = C.c_int32()
    n = _ref_host_lib().ref_host_blockify(C.c_int(wx), C.c_int(wy), C.c_int(wz), _p(ow), _p(cur), C.c_int(static_count),
                                          C.c_int(len(models)), _p(rot), _p(sh), _p(sz), _p(src), _p(dst), C.c_int32(cap),
                                          C.byref(queued))
    return cur, src[:n].copy(), dst[:n].copy(), int(queued.value)


def ref_host_build_requests(world, special=None):
    """The reference's update_radiance list rule (internal_render.cpp:1132-1169).  Coordinates are i8vec4 there:
    worlds up to 127 blocks per axis.  -> int16 [n,4] like build_requests()."""
    world = np.ascontiguousarray(world, np.int16)
    wz, wy, wx = world.shape
    assert max(wx, wy, wz) <= 127
    sp = None if special is None else np.ascontiguousarray(special, np.int8).reshape(-1, 4)
    nsp = 0 if sp is None else sp.shape[0]
    cap = world.size + nsp
    out = np.zeros((cap, 4), np.int8)
    n = _ref_host_lib().ref_host_build_requests(C.c_int(wx), C.c_int(wy), C.c_int(wz), _p(world), _p(sp), C.c_int(nsp),
                                                _p(out), C.c_longlong(cap))
    assert n >= 0
    return out[:n].astype(np.int16)
