"""ctypes binding of include/clum.h (lum_b200/liblum.so): Lum's C99 API, headless B200 edition.

Plumbing for tests and examples only; the product is the native library.  The function names,
argument order and POD layouts are the reference's (src/renderer/api/crenderer.h:20-114)."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "liblum.so")
_lib = None


class LumSettings(C.Structure):
    _fields_ = [("world_size", C.c_int * 3), ("static_block_palette_size", C.c_int),
                ("maxParticleCount", C.c_int), ("timestampCount", C.c_int), ("fif", C.c_int),
                ("vsync", C.c_char), ("fullscreen", C.c_char), ("debug", C.c_char), ("profile", C.c_char)]


class LumMeshTransform(C.Structure):
    _fields_ = [("rot", C.c_float * 4), ("shift", C.c_float * 3)]


_P = C.c_void_p
_S = C.POINTER(LumSettings)
_I3 = C.POINTER(C.c_int)
# every symbol include/clum.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "lum_create_instance": (_P, [_S]),
    "lum_destroy_instance": (None, [_P]),
    "lum_demo_load_scene": (None, [_P, C.c_char_p]),
    "lum_init": (None, [_P, _S]),
    "lum_load_world": (None, [_P, _P]),
    "lum_load_world_file": (None, [_P, C.c_char_p]),
    "lum_set_world_block": (None, [_P, C.c_int, C.c_int, C.c_int, C.c_int16]),
    "lum_get_world_block": (C.c_int16, [_P, C.c_int, C.c_int, C.c_int]),
    "lum_load_block": (None, [_P, C.c_int16, _P]),
    "lum_load_block_file": (None, [_P, C.c_int16, C.c_char_p]),
    "lum_upload_block_palette_to_gpu": (None, [_P]),
    "lum_upload_material_palette_to_gpu": (None, [_P]),
    "lum_load_mesh_file": (_P, [_P, C.c_char_p, C.c_bool]),
    "lum_load_mesh_data": (_P, [_P, _P, C.c_int, C.c_int, C.c_int]),
    "lum_free_mesh": (None, [_P, _P]),
    "lum_load_foliage": (_P, [_P, C.c_char_p, C.c_int, C.c_int]),
    "lum_load_volumetric": (_P, [_P, C.c_float, C.c_float, C.POINTER(C.c_uint8)]),
    "lum_load_liquid": (_P, [_P, C.c_int, C.c_int]),
    "lum_start_frame": (None, [_P]),
    "lum_draw_world": (None, [_P]),
    "lum_draw_particles": (None, [_P]),
    "lum_draw_model": (None, [_P, _P, C.POINTER(LumMeshTransform)]),
    "lum_draw_foliage_block": (None, [_P, _P, _I3]),
    "lum_draw_liquid_block": (None, [_P, _P, _I3]),
    "lum_draw_volumetric_block": (None, [_P, _P, _I3]),
    "lum_prepare_frame": (None, [_P]),
    "lum_submit_frame": (None, [_P]),
    "lum_cleanup": (None, [_P]),
    "lum_get_glfw_ptr": (_P, [_P]),
    "lum_get_should_close": (C.c_bool, [_P]),
    "lum_set_should_close": (C.c_bool, [_P, C.c_bool]),
    "lum_wait_idle": (None, [_P]),
    "lum_is_profiling_enabled": (C.c_bool, [_P]),
    "lum_get_current_timestamp_count": (C.c_int, [_P]),
    "lum_get_timestamp_name": (None, [_P, C.c_int, C.c_char_p, C.c_size_t]),
    "lum_get_timestamp": (C.c_double, [_P, C.c_int]),
    "lum_get_timestamp_difference": (C.c_double, [_P, C.c_int, C.c_int]),
    "lum_b200_request_radiance_update": (None, [_P, C.c_int, C.c_int, C.c_int]),
    "lum_b200_set_camera": (None, [_P, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "lum_b200_set_radiance_shift": (None, [_P, C.c_int, C.c_int, C.c_int]),
    "lum_b200_set_radiance_controller": (None, [_P, C.c_bool, C.c_int]),
    "lum_b200_get_magic_size": (C.c_int, [_P]),
    "lum_b200_set_radiance_time_override": (None, [_P, C.c_float]),
    "lum_b200_get_request_count": (C.c_longlong, [_P]),
    "lum_b200_get_frame_index": (C.c_int, [_P]),
    "lum_b200_download_radiance_field": (None, [_P, _P]),
    "lum_b200_upload_radiance_field": (None, [_P, _P]),
    "lum_b200_glossy": (None, [_P, C.c_int, C.c_int, _P, _P, _P, _P]),
    "lum_b200_download_block_palette": (None, [_P, C.c_int, C.c_int, _P]),
    "lum_b200_download_world": (None, [_P, _P]),
    "lum_b200_get_material_palette": (None, [_P, _P]),
    "lum_b200_set_material_palette": (None, [_P, _P]),
    "lum_b200_get_model_size": (None, [_P, _I3]),
    "lum_b200_read_vox": (C.c_int, [C.c_char_p, _I3, _P, C.c_size_t, _P]),
}


def load_library():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is not built: run `python -m lum_b200.build`")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def read_vox(path):
    """-> (voxels uint8 [sz,sy,sx], materials float32 [256,6]) as Lum's loaders decode the file."""
    lib = load_library()
    size = (C.c_int * 3)()
    if lib.lum_b200_read_vox(os.fsencode(path), size, None, 0, None) != 0:
        raise ValueError(f"{path}: not a readable .vox file")
    sx, sy, sz = size
    vox = np.zeros((sz, sy, sx), np.uint8)
    mats = np.zeros((256, 6), np.float32)
    lib.lum_b200_read_vox(os.fsencode(path), size, _ptr(vox), vox.size, _ptr(mats))
    return vox, mats


class Renderer:
    """Thin object wrapper over the lum_* calls (the call sequence of examples/cdemo.c)."""

    def __init__(self, world_size=(48, 48, 16), static_block_palette_size=15, profile=True):
        self.lib = load_library()
        self.settings = LumSettings((C.c_int * 3)(*world_size), static_block_palette_size, 8128, 1024, 2,
                                    b"\0", b"\0", b"\0", b"\1" if profile else b"\0")
        self.world_size = tuple(int(v) for v in world_size)
        self.cells = self.world_size[0] * self.world_size[1] * self.world_size[2]
        self.h = self.lib.lum_create_instance(C.byref(self.settings))
        self.lib.lum_init(self.h, C.byref(self.settings))
        self._keep = []

    def close(self):
        if self.h:
            self.lib.lum_cleanup(self.h)      # also deletes the instance
            self.h = None

    def load_world(self, world):
        w = np.ascontiguousarray(world, np.int16)
        assert w.size == self.cells
        self.lib.lum_load_world(self.h, _ptr(w))

    def load_world_file(self, path):
        self.lib.lum_load_world_file(self.h, os.fsencode(path))

    def set_world_block(self, x, y, z, block):
        self.lib.lum_set_world_block(self.h, x, y, z, block)

    def get_world_block(self, x, y, z):
        return self.lib.lum_get_world_block(self.h, x, y, z)

    def load_block(self, block, voxels_xyz):
        b = np.ascontiguousarray(voxels_xyz, np.uint8)
        assert b.shape == (16, 16, 16)
        self.lib.lum_load_block(self.h, block, _ptr(b))

    def load_block_file(self, block, path):
        self.lib.lum_load_block_file(self.h, block, os.fsencode(path))

    def upload_block_palette(self):
        self.lib.lum_upload_block_palette_to_gpu(self.h)

    def upload_material_palette(self):
        self.lib.lum_upload_material_palette_to_gpu(self.h)

    def set_material_palette(self, materials):
        m = np.ascontiguousarray(materials, np.float32)
        assert m.shape == (256, 6)
        self.lib.lum_b200_set_material_palette(self.h, _ptr(m))

    def get_material_palette(self):
        m = np.zeros((256, 6), np.float32)
        self.lib.lum_b200_get_material_palette(self.h, _ptr(m))
        return m

    def load_mesh_file(self, path, try_extract_palette=True):
        return self.lib.lum_load_mesh_file(self.h, os.fsencode(path), try_extract_palette)

    def load_mesh_data(self, voxels_zyx):
        v = np.ascontiguousarray(voxels_zyx, np.uint8)
        sz, sy, sx = v.shape
        return self.lib.lum_load_mesh_data(self.h, _ptr(v), sx, sy, sz)

    def free_mesh(self, model):
        self.lib.lum_free_mesh(self.h, model)

    def model_size(self, model):
        s = (C.c_int * 3)()
        self.lib.lum_b200_get_model_size(model, s)
        return tuple(s)

    def start_frame(self):
        self.lib.lum_start_frame(self.h)

    def draw_world(self):
        self.lib.lum_draw_world(self.h)

    def draw_model(self, model, rot_wxyz=(1, 0, 0, 0), shift=(0, 0, 0)):
        t = LumMeshTransform((C.c_float * 4)(*[float(v) for v in rot_wxyz]), (C.c_float * 3)(*[float(v) for v in shift]))
        self.lib.lum_draw_model(self.h, model, C.byref(t))

    def request_radiance_update(self, x, y, z):
        self.lib.lum_b200_request_radiance_update(self.h, x, y, z)

    def prepare_frame(self):
        self.lib.lum_prepare_frame(self.h)

    def submit_frame(self):
        self.lib.lum_submit_frame(self.h)

    def wait_idle(self):
        self.lib.lum_wait_idle(self.h)

    def set_camera(self, pos, direction):
        self.lib.lum_b200_set_camera(self.h, (C.c_double * 3)(*pos), (C.c_double * 3)(*direction))

    def set_radiance_shift(self, dx, dy, dz):
        self.lib.lum_b200_set_radiance_shift(self.h, dx, dy, dz)

    def set_radiance_controller(self, enabled, magic_size=0):
        self.lib.lum_b200_set_radiance_controller(self.h, enabled, magic_size)

    def set_radiance_time_override(self, ms):
        self.lib.lum_b200_set_radiance_time_override(self.h, float(ms))

    def magic_size(self):
        return self.lib.lum_b200_get_magic_size(self.h)

    def request_count(self):
        return self.lib.lum_b200_get_request_count(self.h)

    def frame_index(self):
        return self.lib.lum_b200_get_frame_index(self.h)

    def download_field(self):
        f = np.empty(self.cells, np.uint32)
        self.lib.lum_b200_download_radiance_field(self.h, _ptr(f))
        return f

    def upload_field(self, field):
        f = np.ascontiguousarray(field, np.uint32)
        assert f.size == self.cells
        self.lib.lum_b200_upload_radiance_field(self.h, _ptr(f))

    def glossy(self, mat_norm, depth, frame=None):
        """The glossy pass with the renderer's camera: mat_norm uint8 [h,w,4], depth float32 [h,w] ->
        (frame float32 [h,w,4], covered uint8 [h,w])."""
        mn = np.ascontiguousarray(mat_norm, np.uint8)
        h, w = mn.shape[:2]
        dp = np.ascontiguousarray(depth, np.float32)
        fr = np.zeros((h, w, 4), np.float32) if frame is None else np.ascontiguousarray(frame, np.float32)
        cov = np.zeros((h, w), np.uint8)
        self.lib.lum_b200_glossy(self.h, w, h, _ptr(mn), _ptr(dp), _ptr(fr), _ptr(cov))
        return fr, cov

    def download_block_palette(self, first, count):
        b = np.empty((count, 16, 16, 16), np.uint8)
        self.lib.lum_b200_download_block_palette(self.h, first, count, _ptr(b))
        return b

    def download_world(self):
        wx, wy, wz = self.world_size
        w = np.empty((wz, wy, wx), np.int16)
        self.lib.lum_b200_download_world(self.h, _ptr(w))
        return w

    def timestamps(self):
        n = self.lib.lum_get_current_timestamp_count(self.h)
        out = []
        for i in range(n):
            buf = C.create_string_buffer(64)
            self.lib.lum_get_timestamp_name(self.h, i, buf, 64)
            out.append((buf.value.decode(), self.lib.lum_get_timestamp(self.h, i)))
        return out
