"""lum_b200 -- B200-native radiance-probe update stage for the Lum voxel renderer.

Python here is plumbing only: a ctypes binding of the C-ABI in include/lumcu.h (the product is
the CUDA library lum_b200/liblumcu.so) used by tests/ and bench.py.  There is no CPU fallback:
importing works anywhere, but creating a stage needs the built library and a CUDA device.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("LUMCU_LIB") or os.path.join(HERE, "liblumcu.so")   # LUMCU_LIB: tuning builds

DEBUG_PRE_BLEND = 1
DEBUG_RAYS = 2
SHARD_CONTIGUOUS, SHARD_INTERLEAVED = 0, 1      # lumcu_set_shard_mode

_lib = None


class LumcuError(RuntimeError):
    pass


class LumcuDesc(C.Structure):
    _fields_ = [("world_size", C.c_int32 * 3), ("palette_capacity", C.c_int32),
                ("rays_per_probe", C.c_int32), ("reactiveness", C.c_float),
                ("device", C.c_int32), ("max_requests", C.c_int64)]


class LumcuGlossyView(C.Structure):
    _fields_ = [("campos", C.c_float * 3), ("camdir", C.c_float * 3), ("horizline_scaled", C.c_float * 3),
                ("vertiline_scaled", C.c_float * 3), ("global_light_dir", C.c_float * 3),
                ("frame_size", C.c_float * 2)]


# every symbol include/lumcu.h declares: name -> (restype, argtypes)
_P = C.c_void_p
SYMBOLS = {
    "lumcu_last_error": (C.c_char_p, []),
    "lumcu_version": (C.c_int, []),
    "lumcu_create": (C.c_int, [C.POINTER(LumcuDesc), C.POINTER(_P)]),
    "lumcu_destroy": (C.c_int, [_P]),
    "lumcu_upload_world": (C.c_int, [_P, _P]),
    "lumcu_upload_world_async": (C.c_int, [_P, _P]),
    "lumcu_upload_bricks": (C.c_int, [_P, C.c_int32, C.c_int32, _P]),
    "lumcu_upload_materials": (C.c_int, [_P, _P]),
    "lumcu_set_requests": (C.c_int, [_P, _P, C.c_int64]),
    "lumcu_build_requests": (C.c_int, [_P, _P, C.c_int32, C.POINTER(C.c_int64)]),
    "lumcu_request_count": (C.c_int, [_P, C.POINTER(C.c_int64)]),
    "lumcu_download_requests": (C.c_int, [_P, _P, C.c_int64]),
    "lumcu_update": (C.c_int, [_P, C.c_int32, C.c_int32, C.c_int32]),
    "lumcu_update_range": (C.c_int, [_P, C.c_int32, C.c_int32, C.c_int32, C.c_int64, C.c_int64]),
    "lumcu_slab_range": (C.c_int, [_P, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                   C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "lumcu_update_slab": (C.c_int, [_P, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32]),
    "lumcu_shift": (C.c_int, [_P, C.c_int32, C.c_int32, C.c_int32]),
    "lumcu_ipc_export": (C.c_int, [_P, _P]),
    "lumcu_peers_open": (C.c_int, [_P, C.c_int32, C.c_int32, _P]),
    "lumcu_peers_close": (C.c_int, [_P]),
    "lumcu_update_slab_exchange": (C.c_int, [_P, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32]),
    "lumcu_update_shard": (C.c_int, [_P, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32]),
    "lumcu_update_shard_exchange": (C.c_int, [_P, C.c_int32, C.c_int32, C.c_int32]),
    "lumcu_upload_model": (C.c_int, [_P, _P, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_int32)]),
    "lumcu_free_model": (C.c_int, [_P, C.c_int32]),
    "lumcu_copy_bricks": (C.c_int, [_P, _P, _P, C.c_int32]),
    "lumcu_map_model": (C.c_int, [_P, C.c_int32, _P, _P, _P, C.c_int32]),
    "lumcu_download_bricks": (C.c_int, [_P, C.c_int32, C.c_int32, _P]),
    "lumcu_download_world": (C.c_int, [_P, _P]),
    "lumcu_upload_field": (C.c_int, [_P, _P]),
    "lumcu_download_field": (C.c_int, [_P, _P]),
    "lumcu_download_field_async": (C.c_int, [_P, _P, C.c_int64, C.c_int64]),
    "lumcu_record_fence": (C.c_int, [_P, C.c_int32]),
    "lumcu_wait_fence": (C.c_int, [_P, C.c_int32]),
    "lumcu_direction_table": (C.c_int, [_P, _P]),
    "lumcu_set_debug": (C.c_int, [_P, C.c_int32]),
    "lumcu_debug_extent": (C.c_int, [_P, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "lumcu_download_debug": (C.c_int, [_P, _P, _P, _P, _P]),
    "lumcu_debug_unpack_table": (C.c_int, [_P, _P]),
    "lumcu_wait_idle": (C.c_int, [_P]),
    "lumcu_last_update_ms": (C.c_int, [_P, C.POINTER(C.c_float)]),
    "lumcu_set_kernel_variant": (C.c_int, [_P, C.c_int32]),
    "lumcu_set_profiling": (C.c_int, [_P, C.c_int32]),
    "lumcu_set_shard_mode": (C.c_int, [_P, C.c_int32]),
    "lumcu_shard_chunk": (C.c_int, [_P, C.c_int32, C.c_int32, C.POINTER(C.c_int32)]),
    "lumcu_last_kernel_ms": (C.c_int, [_P, C.POINTER(C.c_float)]),
    "lumcu_launch_count": (C.c_int, [_P, C.POINTER(C.c_uint64)]),
    "lumcu_glossy": (C.c_int, [_P, C.POINTER(LumcuGlossyView), C.c_int32, C.c_int32, _P, _P, _P, _P]),
    "lumcu_glossy_device": (C.c_int, [_P, C.POINTER(LumcuGlossyView), C.c_int32, C.c_int32,
                                      C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64]),
    "lumcu_field_device_ptr": (C.c_int, [_P, C.POINTER(C.c_uint64)]),
    "lumcu_world_device_ptr": (C.c_int, [_P, C.POINTER(C.c_uint64)]),
    "lumcu_stream": (C.c_int, [_P, C.POINTER(C.c_uint64)]),
}


def load_library():
    """dlopen liblumcu.so and bind every declared symbol.  Fails loudly if it is not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise LumcuError(f"{LIB_PATH} is not built: run `python -m lum_b200.build` "
                             "(there is no CPU fallback for the radiance stage)")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(lib, name)      # AttributeError if the library lacks a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class RadianceStage:
    """One device context of the radiance stage (LumcuCtx).  Arrays use Lum's layouts:
    world int16 [wz,wy,wx]; bricks uint8 [n,16,16,16] (x,y,z); materials float32 [256,6];
    field uint32 [wz,wy,wx]; requests int16 [n,4]."""

    def __init__(self, world_size, palette_capacity, rays_per_probe=64, reactiveness=0.01,
                 device=0, max_requests=0):
        self.lib = load_library()
        self.wx, self.wy, self.wz = (int(v) for v in world_size)
        self.rays = int(rays_per_probe)
        self.cells = self.wx * self.wy * self.wz
        d = LumcuDesc((C.c_int32 * 3)(self.wx, self.wy, self.wz), int(palette_capacity), self.rays,
                      float(reactiveness), int(device), int(max_requests))
        self.ctx = _P()
        self._ck(self.lib.lumcu_create(C.byref(d), C.byref(self.ctx)))

    def _ck(self, rc):
        if rc != 0:
            raise LumcuError(f"lumcu error {rc}: {self.lib.lumcu_last_error().decode()}")

    def close(self):
        if getattr(self, "ctx", None):
            self.lib.lumcu_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- uploads --------------------------------------------------------------------------
    def upload_world(self, world):
        w = np.ascontiguousarray(world, np.int16)
        assert w.size == self.cells
        self._ck(self.lib.lumcu_upload_world(self.ctx, _ptr(w)))

    def upload_world_async(self, world_pinned):
        """world_pinned: an int16 array in page-locked memory that stays untouched until the stage is next synchronised."""
        assert world_pinned.dtype == np.int16 and world_pinned.size == self.cells and world_pinned.flags["C_CONTIGUOUS"]
        self._ck(self.lib.lumcu_upload_world_async(self.ctx, _ptr(world_pinned)))

    def upload_bricks(self, first_id, bricks):
        b = np.ascontiguousarray(bricks, np.uint8).reshape(-1, 4096)
        self._ck(self.lib.lumcu_upload_bricks(self.ctx, int(first_id), b.shape[0], _ptr(b)))

    def upload_materials(self, materials):
        m = np.ascontiguousarray(materials, np.float32)
        assert m.size == 256 * 6
        self._ck(self.lib.lumcu_upload_materials(self.ctx, _ptr(m)))

    def upload_field(self, field):
        f = np.ascontiguousarray(field, np.uint32)
        assert f.size == self.cells
        self._ck(self.lib.lumcu_upload_field(self.ctx, _ptr(f)))

    def set_requests(self, requests):
        r = np.ascontiguousarray(requests, np.int16).reshape(-1, 4)
        self._ck(self.lib.lumcu_set_requests(self.ctx, _ptr(r), r.shape[0]))

    def build_requests(self, special=None, sync=True):
        """sync=False: nothing waits for the GPU; the count stays on the device (request_count() fetches it)."""
        sp = None if special is None else np.ascontiguousarray(special, np.int16).reshape(-1, 4)
        n = C.c_int64()
        self._ck(self.lib.lumcu_build_requests(self.ctx, _ptr(sp), 0 if sp is None else sp.shape[0], C.byref(n) if sync else None))
        return n.value if sync else None

    # -- the hot path ---------------------------------------------------------------------
    def update(self, time_seed, size=1, shift=0):
        self._ck(self.lib.lumcu_update(self.ctx, int(time_seed), int(size), int(shift)))

    def update_range(self, time_seed, size, shift, wg_begin, wg_end):
        self._ck(self.lib.lumcu_update_range(self.ctx, int(time_seed), int(size), int(shift),
                                             int(wg_begin), int(wg_end)))

    def update_slab(self, time_seed, size, shift, z_begin, z_end):
        self._ck(self.lib.lumcu_update_slab(self.ctx, int(time_seed), int(size), int(shift),
                                            int(z_begin), int(z_end)))

    def slab_range(self, size, shift, z_begin, z_end):
        b, e = C.c_int64(), C.c_int64()
        self._ck(self.lib.lumcu_slab_range(self.ctx, int(size), int(shift), int(z_begin), int(z_end),
                                           C.byref(b), C.byref(e)))
        return b.value, e.value

    # -- multi-GPU exchange over peer memory -------------------------------------------------
    def ipc_export(self):
        blob = np.zeros(128, np.uint8)
        self._ck(self.lib.lumcu_ipc_export(self.ctx, _ptr(blob)))
        return blob

    def peers_open(self, rank, blobs):
        b = np.ascontiguousarray(blobs, np.uint8).reshape(-1, 128)
        self._ck(self.lib.lumcu_peers_open(self.ctx, int(rank), b.shape[0], _ptr(b)))

    def peers_close(self):
        self._ck(self.lib.lumcu_peers_close(self.ctx))

    def update_slab_exchange(self, time_seed, size, shift, z_begin, z_end):
        self._ck(self.lib.lumcu_update_slab_exchange(self.ctx, int(time_seed), int(size), int(shift),
                                                     int(z_begin), int(z_end)))

    # -- glossy reflections (glossy.frag) ----------------------------------------------------
    @staticmethod
    def glossy_view(view):
        """dict(campos, camdir, horizline_scaled, vertiline_scaled, global_light_dir, frame_size) -> LumcuGlossyView"""
        v = LumcuGlossyView()
        for k in ("campos", "camdir", "horizline_scaled", "vertiline_scaled", "global_light_dir"):
            getattr(v, k)[:] = [float(np.float32(x)) for x in view[k]]
        v.frame_size[:] = [float(np.float32(x)) for x in view["frame_size"]]
        return v

    def glossy(self, view, mat_norm, depth, frame=None):
        """mat_norm uint8 [h,w,4]; depth float32 [h,w]; returns (frame float32 [h,w,4], covered uint8 [h,w])."""
        mn = np.ascontiguousarray(mat_norm, np.uint8)
        h, w = mn.shape[:2]
        dp = np.ascontiguousarray(depth, np.float32)
        assert mn.shape == (h, w, 4) and dp.shape == (h, w)
        fr = np.zeros((h, w, 4), np.float32) if frame is None else np.ascontiguousarray(frame, np.float32)
        cov = np.zeros((h, w), np.uint8)
        v = self.glossy_view(view)
        self._ck(self.lib.lumcu_glossy(self.ctx, C.byref(v), w, h, _ptr(mn), _ptr(dp), _ptr(fr), _ptr(cov)))
        return fr, cov

    def update_shard(self, time_seed, size, shift, rank, nranks):
        self._ck(self.lib.lumcu_update_shard(self.ctx, int(time_seed), int(size), int(shift), int(rank), int(nranks)))

    def update_shard_exchange(self, time_seed, size=1, shift=0):
        self._ck(self.lib.lumcu_update_shard_exchange(self.ctx, int(time_seed), int(size), int(shift)))

    def shift(self, dx, dy, dz):
        self._ck(self.lib.lumcu_shift(self.ctx, int(dx), int(dy), int(dz)))

    # -- feeders: blockify copies + map.comp ------------------------------------------------
    def upload_model(self, voxels_zyx):
        v = np.ascontiguousarray(voxels_zyx, np.uint8)
        sz, sy, sx = v.shape
        h = C.c_int32()
        self._ck(self.lib.lumcu_upload_model(self.ctx, _ptr(v), sx, sy, sz, C.byref(h)))
        return h.value

    def free_model(self, model):
        self._ck(self.lib.lumcu_free_model(self.ctx, int(model)))

    def copy_bricks(self, src_ids, dst_ids):
        s = np.ascontiguousarray(src_ids, np.int32)
        d = np.ascontiguousarray(dst_ids, np.int32)
        assert s.shape == d.shape
        self._ck(self.lib.lumcu_copy_bricks(self.ctx, _ptr(s), _ptr(d), s.size))

    def map_model(self, model, inverse_trans, shift, area, static_block_count):
        m = np.ascontiguousarray(inverse_trans, np.float32).reshape(16)
        sh = np.ascontiguousarray(shift, np.int32).reshape(3)
        ar = np.ascontiguousarray(area, np.int32).reshape(3)
        self._ck(self.lib.lumcu_map_model(self.ctx, int(model), _ptr(m), _ptr(sh), _ptr(ar), int(static_block_count)))

    def download_bricks(self, first, count):
        out = np.empty((count, 16, 16, 16), np.uint8)
        self._ck(self.lib.lumcu_download_bricks(self.ctx, int(first), int(count), _ptr(out)))
        return out

    def download_world(self):
        out = np.empty((self.wz, self.wy, self.wx), np.int16)
        self._ck(self.lib.lumcu_download_world(self.ctx, _ptr(out)))
        return out

    # -- readbacks ------------------------------------------------------------------------
    def request_count(self):
        n = C.c_int64()
        self._ck(self.lib.lumcu_request_count(self.ctx, C.byref(n)))
        return n.value

    def download_requests(self):
        n = self.request_count()
        out = np.zeros((max(n, 1), 4), np.int16)
        self._ck(self.lib.lumcu_download_requests(self.ctx, _ptr(out), out.shape[0]))
        return out[:n]

    def download_field(self, out=None):
        if out is None:
            out = np.empty(self.cells, np.uint32)
        self._ck(self.lib.lumcu_download_field(self.ctx, _ptr(out)))
        return out

    def download_field_async(self, out_pinned, first=0, count=None):
        count = self.cells - first if count is None else count
        assert out_pinned.dtype == np.uint32 and out_pinned.size >= count
        self._ck(self.lib.lumcu_download_field_async(self.ctx, _ptr(out_pinned), int(first), int(count)))

    def record_fence(self, slot):
        self._ck(self.lib.lumcu_record_fence(self.ctx, int(slot)))

    def wait_fence(self, slot):
        self._ck(self.lib.lumcu_wait_fence(self.ctx, int(slot)))

    def direction_table(self):
        out = np.zeros((self.rays, 3), np.float32)
        self._ck(self.lib.lumcu_direction_table(self.ctx, _ptr(out)))
        return out

    def set_debug(self, flags):
        self._ck(self.lib.lumcu_set_debug(self.ctx, int(flags)))

    def debug_extent(self):
        n, b = C.c_int64(), C.c_int64()
        self._ck(self.lib.lumcu_debug_extent(self.ctx, C.byref(n), C.byref(b)))
        return int(n.value), int(b.value)

    def download_debug(self, nwg=None, pre_blend=True, rays=False):
        """Debug buffers of the last instrumented update.  The library copies debug_extent()[0] workgroups, so the
        buffers are sized from that (`nwg` can only make them larger)."""
        have = self.debug_extent()[0]
        nwg = have if nwg is None else max(int(nwg), have)      # never smaller than what the library copies
        pre = np.zeros((nwg, 3), np.float32) if pre_blend else None
        rs = np.zeros(nwg * self.rays, np.uint32) if rays else None
        rh = np.zeros(nwg * self.rays, np.uint8) if rays else None
        tot = (C.c_uint64 * 3)()
        self._ck(self.lib.lumcu_download_debug(self.ctx, _ptr(pre), _ptr(rs), _ptr(rh), tot))
        return dict(pre_blend=pre, ray_samples=rs, ray_hits=rh, totals=(int(tot[0]), int(tot[1]), int(tot[2])))

    def debug_unpack_table(self):
        out = np.zeros(1024, np.float32)
        self._ck(self.lib.lumcu_debug_unpack_table(self.ctx, _ptr(out)))
        return out

    def wait_idle(self):
        self._ck(self.lib.lumcu_wait_idle(self.ctx))

    def last_update_ms(self):
        ms = C.c_float()
        self._ck(self.lib.lumcu_last_update_ms(self.ctx, C.byref(ms)))
        return ms.value

    def set_kernel_variant(self, variant):
        self._ck(self.lib.lumcu_set_kernel_variant(self.ctx, int(variant)))

    def set_shard_mode(self, mode):
        """SHARD_INTERLEAVED (default) or SHARD_CONTIGUOUS (lumcu_set_shard_mode)."""
        self._ck(self.lib.lumcu_set_shard_mode(self.ctx, int(mode)))

    def shard_chunk(self, size, nranks):
        ch = C.c_int32()
        self._ck(self.lib.lumcu_shard_chunk(self.ctx, int(size), int(nranks), C.byref(ch)))
        return ch.value

    def set_profiling(self, on):
        self._ck(self.lib.lumcu_set_profiling(self.ctx, int(bool(on))))

    def last_kernel_ms(self):
        ms = C.c_float()
        self._ck(self.lib.lumcu_last_kernel_ms(self.ctx, C.byref(ms)))
        return ms.value

    def launch_count(self):
        n = C.c_uint64()
        self._ck(self.lib.lumcu_launch_count(self.ctx, C.byref(n)))
        return n.value

    def field_device_ptr(self):
        p = C.c_uint64()
        self._ck(self.lib.lumcu_field_device_ptr(self.ctx, C.byref(p)))
        return p.value

    def stream(self):
        p = C.c_uint64()
        self._ck(self.lib.lumcu_stream(self.ctx, C.byref(p)))
        return p.value
