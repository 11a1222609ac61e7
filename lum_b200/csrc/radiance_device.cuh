// radiance_device.cuh -- device-side arithmetic of the radiance-probe update.
//
// Restates shaders/radiance.comp of the reference (platonvin/lum) for sm_100a.  The sample
// sequence of CastRay_fast (radiance.comp:183-291) is reproduced exactly -- it is a 0.5-voxel
// fixed-step sampler with empty-cell skipping and a <=5-step DDA refinement, not a true DDA --
// so that results are bit-identical to the CPU oracle.  Arithmetic rules (DESIGN.md "Oracle
// semantics"): binary32, one rounding per written operation (this translation unit is compiled
// with -fmad=false; '/' and sqrtf are IEEE under nvcc's default -prec-div/-prec-sqrt), GLSL
// min/max as `y<x?y:x` / `x<y?y:x`, float->int = cvt.rzi (truncate, saturate, NaN->0).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace lumcu {

struct SceneView {
    const int16_t *__restrict__ world;      // [z][y][x] block ids
    const uint8_t *__restrict__ bricks;     // [id][x][y][z] material ids
    const float *__restrict__ materials;    // [256][6]
    const uint32_t *__restrict__ field;     // [z][y][x] packed A2B10G10R10 (read side)
    int wx, wy, wz;
    int capacity;                           // bricks addressable
    int brick0_empty;                       // brick 0 is all zeros (always, unless the user loaded a block 0)
};

struct f3 { float x, y, z; };
struct i3 { int x, y, z; };

__device__ __forceinline__ float gmin(float x, float y) { return (y < x) ? y : x; }
__device__ __forceinline__ float gmax(float x, float y) { return (x < y) ? y : x; }
__device__ __forceinline__ int f2i(float f) { return __float2int_rz(f); }

// radiance.comp:51-59
__device__ __forceinline__ f3 hash33(uint32_t x, uint32_t y, uint32_t z) {
    const uint32_t k = 1103515245u;
#pragma unroll
    for (int r = 0; r < 3; r++) {
        uint32_t a = ((x >> 8) ^ y) * k;
        uint32_t b = ((y >> 8) ^ z) * k;
        uint32_t c = ((z >> 8) ^ x) * k;
        x = a; y = b; z = c;
    }
    const float scale = 1.0f / 4294967296.0f;   // 1.0/float(0xffffffffU)
    return {(float)x * scale, (float)y * scale, (float)z * scale};
}

// OpSDiv by 16 (truncating) for any sign; OpSMod 16 is `& 15` (radiance.comp:124-125).
__device__ __forceinline__ int sdiv16(int a) { return (a + ((a >> 31) & 15)) >> 4; }

// GetVoxel (radiance.comp:121-143).  Out-of-range texel fetches return 0.
__device__ __forceinline__ int get_voxel(const SceneView &s, int ix, int iy, int iz, int &block_id) {
    int bx = sdiv16(ix), by = sdiv16(iy), bz = sdiv16(iz);
    int id = 0;
    if ((unsigned)bx < (unsigned)s.wx && (unsigned)by < (unsigned)s.wy && (unsigned)bz < (unsigned)s.wz)
        id = __ldg(s.world + (bx + s.wx * (by + s.wy * bz)));
    int v = 0;
    if ((unsigned)id < (unsigned)s.capacity && !(id == 0 && s.brick0_empty))
        v = __ldg(s.bricks + ((size_t)id * 4096u + (size_t)(((ix & 15) << 8) | ((iy & 15) << 4) | (iz & 15))));
    block_id = id;
    return v;
}

// c/1023 correctly rounded without the division subroutine: q0 = c*RN(1/1023), one Markstein
// correction step with explicit FMAs.  Bit-identical to IEEE c/1023.0f for every c in [0,1023]
// (verified exhaustively; tests/test_parity_gpu.py::test_unpack_all_codes).
__device__ __forceinline__ float unorm10_to_float(uint32_t c) {
    const float r = 1.0f / 1023.0f;
    float fc = (float)c;
    float q0 = fc * r;
    float rem = __fmaf_rn(-1023.0f, q0, fc);
    return __fmaf_rn(rem, r, q0);
}
__device__ __forceinline__ f3 unpack10(uint32_t w) {
    return {unorm10_to_float(w & 1023u), unorm10_to_float((w >> 10) & 1023u), unorm10_to_float((w >> 20) & 1023u)};
}
__device__ __forceinline__ uint32_t unorm10(float v) {
    float c = gmin(gmax(v, 0.0f), 1.0f);
    if (!(c == c)) c = 0.0f;
    return (uint32_t)__float2int_rn(c * 1023.0f);
}
__device__ __forceinline__ uint32_t pack10(float r, float g, float b) {
    return unorm10(r) | (unorm10(g) << 10) | (unorm10(b) << 20) | (3u << 30);
}

// REPEAT wrap of a texel index.  Hits lie inside the world, so i is within one period of [0,n)
// except for degenerate rays; the integer modulo only runs for those.
__device__ __forceinline__ int wrapi(int i, int n) {
    if (i < 0) i += n; else if (i >= n) i -= n;
    if ((unsigned)i >= (unsigned)n) { i %= n; if (i < 0) i += n; }
    return i;
}
__device__ __forceinline__ f3 lerp3(f3 a, f3 b, float w) {
    return {a.x + w * (b.x - a.x), a.y + w * (b.y - a.y), a.z + w * (b.z - a.z)};
}

// sample_radiance (radiance.comp:161-166): LINEAR + REPEAT gather at block-space bp.
__device__ __forceinline__ f3 gather_field(const SceneView &s, f3 bp) {
    float fx = floorf(bp.x), fy = floorf(bp.y), fz = floorf(bp.z);
    float wx = bp.x - fx, wy = bp.y - fy, wz = bp.z - fz;
    int x0 = wrapi(f2i(fx), s.wx), y0 = wrapi(f2i(fy), s.wy), z0 = wrapi(f2i(fz), s.wz);
    int x1 = x0 + 1 == s.wx ? 0 : x0 + 1;
    int y1 = y0 + 1 == s.wy ? 0 : y0 + 1;
    int z1 = z0 + 1 == s.wz ? 0 : z0 + 1;
    const uint32_t *f = s.field;
    int r00 = s.wx * (y0 + s.wy * z0), r10 = s.wx * (y1 + s.wy * z0);
    int r01 = s.wx * (y0 + s.wy * z1), r11 = s.wx * (y1 + s.wy * z1);
    uint32_t t000 = __ldg(f + r00 + x0), t100 = __ldg(f + r00 + x1);
    uint32_t t010 = __ldg(f + r10 + x0), t110 = __ldg(f + r10 + x1);
    uint32_t t001 = __ldg(f + r01 + x0), t101 = __ldg(f + r01 + x1);
    uint32_t t011 = __ldg(f + r11 + x0), t111 = __ldg(f + r11 + x1);
    f3 c00 = lerp3(unpack10(t000), unpack10(t100), wx);
    f3 c10 = lerp3(unpack10(t010), unpack10(t110), wx);
    f3 c01 = lerp3(unpack10(t001), unpack10(t101), wx);
    f3 c11 = lerp3(unpack10(t011), unpack10(t111), wx);
    return lerp3(lerp3(c00, c10, wy), lerp3(c01, c11, wy), wz);
}

struct RayResult {
    f3 light;
    int samples;   // marching samples (coarse + refinement GetVoxel calls)
    int hit;
};

// ---- ray setup (radiance.comp:370-381) ---------------------------------------------------
// gid_x = gl_GlobalInvocationID.x, ray = gid_x % R, fib = direction table entry of `ray`.
// Outputs origin o, direction d and inv = 1/d (radiance.comp:198).
__device__ __forceinline__ void setup_ray(int px, int py, int pz, uint32_t gid_x, int ray, f3 fib,
                                          int time_seed, f3 &o, f3 &d, f3 &inv) {
    uint32_t sx = gid_x + (uint32_t)time_seed, sy = (uint32_t)time_seed, sz = (uint32_t)time_seed;
    f3 r1 = hash33(sx, sy, sz);
    f3 r2 = hash33(sx + 1u, sy, sz);
    f3 j = {fib.x + 0.1f * ((r1.x - 0.5f) * 2.0f), fib.y + 0.1f * ((r1.y - 0.5f) * 2.0f),
            fib.z + 0.1f * ((r1.z - 0.5f) * 2.0f)};
    float len = sqrtf(j.x * j.x + j.y * j.y + j.z * j.z);
    d = {j.x / len, j.y / len, j.z / len};
    o = {(float)px * 16.0f + 8.0f, (float)py * 16.0f + 8.0f, (float)pz * 16.0f + 8.0f};
    o.x += d.x * 6.0f; o.y += d.y * 6.0f; o.z += d.z * 6.0f;
    o.x += (r2.x - 0.5f) * 2.0f; o.y += (r2.y - 0.5f) * 2.0f; o.z += (r2.z - 0.5f) * 2.0f;
    if (ray & 1) { d.x = -d.x; d.y = -d.y; d.z = -d.z; }
    inv = {1.0f / d.x, 1.0f / d.y, 1.0f / d.z};
}

// ---- one iteration of the coarse march (radiance.comp:205-244) ----------------------------
// Advances t; returns 0 = keep marching, 1 = hit (voxel != 0), 2 = miss (left the world or t >= 128).
struct MarchConsts { float wsx, wsy, wsz; };
__device__ __forceinline__ int march_step(const SceneView &s, const MarchConsts &mc, const f3 &o, const f3 &d,
                                          const f3 &inv, float &t) {
    t += 0.5f;
    f3 p = {o.x + d.x * t, o.y + d.y * t, o.z + d.z * t};
    int ix = f2i(p.x), iy = f2i(p.y), iz = f2i(p.z);
    int id;
    int voxel = get_voxel(s, ix, iy, iz, id);
    // Branch-free tail: the empty-cell skip is computed by every lane and selected.  (Measured on
    // B200: the compiler's two divergent paths here cost 7 % of the whole kernel.)
    const float bx = (float)(sdiv16(ix) * 16) + (d.x > 0.0f ? 16.0f : 0.0f);
    const float by = (float)(sdiv16(iy) * 16) + (d.y > 0.0f ? 16.0f : 0.0f);
    const float bz = (float)(sdiv16(iz) * 16) + (d.z > 0.0f ? 16.0f : 0.0f);
    const float ex = bx * inv.x + (-p.x * inv.x);
    const float ey = by * inv.y + (-p.y * inv.y);
    const float ez = bz * inv.z + (-p.z * inv.z);
    const float tn = t + gmax(gmin(ex, gmin(ey, ez)), 0.01f);
    t = (id == 0) ? tn : t;
    const bool out = p.x < 0.0f || p.y < 0.0f || p.z < 0.0f || p.x >= mc.wsx || p.y >= mc.wsy || p.z >= mc.wsz;
    return voxel != 0 ? 1 : ((out || t >= 128.0f) ? 2 : 0);
}

// ---- after the march: refinement + hit shading, or sky (radiance.comp:246-288,316-347) ----
// t = fraction at loop exit; hit = CastRay_fast's return value.  refine_samples counts GetVoxel calls.
__device__ __forceinline__ f3 finish_ray(const SceneView &s, const f3 &o, const f3 &d, const f3 &inv, float t,
                                         bool hit, f3 sun, int &refine_samples) {
    refine_samples = 0;
    if (hit) {
        float tb = t - 1.5f;
        f3 h = {o.x + d.x * tb, o.y + d.y * tb, o.z + d.z * tb};
        int stx = d.x > 0.0f ? 1 : -1, sty = d.y > 0.0f ? 1 : -1, stz = d.z > 0.0f ? 1 : -1;
        float c1x = (floorf(h.x) - h.x) / d.x, c1y = (floorf(h.y) - h.y) / d.y, c1z = (floorf(h.z) - h.z) / d.z;
        f3 tMax = {gmax(c1x, c1x + inv.x), gmax(c1y, c1y + inv.y), gmax(c1z, c1z + inv.z)};
        f3 tDelta = {fabsf(inv.x), fabsf(inv.y), fabsf(inv.z)};   // 1/|d| == |1/d| exactly
        int vx = f2i(h.x), vy = f2i(h.y), vz = f2i(h.z);
        float ax = 0.0f, ay = 0.0f, az = 0.0f;
        int id;
        int vid = get_voxel(s, vx, vy, vz, id);
        int ns = 1;
        int it = 0;
        while ((it++ <= 4) && (vid == 0)) {
            ax = ((tMax.x <= tMax.y) && (tMax.x <= tMax.z)) ? 1.0f : 0.0f;
            ay = ((tMax.x >= tMax.y) && (tMax.y <= tMax.z)) ? 1.0f : 0.0f;
            az = ((tMax.z <= tMax.x) && (tMax.z <= tMax.y)) ? 1.0f : 0.0f;
            vx += stx * (int)ax; vy += sty * (int)ay; vz += stz * (int)az;
            tMax.x += tDelta.x * ax; tMax.y += tDelta.y * ay; tMax.z += tDelta.z * az;
            vid = get_voxel(s, vx, vy, vz, id);
            ns++;
        }
        refine_samples = ns;
        f3 n = {-((float)stx * ax), -((float)sty * ay), -((float)stz * az)};
        float tf = (tMax.x - tDelta.x) * ax + (tMax.y - tDelta.y) * ay + (tMax.z - tDelta.z) * az;
        const float *m = s.materials + 6 * (vid & 255);
        float mr = __ldg(m), mg = __ldg(m + 1), mb = __ldg(m + 2), me = __ldg(m + 4);
        f3 x = {o.x + tf * d.x, o.y + tf * d.y, o.z + tf * d.z};
        f3 bp = {(x.x + n.x * 16.0f) / 16.0f, (x.y + n.y * 16.0f) / 16.0f, (x.z + n.z * 16.0f) / 16.0f};
        f3 g = gather_field(s, bp);
        return {0.0f + (1.0f * g.x * mr + mr * me), 0.0f + (1.0f * g.y * mg + mg * me),
                0.0f + (1.0f * g.z * mb + mb * me)};
    }
    float glp = -(d.x * sun.x + d.y * sun.y + d.z * sun.z);
    f3 l = {0.0f, 0.0f, 0.0f};
    if (glp > 0.95f) {
        l.x += 1.0f * glp * (0.9f * 3.0f);
        l.y += 1.0f * glp * (0.9f * 3.0f);
        l.z += 1.0f * glp * (0.6f * 3.0f);
    }
    l.x += 1.0f * (0.53f * 0.1f);
    l.y += 1.0f * (0.81f * 0.1f);
    l.z += 1.0f * (0.92f * 0.1f);
    return l;
}

// Whole ray, one thread (the one-workgroup-per-CTA kernel).
__device__ __forceinline__ RayResult trace_probe_ray(const SceneView &s, int px, int py, int pz,
                                                     uint32_t gid_x, int ray, f3 fib, int time_seed,
                                                     f3 sun) {
    f3 o, d, inv;
    setup_ray(px, py, pz, gid_x, ray, fib, time_seed, o, d, inv);
    const MarchConsts mc = {(float)s.wx * 16.0f, (float)s.wy * 16.0f, (float)s.wz * 16.0f};
    float t = 0.0f;
    int samples = 0, st;
    do { st = march_step(s, mc, o, d, inv, t); samples++; } while (st == 0);
    __syncwarp();   // reconverge before the long tail (early leavers would run it in separate passes)
    RayResult res;
    int nr;
    res.light = finish_ray(s, o, d, inv, t, st == 1, sun, nr);
    res.hit = st == 1;
    res.samples = samples + nr;
    return res;
}

}  // namespace lumcu
