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
    const float4 *__restrict__ mat4;        // [256] {r, g, b, emmit}: the four floats GetMat's callers use, one 16-byte load
    const uint32_t *__restrict__ field;     // [z][y][x] packed A2B10G10R10 (read side)
    int wx, wy, wz;
    int capacity;                           // bricks addressable
    int brick0_empty;                       // brick 0 is all zeros (always, unless the user loaded a block 0)
    const uint32_t *__restrict__ flags;     // bit 0: the world holds ids outside [0, capacity) (validate_world_kernel)
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

// c/1023 correctly rounded without the division subroutine: with r = RN(1/1023) and rl = RN(1/1023 - r),
// c/1023 = c*r + c*rl up to a relative 2^-48, and fma(c, r, RN(c*rl)) rounds that sum once.  Bit-identical to
// IEEE c/1023.0f for every c in [0,1023] (verified exhaustively; tests/test_parity_gpu.py::test_unpack_all_codes).
__device__ __forceinline__ float unorm10_to_float(uint32_t c) {
    const float r = 1.0f / 1023.0f;
    const float rl = __uint_as_float(0x2b802008u);      // RN(1/1023 - r) = 0x1.00401p-40
    const float fc = (float)c;
    return __fmaf_rn(fc, r, fc * rl);
}
// The G and B fields are converted where they sit: float(w & (1023 << k)) = c * 2^k exactly (10 significant
// bits), and scaling r and rl by 2^-k is exact too, so the same fma yields the same bits without the shift.
template <int K>
__device__ __forceinline__ float unorm10_field(uint32_t w) {
    const float r = __uint_as_float(0x3a802008u - ((uint32_t)K << 23));     // RN(1/1023) * 2^-K
    const float rl = __uint_as_float(0x2b802008u - ((uint32_t)K << 23));    // RN(1/1023 - r) * 2^-K
    const float f = (float)(w & (1023u << K));
    return __fmaf_rn(f, r, f * rl);
}
__device__ __forceinline__ f3 unpack10(uint32_t w) {
    return {unorm10_field<0>(w), unorm10_field<10>(w), unorm10_field<20>(w)};
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
// r1, r2 = the two Random3D() draws of the ray (radiance.comp:60-66), already computed.
__device__ __forceinline__ void setup_ray_from_draws(int px, int py, int pz, int ray, f3 fib, f3 r1, f3 r2,
                                                     f3 &o, f3 &d, f3 &inv) {
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
__device__ __forceinline__ void setup_ray(int px, int py, int pz, uint32_t gid_x, int ray, f3 fib,
                                          int time_seed, f3 &o, f3 &d, f3 &inv) {
    uint32_t sx = gid_x + (uint32_t)time_seed, sy = (uint32_t)time_seed, sz = (uint32_t)time_seed;
    f3 r1 = hash33(sx, sy, sz);
    f3 r2 = hash33(sx + 1u, sy, sz);
    setup_ray_from_draws(px, py, pz, ray, fib, r1, r2, o, d, inv);
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
    return voxel != 0 ? 1 : ((out || !(t < 128.0f)) ? 2 : 0);   // !(t<128): a NaN fraction ends the ray (decree 13)
}

// 16*b + c as ONE multiply-add on the FMA pipe (left to itself the compiler re-derives 16*b as `i & ~15` on the
// ALU pipe, the one this loop is short of, and adds separately).
// (The #else branches exist for tests/device_host, which compiles this header with g++ to run the device
// arithmetic against the oracle where there is no GPU; nvcc's device pass never sees them.)
__device__ __forceinline__ int mad16(int b, int c) {
#ifdef __CUDA_ARCH__
    int r;
    asm("mad.lo.s32 %0, %1, 16, %2;" : "=r"(r) : "r"(b), "r"(c));
    return r;
#else
    return (int)((unsigned)b * 16u + (unsigned)c);
#endif
}
__device__ __forceinline__ int msub16(int b, int c) {     // c - 16*b
#ifdef __CUDA_ARCH__
    int r;
    asm("mad.lo.s32 %0, %1, -16, %2;" : "=r"(r) : "r"(b), "r"(c));
    return r;
#else
    return (int)((unsigned)c - (unsigned)b * 16u);
#endif
}

// ---- the same iteration for the validated device mirror -----------------------------------
// Preconditions (checked on the device, see SceneView::flags): brick 0 is empty and every world id is
// in [0, capacity).  Differences from march_step, none of which can change a result:
//   * `out` first, as ONE unsigned compare per axis on the float bits: negative floats compare above
//     every positive bound, and pos is never -0.0 or NaN here (origin + dir*t with a finite t and an
//     origin that is a sum with a non-zero first term), so this is exactly pos<0 || pos>=16*world_size;
//   * an `out` lane does not fetch: it ends as "pending" and the one sample the reference still takes
//     there (trunc-division aliases pos in (-1,0) onto voxel row 0) is taken by
//     resolve_pending() in the shading phase, with the general get_voxel;
//   * in-world lanes need no bounds test, no negative-coordinate division and no capacity test;
//   * the cell corner float(16*b + {0,16}) is formed on the FMA pipe: bits = 0x4B000000 + 16*b + off is the
//     float 2^23 + corner, minus 2^23 is exact.  cm.* = 0x4B000000 + (dir > 0 ? 16 : 0), per ray.
// The loop of the pooled kernel is bound by the ALU pipe (ncu: 65 % of its peak, every other pipe < 25 %),
// so the integer work is written as multiply-adds (IMAD, FMA-heavy pipe) wherever that is exact.
// t_pre = the fraction at which pos was sampled (what a pending lane stores).
__device__ __forceinline__ bool march_step_fast(const SceneView &s, const uint3 &wsbits, const f3 &o, const f3 &d,
                                                const f3 &inv, const i3 &cm, float &t, float &t_pre, bool &hit) {
    t += 0.5f;
    t_pre = t;
    const float px = o.x + d.x * t, py = o.y + d.y * t, pz = o.z + d.z * t;
    const bool out = __float_as_uint(px) >= wsbits.x || __float_as_uint(py) >= wsbits.y || __float_as_uint(pz) >= wsbits.z;
    const int ix = f2i(px), iy = f2i(py), iz = f2i(pz);
    const int bx = ix >> 4, by = iy >> 4, bz = iz >> 4;
    // Addresses are formed by every lane and only the loads are predicated (no divergent branch in the loop).
    const int16_t *wp = s.world + (bx + s.wx * (by + s.wy * bz));
    int id = 0;
    if (!out) id = __ldg(wp);
    // ((ix&15)<<8 | (iy&15)<<4 | iz&15) == V - 16*B with V = (ix*16+iy)*16+iz, B = (bx*16+by)*16+bz
    const int V = mad16(mad16(ix, iy), iz), B = mad16(mad16(bx, by), bz);
    const uint8_t *bp = s.bricks + (uint32_t)(id * 4096 + msub16(B, V));
    int voxel = 0;
    if (id != 0) voxel = __ldg(bp);
    const float cx = __int_as_float(mad16(bx, cm.x)) - 8388608.0f;
    const float cy = __int_as_float(mad16(by, cm.y)) - 8388608.0f;
    const float cz = __int_as_float(mad16(bz, cm.z)) - 8388608.0f;
    const float ex = cx * inv.x + (-px * inv.x);
    const float ey = cy * inv.y + (-py * inv.y);
    const float ez = cz * inv.z + (-pz * inv.z);
    const float tn = t + gmax(gmin(ex, gmin(ey, ez)), 0.01f);
    t = (id == 0) ? tn : t;
    hit = voxel != 0;
    return hit || out || !(t < 128.0f);
}

// ---- two iterations at once, the second one's block-id fetch issued early (the loop body of every fast march) ----
// A step is two dependent fetches (block id, then voxel), so two steps are four fetch latencies in a row.  But the
// position of step B depends on step A's block id only (the empty-cell skip), not on A's voxel: B's block id can be
// requested while A's voxel is still in flight -- id_A -> (voxel_A | id_B) -> voxel_B, three latencies instead of
// four.  The early fetch is speculative only in that a ray ending at A does not need it; its address is in-world by
// B's own bounds test, and a lane that is done at A would idle through B under lock-step anyway.  Every value a
// continuing ray sees is computed exactly as by two calls of march_step_fast.  steps = iterations the ray took.
__device__ __forceinline__ bool march_two_steps_fast(const SceneView &s, const uint3 &wsbits, const f3 &o, const f3 &d,
                                                     const f3 &inv, const i3 &cm, float &t, float &t_pre, bool &hit,
                                                     int &steps) {
    // step A up to its block id
    const float tA = t + 0.5f;
    const float pxA = o.x + d.x * tA, pyA = o.y + d.y * tA, pzA = o.z + d.z * tA;
    const bool outA = __float_as_uint(pxA) >= wsbits.x || __float_as_uint(pyA) >= wsbits.y || __float_as_uint(pzA) >= wsbits.z;
    const int ixA = f2i(pxA), iyA = f2i(pyA), izA = f2i(pzA);
    const int bxA = ixA >> 4, byA = iyA >> 4, bzA = izA >> 4;
    int idA = 0;
    if (!outA) idA = __ldg(s.world + (bxA + s.wx * (byA + s.wy * bzA)));
    const int VA = mad16(mad16(ixA, iyA), izA), BA = mad16(mad16(bxA, byA), bzA);
    int voxA = 0;
    if (idA != 0) voxA = __ldg(s.bricks + (uint32_t)(idA * 4096 + msub16(BA, VA)));
    const float cxA = __int_as_float(mad16(bxA, cm.x)) - 8388608.0f;
    const float cyA = __int_as_float(mad16(byA, cm.y)) - 8388608.0f;
    const float czA = __int_as_float(mad16(bzA, cm.z)) - 8388608.0f;
    const float exA = cxA * inv.x + (-pxA * inv.x);
    const float eyA = cyA * inv.y + (-pyA * inv.y);
    const float ezA = czA * inv.z + (-pzA * inv.z);
    const float tnA = tA + gmax(gmin(exA, gmin(eyA, ezA)), 0.01f);
    const float tA2 = (idA == 0) ? tnA : tA;
    // step B up to its block id: needs tA2 (A's block id), not A's voxel
    const float tB = tA2 + 0.5f;
    const float pxB = o.x + d.x * tB, pyB = o.y + d.y * tB, pzB = o.z + d.z * tB;
    const bool outB = __float_as_uint(pxB) >= wsbits.x || __float_as_uint(pyB) >= wsbits.y || __float_as_uint(pzB) >= wsbits.z;
    const int ixB = f2i(pxB), iyB = f2i(pyB), izB = f2i(pzB);
    const int bxB = ixB >> 4, byB = iyB >> 4, bzB = izB >> 4;
    int idB = 0;
    if (!outB) idB = __ldg(s.world + (bxB + s.wx * (byB + s.wy * bzB)));
    // A ends here
    const bool hitA = voxA != 0;
    const bool doneA = hitA || outA || !(tA2 < 128.0f);
    // rest of step B (its voxel is only fetched for a ray that got this far)
    const int VB = mad16(mad16(ixB, iyB), izB), BB = mad16(mad16(bxB, byB), bzB);
    int voxB = 0;
    if (idB != 0 && !doneA) voxB = __ldg(s.bricks + (uint32_t)(idB * 4096 + msub16(BB, VB)));
    const float cxB = __int_as_float(mad16(bxB, cm.x)) - 8388608.0f;
    const float cyB = __int_as_float(mad16(byB, cm.y)) - 8388608.0f;
    const float czB = __int_as_float(mad16(bzB, cm.z)) - 8388608.0f;
    const float exB = cxB * inv.x + (-pxB * inv.x);
    const float eyB = cyB * inv.y + (-pyB * inv.y);
    const float ezB = czB * inv.z + (-pzB * inv.z);
    const float tnB = tB + gmax(gmin(exB, gmin(eyB, ezB)), 0.01f);
    const float tB2 = (idB == 0) ? tnB : tB;
    const bool hitB = voxB != 0;
    const bool doneB = hitB || outB || !(tB2 < 128.0f);
    t = doneA ? tA2 : tB2;
    t_pre = doneA ? tA : tB;
    hit = doneA ? hitA : hitB;
    steps = doneA ? 1 : 2;
    return doneA || doneB;
}

// A ray that ended without an in-world hit stored -t_pre.  If its last position was outside the world the
// reference still sampled there before the bounds test (radiance.comp:212-237): take that one sample with
// the general get_voxel.  Returns true if it hit after all; the fraction is then t_pre itself (brick 0 is
// empty under the fast path's precondition, so a solid voxel implies block id != 0, i.e. no skip was added).
__device__ __forceinline__ bool resolve_pending(const SceneView &s, const MarchConsts &mc, const f3 &o, const f3 &d,
                                                float t_pre) {
    const float px = o.x + d.x * t_pre, py = o.y + d.y * t_pre, pz = o.z + d.z * t_pre;
    const bool out = px < 0.0f || py < 0.0f || pz < 0.0f || px >= mc.wsx || py >= mc.wsy || pz >= mc.wsz;
    if (!out) return false;                 // in-world sample: the march already saw voxel == 0
    int id;
    return get_voxel(s, f2i(px), f2i(py), f2i(pz), id) != 0;
}

// ---- refinement after a coarse hit (radiance.comp:246-288 == glossy.frag:182-224) --------------
// t = fraction at loop exit.  Outputs the face normal, the refined fraction and the material id;
// refine_samples counts GetVoxel calls.
// a / d for a = floor(h) - h.  An empty-cell skip lands samples on cell faces, so ~3 % of the hits have an h
// that is an exact integer on one axis, a == +0, and a zero numerator sends nvcc's IEEE division into its
// out-of-line slow path (ncu: 42 warp-instructions per 32 rays at 1.6 active lanes).  PROBE_DIRS: d is a
// component of a probe ray direction -- zero or of magnitude >= 2^-50 by construction (setup_ray), never
// denormal -- so for a == 0 the quotient is (1/d) * 0: +-0 with d's sign, NaN for d == +-0, exactly as 0/d.
template <bool PROBE_DIRS>
__device__ __forceinline__ float div_frac(float a, float d) {
    if constexpr (PROBE_DIRS) {
        const bool z = a == 0.0f;
        const float q = (z ? 1.0f : a) / d;
        return z ? q * 0.0f : q;
    } else {
        return a / d;
    }
}

// GetVoxel for the validated device mirror (brick 0 empty, ids inside the palette): a voxel inside the world
// needs neither the sign-correct division nor the block bounds test nor the capacity test; anything else takes
// the general path.  Same result as get_voxel for every input.
template <bool VALID>
__device__ __forceinline__ int get_voxel_v(const SceneView &s, int ix, int iy, int iz) {
    if constexpr (VALID) {
        if ((unsigned)ix < (unsigned)(s.wx << 4) && (unsigned)iy < (unsigned)(s.wy << 4) && (unsigned)iz < (unsigned)(s.wz << 4)) {
            const int bx = ix >> 4, by = iy >> 4, bz = iz >> 4;
            const int id = __ldg(s.world + (bx + s.wx * (by + s.wy * bz)));
            const int V = (ix * 16 + iy) * 16 + iz, B = (bx * 16 + by) * 16 + bz;
            int v = 0;
            if (id != 0) v = __ldg(s.bricks + (uint32_t)(id * 4096 + (V - 16 * B)));
            return v;
        }
    }
    int id;
    return get_voxel(s, ix, iy, iz, id);
}

template <bool PROBE_DIRS, bool VALID>
__device__ __forceinline__ void refine_hit(const SceneView &s, const f3 &o, const f3 &d, const f3 &inv, float t,
                                           f3 &n, float &tf, int &vid, int &refine_samples) {
    float tb = t - 1.5f;
    f3 h = {o.x + d.x * tb, o.y + d.y * tb, o.z + d.z * tb};
    int stx = d.x > 0.0f ? 1 : -1, sty = d.y > 0.0f ? 1 : -1, stz = d.z > 0.0f ? 1 : -1;
    float c1x = div_frac<PROBE_DIRS>(floorf(h.x) - h.x, d.x), c1y = div_frac<PROBE_DIRS>(floorf(h.y) - h.y, d.y),
          c1z = div_frac<PROBE_DIRS>(floorf(h.z) - h.z, d.z);
    f3 tMax = {gmax(c1x, c1x + inv.x), gmax(c1y, c1y + inv.y), gmax(c1z, c1z + inv.z)};
    f3 tDelta = {fabsf(inv.x), fabsf(inv.y), fabsf(inv.z)};   // 1/|d| == |1/d| exactly
    int vx = f2i(h.x), vy = f2i(h.y), vz = f2i(h.z);
    float ax = 0.0f, ay = 0.0f, az = 0.0f;
    vid = get_voxel_v<VALID>(s, vx, vy, vz);
    int ns = 1;
    int it = 0;
    while ((it++ <= 4) && (vid == 0)) {
        ax = ((tMax.x <= tMax.y) && (tMax.x <= tMax.z)) ? 1.0f : 0.0f;
        ay = ((tMax.x >= tMax.y) && (tMax.y <= tMax.z)) ? 1.0f : 0.0f;
        az = ((tMax.z <= tMax.x) && (tMax.z <= tMax.y)) ? 1.0f : 0.0f;
        vx += stx * (int)ax; vy += sty * (int)ay; vz += stz * (int)az;
        tMax.x += tDelta.x * ax; tMax.y += tDelta.y * ay; tMax.z += tDelta.z * az;
        vid = get_voxel_v<VALID>(s, vx, vy, vz);
        ns++;
    }
    refine_samples = ns;
    n = {-((float)stx * ax), -((float)sty * ay), -((float)stz * az)};
    tf = (tMax.x - tDelta.x) * ax + (tMax.y - tDelta.y) * ay + (tMax.z - tDelta.z) * az;
}

// ---- after the march: refinement + hit shading, or sky (radiance.comp:246-288,316-347) ----
// t = fraction at loop exit; hit = CastRay_fast's return value.  refine_samples counts GetVoxel calls.
template <bool VALID>
__device__ __forceinline__ f3 finish_ray(const SceneView &s, const f3 &o, const f3 &d, const f3 &inv, float t,
                                         bool hit, f3 sun, int &refine_samples) {
    refine_samples = 0;
    if (hit) {
        f3 n;
        float tf;
        int vid;
        refine_hit<true, VALID>(s, o, d, inv, t, n, tf, vid, refine_samples);
        const float4 m = __ldg(s.mat4 + (vid & 255));     // GetMat (radiance.comp:145-157): r, g, b, emmit
        const float mr = m.x, mg = m.y, mb = m.z, me = m.w;
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

// ---- CastRay_fast's coarse loop for one ray, start to end (one-thread-per-ray kernels) ----------
// fast = the scene passed validation (see march_step_fast).  Returns hit; t = fraction at loop exit.
__device__ __forceinline__ bool march_ray(const SceneView &s, const MarchConsts &mc, bool fast, const f3 &o, const f3 &d,
                                          const f3 &inv, float &t) {
    t = 0.0f;
    if (fast) {
        const uint3 wsbits = {__float_as_uint(mc.wsx), __float_as_uint(mc.wsy), __float_as_uint(mc.wsz)};
        const i3 cm = {0x4B000000 + (d.x > 0.0f ? 16 : 0), 0x4B000000 + (d.y > 0.0f ? 16 : 0), 0x4B000000 + (d.z > 0.0f ? 16 : 0)};
        float t_pre;
        bool hit;
        int steps;     // the fused pair of steps here too (demo world: 34 -> 32 us per update, r02 sweep)
        while (!march_two_steps_fast(s, wsbits, o, d, inv, cm, t, t_pre, hit, steps)) {}
        if (hit) return true;
        // Ended without an in-world hit.  Unlike the probe rays, whose origins are built by setup_ray, these
        // origins come from the caller: a position of -0.0 or NaN is "out" for the bit test but not for
        // pos < 0 -- such a ray (never seen in practice) is simply redone by the general loop below.
        const float px = o.x + d.x * t_pre, py = o.y + d.y * t_pre, pz = o.z + d.z * t_pre;
        const bool fast_out = __float_as_uint(px) >= wsbits.x || __float_as_uint(py) >= wsbits.y || __float_as_uint(pz) >= wsbits.z;
        const bool out = px < 0.0f || py < 0.0f || pz < 0.0f || px >= mc.wsx || py >= mc.wsy || pz >= mc.wsz;
        if (fast_out == out) {
            if (!out) return false;                                   // in-world, ended by distance
            int id;
            if (get_voxel(s, f2i(px), f2i(py), f2i(pz), id) == 0) return false;
            t = t_pre;                                                // the deferred sample hit (resolve_pending)
            return true;
        }
        t = 0.0f;
    }
    int st;
    do { st = march_step(s, mc, o, d, inv, t); } while (st == 0);
    return st == 1;
}

// Whole ray, one thread (the one-workgroup-per-CTA kernel).  VALID: the scene passed validation, so the fast
// march and the in-world fetches apply (see march_step_fast).
template <bool VALID>
__device__ __forceinline__ RayResult trace_probe_ray(const SceneView &s, int px, int py, int pz,
                                                     uint32_t gid_x, int ray, f3 fib, int time_seed,
                                                     f3 sun) {
    f3 o, d, inv;
    setup_ray(px, py, pz, gid_x, ray, fib, time_seed, o, d, inv);
    const MarchConsts mc = {(float)s.wx * 16.0f, (float)s.wy * 16.0f, (float)s.wz * 16.0f};
    float t = 0.0f;
    int samples = 0;
    bool hit;
    if constexpr (VALID) {
        const uint3 wsbits = {__float_as_uint(mc.wsx), __float_as_uint(mc.wsy), __float_as_uint(mc.wsz)};
        const i3 cm = {0x4B000000 + (d.x > 0.0f ? 16 : 0), 0x4B000000 + (d.y > 0.0f ? 16 : 0), 0x4B000000 + (d.z > 0.0f ? 16 : 0)};
        float t_pre;
        bool done;
        int steps;
        do { done = march_two_steps_fast(s, wsbits, o, d, inv, cm, t, t_pre, hit, steps); samples += steps; } while (!done);
        if (!hit && resolve_pending(s, mc, o, d, t_pre)) { hit = true; t = t_pre; }
    } else {
        int st;
        do { st = march_step(s, mc, o, d, inv, t); samples++; } while (st == 0);
        hit = st == 1;
    }
    __syncwarp();   // reconverge before the long tail (early leavers would run it in separate passes)
    RayResult res;
    int nr;
    res.light = finish_ray<VALID>(s, o, d, inv, t, hit, sun, nr);
    res.hit = hit;
    res.samples = samples + nr;
    return res;
}

// ---- glossy.frag:439-491 main() for one fragment (the body of glossy_kernel) ------------------
struct GlossyArgs {
    float campos[3], camdir[3], hs[3], vs[3], gld[3];   // LumcuGlossyView
    float fw, fh;
    int width, height;
    const uchar4 *__restrict__ mat_norm;                // [h][w] R8G8B8A8_UINT
    const float *__restrict__ depth;                    // [h][w]
    float4 *__restrict__ frame;                         // [h][w]
    uint8_t *__restrict__ covered;                      // [h][w] or null
};

// The stencil test of the pass (fillStencilGlossy.frag: discard if roughness > 0.5) is evaluated here instead of
// by the ROP.  (px, py) is inside the frame.
__device__ __forceinline__ void glossy_fragment(const SceneView &s, const GlossyArgs &a, int px, int py) {
    const size_t pi = (size_t)py * a.width + px;
    const uchar4 mn = a.mat_norm[pi];
    const float4 m = __ldg(s.mat4 + (int)mn.x);                         // GetMat(load_mat(pix)) :443, :85-96
    const float mr = m.x, mg = m.y, mb = m.z, me = m.w, rough = __ldg(s.materials + 6 * (int)mn.x + 5);
    const bool cov = !(rough > 0.5f);
    if (a.covered) a.covered[pi] = cov ? 1 : 0;
    if (!cov) return;
    const bool fast = s.brick0_empty && !(__ldg(s.flags) & 1u);
    const MarchConsts mc = {(float)s.wx * 16.0f, (float)s.wy * 16.0f, (float)s.wz * 16.0f};

    f3 d = {a.camdir[0], a.camdir[1], a.camdir[2]};                     // :444
    const float fcx = (float)px + 0.5f, fcy = (float)py + 0.5f;         // gl_FragCoord.xy
    const float cx = fcx / a.fw * 2.0f - 1.0f, cy = fcy / a.fh * 2.0f - 1.0f;   // :446
    const float dep = __ldg(a.depth + pi) * 1000.0f;                    // load_depth :377-382 (nearest texel = pix)
    f3 o = {((a.campos[0] + a.hs[0] * cx) + a.vs[0] * cy) + a.camdir[0] * dep,   // get_origin_from_depth :41-47
            ((a.campos[1] + a.hs[1] * cx) + a.vs[1] * cy) + a.camdir[1] * dep,
            ((a.campos[2] + a.hs[2] * cx) + a.vs[2] * cy) + a.camdir[2] * dep};
    f3 n = {(float)mn.y / 255.0f * 2.0f - 1.0f, (float)mn.z / 255.0f * 2.0f - 1.0f, (float)mn.w / 255.0f * 2.0f - 1.0f};  // load_norm :367-371

    // ProcessHit(origin, direction, 0, normal, mat, ...) :454-456, :312-326
    o = {o.x + 0.0f * d.x, o.y + 0.0f * d.y, o.z + 0.0f * d.z};
    f3 g = gather_field(s, {(o.x + n.x * 16.0f) / 16.0f, (o.y + n.y * 16.0f) / 16.0f, (o.z + n.z * 16.0f) / 16.0f});
    f3 refl = {1.0f * mr, 1.0f * mg, 1.0f * mb};
    f3 light = {0.0f + refl.x * (me + g.x), 0.0f + refl.y * (me + g.y), 0.0f + refl.z * (me + g.z)};
    {   // reflect(I, N) = I - 2 * dot(N, I) * N (GLSL 8.5), evaluated left to right (decree 14)
        const float k = 2.0f * ((n.x * d.x + n.y * d.y) + n.z * d.z);
        d = {d.x - k * n.x, d.y - k * n.y, d.z - k * n.z};
    }

    // trace_glossy_ray :328-360
    const f3 inv = {1.0f / d.x, 1.0f / d.y, 1.0f / d.z};
    float t;
    const bool hit = march_ray(s, mc, fast, o, d, inv, t);
    __syncwarp();
    if (hit) {
        f3 n2;
        float tf;
        int vid, nr;
        if (fast) refine_hit<false, true>(s, o, d, inv, t, n2, tf, vid, nr);
        else refine_hit<false, false>(s, o, d, inv, t, n2, tf, vid, nr);
        const float4 m2 = __ldg(s.mat4 + (vid & 255));
        const float r2 = m2.x, g2 = m2.y, b2 = m2.z, e2 = m2.w;
        const f3 x = {o.x + tf * d.x, o.y + tf * d.y, o.z + tf * d.z};
        const f3 gg = gather_field(s, {(x.x + n2.x * 16.0f) / 16.0f, (x.y + n2.y * 16.0f) / 16.0f, (x.z + n2.z * 16.0f) / 16.0f});
        refl = {refl.x * r2, refl.y * g2, refl.z * b2};
        light = {light.x + refl.x * (e2 + gg.x), light.y + refl.y * (e2 + gg.y), light.z + refl.z * (e2 + gg.z)};
    } else {
        const float glp = -((d.x * a.gld[0] + d.y * a.gld[1]) + d.z * a.gld[2]);   // :346
        if (glp > 0.9f) {                                                            // :347-350
            const float w = glp - 0.9f;
            light.x += (0.9f * 0.5f) * refl.x * w * 10.0f;
            light.y += (0.9f * 0.5f) * refl.y * w * 10.0f;
            light.z += (0.6f * 0.5f) * refl.z * w * 10.0f;
        }
        light.x += (0.53f * 0.1f) * refl.x;                                          // :353
        light.y += (0.81f * 0.1f) * refl.y;
        light.z += (0.92f * 0.1f) * refl.z;
    }
    a.frame[pi] = make_float4(light.x / 8.0f, light.y / 8.0f, light.z / 8.0f, 1.0f - rough);   // :490, encode_color :435
}

// ---- shaders/map.comp:32-67 for one invocation (the body of map_model_kernel) ----------------
struct MapArgs {
    float m[16];                 // inverse_trans, column-major (glm mat4)
    int shift[3];
    int ext[3];                  // invocations per axis = 4*ceil(area/4)
    int msx, msy, msz;           // model size
    int static_count;
};

// (gx, gy, gz) = gl_GlobalInvocationID inside the dispatched area.
__device__ __forceinline__ void map_model_invocation(const int16_t *__restrict__ world, uint8_t *__restrict__ bricks,
                                                     const uint8_t *__restrict__ model, int wx, int wy, int wz,
                                                     int capacity, const MapArgs &a, int gx, int gy, int gz) {
    // world_voxel = vec3(gid)/1.0 + vec3(0.5) + vec3(shift)      (:41-43)
    const float px = ((float)gx + 0.5f) + (float)a.shift[0];
    const float py = ((float)gy + 0.5f) + (float)a.shift[1];
    const float pz = ((float)gz + 0.5f) + (float)a.shift[2];
    if (px < 0.0f || py < 0.0f || pz < 0.0f) return;          // decree 11: below-world invocations alias + race in the reference
    const int ix = f2i(px), iy = f2i(py), iz = f2i(pz);       // ivec3(world_voxel) :44
    // model_voxel = (inverse_trans * vec4(world_voxel,1)).xyz   (:50), columns summed 0,1,2,3
    const float mx = ((a.m[0] * px + a.m[4] * py) + a.m[8] * pz) + a.m[12] * 1.0f;
    const float my = ((a.m[1] * px + a.m[5] * py) + a.m[9] * pz) + a.m[13] * 1.0f;
    const float mz = ((a.m[2] * px + a.m[6] * py) + a.m[10] * pz) + a.m[14] * 1.0f;
    const int jx = f2i(mx), jy = f2i(my), jz = f2i(mz);       // ivec3(model_voxel) :51
    const int bx = sdiv16(ix), by = sdiv16(iy), bz = sdiv16(iz);   // iworld_voxel/16 :53
    int id = 0;
    if ((unsigned)bx < (unsigned)wx && (unsigned)by < (unsigned)wy && (unsigned)bz < (unsigned)wz)
        id = world[bx + wx * (by + wy * bz)];
    if (id >= a.static_count && id < capacity) {              // :62 (stores outside the atlas are dropped)
        int v = 0;
        if ((unsigned)jx < (unsigned)a.msx && (unsigned)jy < (unsigned)a.msy && (unsigned)jz < (unsigned)a.msz)
            v = model[jx + a.msx * (jy + a.msy * jz)];
        if (v != 0)
            bricks[(size_t)id * 4096u + (size_t)(((ix & 15) << 8) | ((iy & 15) << 4) | (iz & 15))] = (uint8_t)v;
    }
}

}  // namespace lumcu
