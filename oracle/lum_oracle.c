/*
 * lum_oracle.c -- CPU restatement (ORACLE) of Lum's radiance-probe update path.
 *
 * TEST INFRASTRUCTURE ONLY (see lum_oracle.h).  PARITY PINNED AGAINST THE REFERENCE'S OWN
 * SHADER TEXT, executed on the CPU (oracle/_ref/libref_radiance_glsl.so, recipe in
 * oracle/Makefile): tests/test_ref_glsl.py and tests/golden/ref_glsl_golden.npz.  Driver-defined
 * semantics (filter arithmetic, subgroup sum order, ...) remain decrees -- see lum_oracle.h.
 *
 * Every function follows the statement order of the reference text it cites
 * (paths relative to /root/reference).  Arithmetic rules pinned by decree
 * (DESIGN.md "Oracle semantics"): IEEE binary32, one rounding per written operation,
 * no FMA contraction (build with -ffp-contract=off), GLSL min/max/clamp/mix as
 * defined in the GLSL 4.50 spec section 8.3, float->int conversion truncates toward
 * zero and saturates (NaN -> 0).
 */
#include "lum_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct { float x, y, z; } v3;
typedef struct { int32_t x, y, z; } i3;

/* ---- GLSL built-ins, as the GLSL spec defines them ------------------------------- */
static inline float g_min(float x, float y) { return (y < x) ? y : x; }   /* min(x,y) = y<x ? y : x */
static inline float g_max(float x, float y) { return (x < y) ? y : x; }   /* max(x,y) = x<y ? y : x */
static inline float g_clamp(float x, float lo, float hi) { return g_min(g_max(x, lo), hi); }
static inline float g_mix(float a, float b, float t) { return a * (1.0f - t) + b * t; }
static inline int32_t f2i(float f) {                                       /* ivec3(vec3): trunc, saturating */
    if (!(f == f)) return 0;
    if (f >= 2147483648.0f) return INT32_MAX;
    if (f <= -2147483648.0f) return INT32_MIN;
    return (int32_t)f;
}
static inline float g_dot(v3 a, v3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; } /* summed x->y->z */
static inline v3 g_normalize(v3 v) {
    float l = sqrtf(g_dot(v, v));
    v3 r = { v.x / l, v.y / l, v.z / l };
    return r;
}
/* OpSDiv (trunc) and OpSMod (floor) by 16 -- radiance.comp:124-125, GlslangToSpv.cpp:6797-6804 */
static inline int32_t sdiv16(int32_t a) { return a / 16; }
static inline int32_t smod16(int32_t a) { return a & 15; }

/* ---- radiance.comp:51-59 ----------------------------------------------------------- */
static inline void hash33_u(uint32_t x[3]) {
    const uint32_t k = 1103515245U;
    for (int r = 0; r < 3; r++) {
        uint32_t a = ((x[0] >> 8) ^ x[1]) * k;
        uint32_t b = ((x[1] >> 8) ^ x[2]) * k;
        uint32_t c = ((x[2] >> 8) ^ x[0]) * k;
        x[0] = a; x[1] = b; x[2] = c;
    }
}
static inline v3 hash33_f(const uint32_t s[3]) {
    uint32_t x[3] = { s[0], s[1], s[2] };
    hash33_u(x);
    /* vec3(x) * (1.0/float(0xffffffffU)) : float(0xffffffff) rounds to 2^32 */
    const float scale = 1.0f / 4294967296.0f;
    v3 r = { (float)x[0] * scale, (float)x[1] * scale, (float)x[2] * scale };
    return r;
}
void lum_oracle_hash33(const uint32_t in[3], uint32_t out_u[3], float out_f[3]) {
    uint32_t x[3] = { in[0], in[1], in[2] };
    hash33_u(x);
    out_u[0] = x[0]; out_u[1] = x[1]; out_u[2] = x[2];
    v3 f = hash33_f(in);
    out_f[0] = f.x; out_f[1] = f.y; out_f[2] = f.z;
}

/* ---- radiance.comp:293-314 --------------------------------------------------------- */
void lum_oracle_direction_table(int n, float *out) {
    const float PI = 3.1415926535897932384626433832795f;                 /* :40 */
    const float PHI = sqrtf(5.0f) * 0.5f + 0.5f;                         /* :294 */
    for (int k = 0; k < n; k++) {
        float i = (float)k, fn = (float)n;
        float ab = i * (PHI - 1.0f);
        float mf = ab - floorf(ab);                                      /* madfrac :295 */
        float phi = 2.0f * PI * mf;                                      /* :296 */
        float cosTheta = 1.0f - (2.0f * i + 1.0f) * (1.0f / fn);         /* :297 */
        float sinTheta = sqrtf(g_clamp(1.0f - cosTheta * cosTheta, 0.0f, 1.0f)); /* :298 */
        out[3 * k + 0] = cosf(phi) * sinTheta;                           /* :300-303 */
        out[3 * k + 1] = sinf(phi) * sinTheta;
        out[3 * k + 2] = cosTheta;
    }
}

void lum_oracle_sun_dir(float out[3]) {                                   /* :42 */
    v3 v = { 0.5f, 0.5f, -0.9f };
    v = g_normalize(v);
    out[0] = v.x; out[1] = v.y; out[2] = v.z;
}

/* ---- A2B10G10R10_UNORM_PACK32 (internal_render.hpp:456) ---------------------------- */
static inline uint32_t unorm10(float v) {
    float c = g_clamp(v, 0.0f, 1.0f);
    if (!(c == c)) c = 0.0f;                                             /* NaN -> 0 */
    return (uint32_t)rintf(c * 1023.0f);                                 /* round to nearest even */
}
uint32_t lum_oracle_pack_rgb10a2(const float rgb[3]) {
    return unorm10(rgb[0]) | (unorm10(rgb[1]) << 10) | (unorm10(rgb[2]) << 20) | (3u << 30); /* a = 1 */
}
static inline v3 unpack10(uint32_t w) {
    v3 r = { (float)(w & 1023u) / 1023.0f, (float)((w >> 10) & 1023u) / 1023.0f,
             (float)((w >> 20) & 1023u) / 1023.0f };
    return r;
}
void lum_oracle_unpack_rgb10a2(uint32_t word, float rgb[3]) {
    v3 r = unpack10(word);
    rgb[0] = r.x; rgb[1] = r.y; rgb[2] = r.z;
}

/* ---- GetBlock / voxel_in_palette / GetVoxel (radiance.comp:97-143) ----------------- */
static inline void get_voxel(const LumOracleScene *s, i3 ipos, int *voxel, int *block_id) {
    i3 b = { sdiv16(ipos.x), sdiv16(ipos.y), sdiv16(ipos.z) };           /* :124 */
    i3 r = { smod16(ipos.x), smod16(ipos.y), smod16(ipos.z) };           /* :125 */
    int id = 0;                                                          /* texelFetch out of range -> 0 */
    if (b.x >= 0 && b.x < s->wx && b.y >= 0 && b.y < s->wy && b.z >= 0 && b.z < s->wz)
        id = s->world[(size_t)b.x + (size_t)s->wx * ((size_t)b.y + (size_t)s->wy * (size_t)b.z)];
    int v = 0;
    if (id >= 0 && id < s->palette_capacity)                             /* :105-110,140 */
        v = s->bricks[(size_t)id * 4096u + (size_t)r.x * 256u + (size_t)r.y * 16u + (size_t)r.z];
    *voxel = v; *block_id = id;
}
void lum_oracle_get_voxel(const LumOracleScene *s, int x, int y, int z, int out[2]) {
    i3 p = { x, y, z };
    get_voxel(s, p, &out[0], &out[1]);
}
static inline i3 trunc3(v3 p) { i3 r = { f2i(p.x), f2i(p.y), f2i(p.z) }; return r; }

/* ---- sample_radiance (radiance.comp:161-166), LINEAR + REPEAT sampler -------------- */
static inline int wrap(int i, int n) { int m = i % n; return m < 0 ? m + n : m; }
static inline v3 field_tap(const LumOracleScene *s, const uint32_t *field, int x, int y, int z) {
    return unpack10(field[(size_t)x + (size_t)s->wx * ((size_t)y + (size_t)s->wy * (size_t)z)]);
}
static inline v3 lerp3(v3 a, v3 b, float w) {
    v3 r = { a.x + w * (b.x - a.x), a.y + w * (b.y - a.y), a.z + w * (b.z - a.z) };
    return r;
}
static v3 gather(const LumOracleScene *s, const uint32_t *field, v3 bp) {
    float fx = floorf(bp.x), fy = floorf(bp.y), fz = floorf(bp.z);
    float wx = bp.x - fx, wy = bp.y - fy, wz = bp.z - fz;
    int x0 = wrap(f2i(fx), s->wx), y0 = wrap(f2i(fy), s->wy), z0 = wrap(f2i(fz), s->wz);
    int x1 = wrap(x0 + 1, s->wx), y1 = wrap(y0 + 1, s->wy), z1 = wrap(z0 + 1, s->wz);
    v3 c00 = lerp3(field_tap(s, field, x0, y0, z0), field_tap(s, field, x1, y0, z0), wx);
    v3 c10 = lerp3(field_tap(s, field, x0, y1, z0), field_tap(s, field, x1, y1, z0), wx);
    v3 c01 = lerp3(field_tap(s, field, x0, y0, z1), field_tap(s, field, x1, y0, z1), wx);
    v3 c11 = lerp3(field_tap(s, field, x0, y1, z1), field_tap(s, field, x1, y1, z1), wx);
    v3 c0 = lerp3(c00, c10, wy);
    v3 c1 = lerp3(c01, c11, wy);
    return lerp3(c0, c1, wz);
}
void lum_oracle_gather(const LumOracleScene *s, const uint32_t *field, const float bp[3], float rgb[3]) {
    v3 p = { bp[0], bp[1], bp[2] };
    v3 r = gather(s, field, p);
    rgb[0] = r.x; rgb[1] = r.y; rgb[2] = r.z;
}

/* ---- CastRay_fast (radiance.comp:183-291) ------------------------------------------ */
typedef struct { float color[3]; float emmit; } mat_t;
static int cast_ray_fast(const LumOracleScene *s, v3 origin, v3 dir,
                         float *fraction_out, v3 *normal_out, mat_t *mat_out, int *voxel_out,
                         int *coarse_samples, int *refine_samples) {
    const float max_dist = 16.0f * 8.0f;                                 /* :193 */
    /* :195-196 initial GetVoxel(origin) results are dead (overwritten at :212-213) */
    v3 inv = { 1.0f / dir.x, 1.0f / dir.y, 1.0f / dir.z };               /* :198 */
    v3 corner = { dir.x > 0.0f ? 16.0f : 0.0f, dir.y > 0.0f ? 16.0f : 0.0f, dir.z > 0.0f ? 16.0f : 0.0f }; /* :199-200 */
    float wsx = (float)s->wx * 16.0f, wsy = (float)s->wy * 16.0f, wsz = (float)s->wz * 16.0f;
    float fraction = 0.0f;                                               /* :203 */
    int voxel = 0, block = 0, ns = 0;
    for (;;) {                                                           /* :205 */
        fraction += 0.5f;                                                /* :207 */
        v3 pos = { origin.x + dir.x * fraction, origin.y + dir.y * fraction, origin.z + dir.z * fraction }; /* :209 */
        i3 ipos = trunc3(pos);
        get_voxel(s, ipos, &voxel, &block);                              /* :212-213 (same args twice) */
        ns++;
        if (block == 0) {                                                /* :215 */
            v3 box = { (float)(sdiv16(ipos.x) * 16) + corner.x,          /* :216 */
                       (float)(sdiv16(ipos.y) * 16) + corner.y,
                       (float)(sdiv16(ipos.z) * 16) + corner.z };
            v3 temp = { -pos.x * inv.x, -pos.y * inv.y, -pos.z * inv.z };/* :217 */
            v3 bc = { box.x * inv.x + temp.x, box.y * inv.y + temp.y, box.z * inv.z + temp.z }; /* :218 */
            float _f = g_min(bc.x, g_min(bc.y, bc.z));                   /* :220 */
            float f = g_max(_f, 0.01f);                                  /* :222 */
            fraction += f;                                               /* :225 */
        }
        if (voxel != 0) break;                                           /* :228-231 */
        if (pos.x < 0.0f || pos.y < 0.0f || pos.z < 0.0f ||              /* :233-237 */
            pos.x >= wsx || pos.y >= wsy || pos.z >= wsz) break;
        if (!(fraction < max_dist)) break;                               /* :239-243; NaN-safe form (decree 13): a NaN fraction ends the ray instead of looping forever */
    }
    *coarse_samples = ns;
    *refine_samples = 0;
    *fraction_out = fraction;
    if (voxel == 0) return 0;                                            /* :249,290 */

    v3 before = { origin.x + dir.x * (fraction - 1.5f),                  /* :246 */
                  origin.y + dir.y * (fraction - 1.5f),
                  origin.z + dir.z * (fraction - 1.5f) };
    i3 steps = { dir.x > 0.0f ? 1 : -1, dir.y > 0.0f ? 1 : -1, dir.z > 0.0f ? 1 : -1 }; /* :251-252 */
    v3 fl = { floorf(before.x), floorf(before.y), floorf(before.z) };
    v3 c1 = { (fl.x - before.x) / dir.x, (fl.y - before.y) / dir.y, (fl.z - before.z) / dir.z }; /* :257 */
    v3 c2 = { (fl.x - before.x) / dir.x + 1.0f / dir.x,                  /* :258 */
              (fl.y - before.y) / dir.y + 1.0f / dir.y,
              (fl.z - before.z) / dir.z + 1.0f / dir.z };
    v3 tMax = { g_max(c1.x, c2.x), g_max(c1.y, c2.y), g_max(c1.z, c2.z) }; /* :259-261 */
    v3 tDelta = { 1.0f / fabsf(dir.x), 1.0f / fabsf(dir.y), 1.0f / fabsf(dir.z) }; /* :263 */
    i3 vp = trunc3(before);                                              /* :264 */
    v3 axis = { 0.0f, 0.0f, 0.0f };                                      /* :266 */
    int vid, bid, nr = 0;
    get_voxel(s, vp, &vid, &bid); nr++;                                  /* :267 */
    int iterations = 0;
    while ((iterations++ <= 4) && (vid == 0)) {                          /* :270 */
        axis.x = (float)(int)((tMax.x <= tMax.y) && (tMax.x <= tMax.z)); /* :273 */
        axis.y = (float)(int)((tMax.x >= tMax.y) && (tMax.y <= tMax.z)); /* :274 */
        axis.z = (float)(int)((tMax.z <= tMax.x) && (tMax.z <= tMax.y)); /* :275 */
        vp.x += steps.x * f2i(axis.x);                                   /* :277 */
        vp.y += steps.y * f2i(axis.y);
        vp.z += steps.z * f2i(axis.z);
        tMax.x += tDelta.x * axis.x;                                     /* :278 */
        tMax.y += tDelta.y * axis.y;
        tMax.z += tDelta.z * axis.z;
        get_voxel(s, vp, &vid, &bid); nr++;                              /* :280 */
    }
    normal_out->x = -((float)steps.x * axis.x);                          /* :283 */
    normal_out->y = -((float)steps.y * axis.y);
    normal_out->z = -((float)steps.z * axis.z);
    v3 tFinal = { tMax.x - tDelta.x, tMax.y - tDelta.y, tMax.z - tDelta.z }; /* :284 */
    *fraction_out = g_dot(tFinal, axis);                                 /* :285 */
    const float *m = s->materials + 6 * (size_t)(vid & 255);             /* :287, GetMat :145-157 */
    mat_out->color[0] = m[0]; mat_out->color[1] = m[1]; mat_out->color[2] = m[2];
    mat_out->emmit = m[4];
    *voxel_out = vid;
    *refine_samples = nr;
    return 1;
}

/* ---- main() ray setup (radiance.comp:362-381) + trace_ray (:316-347) ---------------- */
void lum_oracle_trace(const LumOracleScene *s, const uint32_t *field,
                      const float *dir_table, const int16_t probe[3],
                      uint32_t gid_x, int ray_index, int time_seed, LumOracleRay *out) {
    uint32_t rs[3] = { gid_x + (uint32_t)time_seed, (uint32_t)time_seed, (uint32_t)time_seed }; /* :370 */
    v3 corner = { (float)probe[0] * 16.0f, (float)probe[1] * 16.0f, (float)probe[2] * 16.0f };  /* :372 */
    v3 pos = { corner.x + 8.0f, corner.y + 8.0f, corner.z + 8.0f };                             /* :373-374 */
    v3 dir = { dir_table[3 * ray_index], dir_table[3 * ray_index + 1], dir_table[3 * ray_index + 2] }; /* :375 */
    v3 r1 = hash33_f(rs); rs[0]++;                                       /* Random3D :62-66 */
    v3 j = { dir.x + 0.1f * ((r1.x - 0.5f) * 2.0f),                      /* :378 */
             dir.y + 0.1f * ((r1.y - 0.5f) * 2.0f),
             dir.z + 0.1f * ((r1.z - 0.5f) * 2.0f) };
    dir = g_normalize(j);
    pos.x += dir.x * 6.0f; pos.y += dir.y * 6.0f; pos.z += dir.z * 6.0f; /* :379 */
    v3 r2 = hash33_f(rs); rs[0]++;
    pos.x += (r2.x - 0.5f) * 2.0f; pos.y += (r2.y - 0.5f) * 2.0f; pos.z += (r2.z - 0.5f) * 2.0f; /* :380 */
    if ((ray_index % 2) != 0) { dir.x = -dir.x; dir.y = -dir.y; dir.z = -dir.z; }               /* :381 */

    float fraction = 0.0f;
    v3 normal = { 0.0f, 0.0f, 0.0f };
    mat_t mat = { { 0.0f, 0.0f, 0.0f }, 0.0f };
    int voxel = 0, ncoarse = 0, nrefine = 0;
    int hit = cast_ray_fast(s, pos, dir, &fraction, &normal, &mat, &voxel, &ncoarse, &nrefine); /* :329 */
    v3 light = { 0.0f, 0.0f, 0.0f };
    if (hit) {                                                           /* :331-333 */
        v3 p = { pos.x + fraction * dir.x, pos.y + fraction * dir.y, pos.z + fraction * dir.z };
        v3 bp = { (p.x + normal.x * 16.0f) / 16.0f, (p.y + normal.y * 16.0f) / 16.0f,
                  (p.z + normal.z * 16.0f) / 16.0f };                    /* :162 */
        v3 g = gather(s, field, bp);
        light.x += 1.0f * g.x * mat.color[0] + mat.color[0] * mat.emmit;
        light.y += 1.0f * g.y * mat.color[1] + mat.color[1] * mat.emmit;
        light.z += 1.0f * g.z * mat.color[2] + mat.color[2] * mat.emmit;
    } else {                                                             /* :334-344 */
        float sun[3]; lum_oracle_sun_dir(sun);
        v3 sd = { sun[0], sun[1], sun[2] };
        float glp = -g_dot(dir, sd);                                     /* :335 */
        if (glp > 0.95f) {                                               /* :336-338 */
            light.x += 1.0f * glp * (0.9f * 3.0f);
            light.y += 1.0f * glp * (0.9f * 3.0f);
            light.z += 1.0f * glp * (0.6f * 3.0f);
        }
        light.x += 1.0f * (0.53f * 0.1f);                                /* :342 */
        light.y += 1.0f * (0.81f * 0.1f);
        light.z += 1.0f * (0.92f * 0.1f);
    }
    out->light[0] = light.x; out->light[1] = light.y; out->light[2] = light.z;
    out->origin[0] = pos.x; out->origin[1] = pos.y; out->origin[2] = pos.z;
    out->dir[0] = dir.x; out->dir[1] = dir.y; out->dir[2] = dir.z;
    out->fraction = fraction;
    out->normal[0] = normal.x; out->normal[1] = normal.y; out->normal[2] = normal.z;
    out->hit = hit; out->coarse_samples = ncoarse; out->refine_samples = nrefine; out->voxel = voxel;
}

/* ---- workgroup reduction (radiance.comp:386-413) + store_probe (:349-354) ---------- */
/* subgroupAdd over a 32-lane subgroup as an xor butterfly 16,8,4,2,1 (decree 3). */
static float butterfly_sum(const float *v, int n) {
    float a[32], b[32];
    for (int i = 0; i < n; i++) a[i] = v[i];
    for (int k = n / 2; k >= 1; k /= 2) {
        for (int i = 0; i < n; i++) b[i] = a[i] + a[i ^ k];
        memcpy(a, b, sizeof(float) * (size_t)n);
    }
    return a[0];
}

int lum_oracle_update(const LumOracleScene *s, const uint32_t *field_old, uint32_t *field_new,
                      const int16_t *req, int64_t nreq, int time_seed, int size, int shift,
                      float *pre_blend, uint32_t *ray_samples, uint8_t *ray_hits,
                      uint64_t totals[3], int nthreads) {
    return lum_oracle_update_ex(s, field_old, field_new, req, nreq, time_seed, size, shift, pre_blend,
                                ray_samples, ray_hits, NULL, totals, nthreads, 0, -1);
}

int lum_oracle_update_ex(const LumOracleScene *s, const uint32_t *field_old, uint32_t *field_new,
                         const int16_t *req, int64_t nreq, int time_seed, int size, int shift,
                         float *pre_blend, uint32_t *ray_samples, uint8_t *ray_hits, uint16_t *ray_coarse,
                         uint64_t totals[3], int nthreads, int64_t wg_begin, int64_t wg_end) {
    const int R = s->rays_per_probe;
    const int sub = R < 32 ? R : 32;                                     /* subgroup width used */
    const int nsub = R / sub;                                            /* SUBGROUSPS_PER_WORKGROUP :403 */
    const int64_t nwg = (size > 0) ? nreq / size : 0;                    /* internal_render.cpp:1204 */
    if (wg_end < 0 || wg_end > nwg) wg_end = nwg;                        /* optional sub-range of the dispatch */
    if (wg_begin < 0) wg_begin = 0;
    size_t cells = (size_t)s->wx * (size_t)s->wy * (size_t)s->wz;
    if (field_new != field_old) memcpy(field_new, field_old, cells * sizeof(uint32_t));
    float *table = (float *)malloc(sizeof(float) * 3 * (size_t)R);
    lum_oracle_direction_table(R, table);
    uint64_t S = 0, H = 0;
    int used = 1;
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
    used = nthreads;
#pragma omp parallel for schedule(dynamic, 16) num_threads(nthreads) reduction(+ : S, H)
#endif
    for (int64_t w = wg_begin; w < wg_end; w++) {
        int64_t rq = w * (int64_t)size + shift;                          /* :368 */
        const int16_t *p = req + 4 * rq;                                 /* :369 */
        float lr[256], lg[256], lb[256];
        for (int i = 0; i < R; i++) {
            LumOracleRay ray;
            lum_oracle_trace(s, field_old, table, p, (uint32_t)(w * R + i), i, time_seed, &ray);
            lr[i] = ray.light[0]; lg[i] = ray.light[1]; lb[i] = ray.light[2];
            uint32_t ns = (uint32_t)(ray.coarse_samples + ray.refine_samples);
            S += ns; H += (uint64_t)ray.hit;
            if (ray_samples) ray_samples[w * R + i] = ns;
            if (ray_hits) ray_hits[w * R + i] = (uint8_t)ray.hit;
            if (ray_coarse) ray_coarse[w * R + i] = (uint16_t)ray.coarse_samples;
        }
        float fr = 0.0f, fg = 0.0f, fb = 0.0f;                           /* :401 */
        for (int g = 0; g < nsub; g++) {                                 /* :386,398,404-406 */
            fr += butterfly_sum(lr + g * sub, sub) / (float)sub;
            fg += butterfly_sum(lg + g * sub, sub) / (float)sub;
            fb += butterfly_sum(lb + g * sub, sub) / (float)sub;
        }
        fr /= (float)nsub; fg /= (float)nsub; fb /= (float)nsub;         /* :407 */
        if (pre_blend) { pre_blend[3 * w] = fr; pre_blend[3 * w + 1] = fg; pre_blend[3 * w + 2] = fb; }
        /* store_probe :349-354; probes outside the image are dropped (imageStore out of range) */
        if (p[0] >= 0 && p[0] < s->wx && p[1] >= 0 && p[1] < s->wy && p[2] >= 0 && p[2] < s->wz) {
            size_t idx = (size_t)p[0] + (size_t)s->wx * ((size_t)p[1] + (size_t)s->wy * (size_t)p[2]);
            v3 old = unpack10(field_old[idx]);
            old.x = g_clamp(old.x, 0.0f, 1.0f); old.y = g_clamp(old.y, 0.0f, 1.0f); old.z = g_clamp(old.z, 0.0f, 1.0f);
            float nl[3] = { g_mix(old.x, fr, s->reactiveness), g_mix(old.y, fg, s->reactiveness),
                            g_mix(old.z, fb, s->reactiveness) };
            field_new[idx] = lum_oracle_pack_rgb10a2(nl);
        }
    }
    free(table);
    if (totals) { totals[0] = S; totals[1] = H; totals[2] = (uint64_t)(wg_end > wg_begin ? wg_end - wg_begin : 0); }
    return used;
}

/* ---- shaders/glossy.frag as a CPU pass ------------------------------- */
/* CastRay_fast of glossy.frag:119-226 is statement for statement the one of radiance.comp:183-291
 * (checked by running both shader texts, tests/test_ref_glsl.py), so cast_ray_fast() above serves both.
 * One invocation per fragment of a width x height target; citations are glossy.frag lines. */
void lum_oracle_glossy(const LumOracleScene *s, const uint32_t *field, const LumOracleGlossyView *v,
                       int width, int height, const uint8_t *mat_norm, const float *depth,
                       float *frame, uint8_t *covered, uint32_t *samples, uint32_t *coarse, int nthreads) {
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 4) num_threads(nthreads)
#endif
    for (int py = 0; py < height; py++)
    for (int px = 0; px < width; px++) {
        size_t pi = (size_t)py * (size_t)width + (size_t)px;
        const uint8_t *mn = mat_norm + 4 * pi;
        const float *m = s->materials + 6 * (size_t)mn[0];               /* GetMat(load_mat(pix)) :443,:85-96 */
        float rough = m[5];
        int cov = !(rough > 0.5f);                                       /* fillStencilGlossy.frag:34-36 */
        if (covered) covered[pi] = (uint8_t)cov;
        if (samples) samples[pi] = 0;
        if (coarse) coarse[pi] = 0;
        if (!cov) continue;
        v3 dir = { v->camdir[0], v->camdir[1], v->camdir[2] };           /* :444 */
        float fcx = (float)px + 0.5f, fcy = (float)py + 0.5f;            /* gl_FragCoord.xy */
        float cx = fcx / v->frame_size[0] * 2.0f - 1.0f;                 /* :446 */
        float cy = fcy / v->frame_size[1] * 2.0f - 1.0f;
        float dep = depth[pi] * 1000.0f;                                 /* load_depth :377-382 */
        v3 origin = {                                                    /* get_origin_from_depth :41-47 */
            ((v->campos[0] + v->horizline_scaled[0] * cx) + v->vertiline_scaled[0] * cy) + v->camdir[0] * dep,
            ((v->campos[1] + v->horizline_scaled[1] * cx) + v->vertiline_scaled[1] * cy) + v->camdir[1] * dep,
            ((v->campos[2] + v->horizline_scaled[2] * cx) + v->vertiline_scaled[2] * cy) + v->camdir[2] * dep };
        v3 normal = { (float)mn[1] / 255.0f * 2.0f - 1.0f,               /* load_norm :367-371 */
                      (float)mn[2] / 255.0f * 2.0f - 1.0f,
                      (float)mn[3] / 255.0f * 2.0f - 1.0f };
        v3 light = { 0.0f, 0.0f, 0.0f }, refl = { 1.0f, 1.0f, 1.0f };    /* :450-451 */
        /* ProcessHit(origin, direction, 0, normal, mat, ...) :454-456, :312-326 */
        origin.x = origin.x + 0.0f * dir.x; origin.y = origin.y + 0.0f * dir.y; origin.z = origin.z + 0.0f * dir.z; /* :316 */
        v3 bp = { (origin.x + normal.x * 16.0f) / 16.0f, (origin.y + normal.y * 16.0f) / 16.0f,
                  (origin.z + normal.z * 16.0f) / 16.0f };               /* sample_radiance :98-103 */
        v3 g = gather(s, field, bp);
        refl.x *= m[0]; refl.y *= m[1]; refl.z *= m[2];                  /* :322 */
        light.x += refl.x * (m[4] + g.x);                                /* :323 */
        light.y += refl.y * (m[4] + g.y);
        light.z += refl.z * (m[4] + g.z);
        {   /* reflect(I,N) = I - 2.0*dot(N,I)*N, GLSL 4.50 s8.5, left to right (decree 14) :325 */
            float k = 2.0f * g_dot(normal, dir);
            dir.x = dir.x - k * normal.x; dir.y = dir.y - k * normal.y; dir.z = dir.z - k * normal.z;
        }
        /* trace_glossy_ray :328-360 */
        float fraction = 0.0f;
        v3 n2 = { 0.0f, 0.0f, 0.0f };
        mat_t mat = { { 0.0f, 0.0f, 0.0f }, 0.0f };
        int voxel = 0, ncoarse = 0, nrefine = 0;
        int hit = cast_ray_fast(s, origin, dir, &fraction, &n2, &mat, &voxel, &ncoarse, &nrefine); /* :341 */
        if (samples) samples[pi] = (uint32_t)(ncoarse + nrefine);
        if (coarse) coarse[pi] = (uint32_t)ncoarse;
        if (hit) {                                                       /* :342-345 */
            v3 x = { origin.x + fraction * dir.x, origin.y + fraction * dir.y, origin.z + fraction * dir.z };
            v3 bp2 = { (x.x + n2.x * 16.0f) / 16.0f, (x.y + n2.y * 16.0f) / 16.0f, (x.z + n2.z * 16.0f) / 16.0f };
            v3 g2 = gather(s, field, bp2);
            refl.x *= mat.color[0]; refl.y *= mat.color[1]; refl.z *= mat.color[2];
            light.x += refl.x * (mat.emmit + g2.x);
            light.y += refl.y * (mat.emmit + g2.y);
            light.z += refl.z * (mat.emmit + g2.z);
        } else {                                                         /* :346-357 */
            v3 gld = { v->global_light_dir[0], v->global_light_dir[1], v->global_light_dir[2] };
            float glp = -g_dot(dir, gld);                                /* :349 */
            if (glp > 0.9f) {                                            /* :350-353 */
                float w = glp - 0.9f;
                light.x += (0.9f * 0.5f) * refl.x * w * 10.0f;
                light.y += (0.9f * 0.5f) * refl.y * w * 10.0f;
                light.z += (0.6f * 0.5f) * refl.z * w * 10.0f;
            }
            light.x += (0.53f * 0.1f) * refl.x;                          /* :356 */
            light.y += (0.81f * 0.1f) * refl.y;
            light.z += (0.92f * 0.1f) * refl.z;
        }
        frame[4 * pi + 0] = light.x / 8.0f;                              /* :490, encode_color :435-437 */
        frame[4 * pi + 1] = light.y / 8.0f;
        frame[4 * pi + 2] = light.z / 8.0f;
        frame[4 * pi + 3] = 1.0f - rough;
    }
}

/* A plausible G-buffer for the glossy tests and bench: what the raster passes would leave for an
 * orthographic camera looking along camdir (start_raygen, internal_render.cpp:1580-1592).  NOT a restatement
 * of anything: it only fabricates inputs (first solid voxel along the view ray by fine sampling). */
void lum_oracle_gbuffer(const LumOracleScene *s, const LumOracleGlossyView *v, int width, int height,
                        uint8_t *mat_norm, float *depth) {
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 4)
#endif
    for (int py = 0; py < height; py++)
    for (int px = 0; px < width; px++) {
        size_t pi = (size_t)py * (size_t)width + (size_t)px;
        float cx = ((float)px + 0.5f) / v->frame_size[0] * 2.0f - 1.0f;
        float cy = ((float)py + 0.5f) / v->frame_size[1] * 2.0f - 1.0f;
        double o[3], d[3] = { v->camdir[0], v->camdir[1], v->camdir[2] };
        for (int k = 0; k < 3; k++) o[k] = (double)v->campos[k] + (double)v->horizline_scaled[k] * cx + (double)v->vertiline_scaled[k] * cy;
        int found = 0, vox = 0;
        i3 prev = { 0x7fffffff, 0, 0 }, cur = { 0, 0, 0 };
        double t = 0.0;
        for (; t < 1000.0; t += 0.125) {
            v3 p = { (float)(o[0] + d[0] * t), (float)(o[1] + d[1] * t), (float)(o[2] + d[2] * t) };
            if (p.x < 0.0f || p.y < 0.0f || p.z < 0.0f || p.x >= 16.0f * s->wx || p.y >= 16.0f * s->wy || p.z >= 16.0f * s->wz) continue;
            cur = trunc3(p);
            int bid;
            get_voxel(s, cur, &vox, &bid);
            if (vox != 0) { found = 1; break; }
            prev = cur;
            if (bid == 0) {     /* empty block: hop over the lattice points that lie strictly inside it */
                double te = 1e30;
                for (int k = 0; k < 3; k++) {
                    double pk = o[k] + d[k] * t, c0 = 16.0 * floor(pk / 16.0);
                    if (d[k] > 0.0) te = fmin(te, (c0 + 16.0 - pk) / d[k]);
                    else if (d[k] < 0.0) te = fmin(te, (c0 - pk) / d[k]);
                }
                double hop = floor((te - 0.5) / 0.125) * 0.125;       /* stay 0.5 short of the face */
                if (hop > 0.0) t += hop;
            }
        }
        uint8_t *mn = mat_norm + 4 * pi;
        if (!found) { mn[0] = 0; mn[1] = mn[2] = mn[3] = 128; depth[pi] = 1.0f; continue; }
        int n[3] = { 0, 0, 0 };
        if (prev.x != 0x7fffffff) {
            if (prev.x != cur.x) n[0] = prev.x - cur.x; else if (prev.y != cur.y) n[1] = prev.y - cur.y; else if (prev.z != cur.z) n[2] = prev.z - cur.z;
        }
        if (!n[0] && !n[1] && !n[2]) n[2] = 1;
        for (int k = 0; k < 3; k++) { int c = n[k] > 0 ? 255 : (n[k] < 0 ? 0 : 128); mn[1 + k] = (uint8_t)c; }
        mn[0] = (uint8_t)vox;
        depth[pi] = (float)(t / 1000.0);
    }
}

/* ---- update_radiance request list (internal_render.cpp:1142-1169) ------------------ */
static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
int64_t lum_oracle_build_requests(int wx, int wy, int wz, const int16_t *world,
                                  const int16_t *special, int nspecial,
                                  int16_t *out, int64_t cap) {
    size_t cells = (size_t)wx * (size_t)wy * (size_t)wz;
    uint8_t *set = (uint8_t *)calloc(cells, 1);
    int64_t n = 0;
    for (int zz = 0; zz < wz; zz++)
    for (int yy = 0; yy < wy; yy++)
    for (int xx = 0; xx < wx; xx++) {
        int sum = 0;
        for (int dz = -1; dz <= 1; dz++)
        for (int dy = -1; dy <= 1; dy++)
        for (int dx = -1; dx <= 1; dx++) {
            int x = clampi(xx + dx, 0, wx - 1), y = clampi(yy + dy, 0, wy - 1), z = clampi(zz + dz, 0, wz - 1);
            sum += world[(size_t)x + (size_t)wx * ((size_t)y + (size_t)wy * (size_t)z)];
        }
        if (sum > 0) {
            if (n < cap) { out[4 * n] = (int16_t)xx; out[4 * n + 1] = (int16_t)yy; out[4 * n + 2] = (int16_t)zz; out[4 * n + 3] = 0; }
            n++;
            set[(size_t)xx + (size_t)wx * ((size_t)yy + (size_t)wy * (size_t)zz)] = 1;
        }
    }
    for (int i = 0; i < nspecial; i++) {                                 /* :1164-1168 */
        const int16_t *u = special + 4 * i;
        if (!set[(size_t)u[0] + (size_t)wx * ((size_t)u[1] + (size_t)wy * (size_t)u[2])]) {
            if (n < cap) { out[4 * n] = u[0]; out[4 * n + 1] = u[1]; out[4 * n + 2] = u[2]; out[4 * n + 3] = u[3]; }
            n++;
        }
    }
    free(set);
    return n;
}

/* ---- shift_radiance (internal_render.cpp:1218-1335) -------------------------------- */
void lum_oracle_shift_field(int wx, int wy, int wz, const uint32_t *in, uint32_t *out,
                            int dx, int dy, int dz) {
    size_t cells = (size_t)wx * (size_t)wy * (size_t)wz;
    memcpy(out, in, cells * sizeof(uint32_t));
    if (abs(dx) >= wx || abs(dy) >= wy || abs(dz) >= wz) return;
    for (int z = 0; z < wz; z++)
    for (int y = 0; y < wy; y++)
    for (int x = 0; x < wx; x++) {
        int sx = x + dx, sy = y + dy, sz = z + dz;
        if (sx < 0 || sx >= wx || sy < 0 || sy >= wy || sz < 0 || sz >= wz) continue;
        out[(size_t)x + (size_t)wx * ((size_t)y + (size_t)wy * (size_t)z)] =
            in[(size_t)sx + (size_t)wx * ((size_t)sy + (size_t)wy * (size_t)sz)];
    }
}

int lum_oracle_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
