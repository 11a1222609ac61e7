// device_on_host.cpp -- TEST INFRASTRUCTURE ONLY (see cuda_host_shim.hpp).
// Runs the radiance-probe update with the per-thread functions of lum_b200/csrc/radiance_device.cuh compiled for
// the host, one "thread" after another, and the workgroup reduction / blend of the kernels restated around them
// (radiance_kernels.cuh phase D: xor butterfly per <=32-ray subgroup, /SUB, mean of subgroups, mix, pack).
//   mode 0: general march (march_step)            -- what the kernels run for an unvalidated scene
//   mode 1: trace_probe_ray<true>                 -- the one-workgroup-per-CTA kernel (fused pair of fast steps)
//   mode 2: fast march, fused two steps           -- march_two_steps_fast, the pooled kernel's loop body
//   mode 3: fast march, one step per iteration    -- march_step_fast, what the fused pair is defined against
// Compiled with g++ -ffp-contract=off (the kernels: nvcc -fmad=false), so one rounding per written operation.
#include "cuda_host_shim.hpp"
#include "../../lum_b200/csrc/radiance_device.cuh"

#include <vector>

using namespace lumcu;

static RayResult trace_fused(const SceneView &s, int px, int py, int pz, uint32_t gid_x, int ray, f3 fib, int time_seed, f3 sun) {
    f3 o, d, inv;
    setup_ray(px, py, pz, gid_x, ray, fib, time_seed, o, d, inv);
    const MarchConsts mc = {(float)s.wx * 16.0f, (float)s.wy * 16.0f, (float)s.wz * 16.0f};
    const uint3 wsbits = {__float_as_uint(mc.wsx), __float_as_uint(mc.wsy), __float_as_uint(mc.wsz)};
    const i3 cm = {0x4B000000 + (d.x > 0.0f ? 16 : 0), 0x4B000000 + (d.y > 0.0f ? 16 : 0), 0x4B000000 + (d.z > 0.0f ? 16 : 0)};
    float t = 0.0f, t_pre = 0.0f;
    int samples = 0, steps = 0;
    bool hit = false, done;
    do { done = march_two_steps_fast(s, wsbits, o, d, inv, cm, t, t_pre, hit, steps); samples += steps; } while (!done);
    // what the pooled kernel stores and pool_shade() reads back: hit ? t : -t_pre
    if (!hit && resolve_pending(s, mc, o, d, t_pre)) { hit = true; t = t_pre; }
    RayResult res;
    int nr;
    res.light = finish_ray<true>(s, o, d, inv, hit ? t : t_pre, hit, sun, nr);
    res.hit = hit;
    res.samples = samples + nr;
    return res;
}

// mode 3: the fast march one step at a time (march_step_fast, what the fused pair is defined against)
static RayResult trace_single(const SceneView &s, int px, int py, int pz, uint32_t gid_x, int ray, f3 fib, int time_seed, f3 sun) {
    f3 o, d, inv;
    setup_ray(px, py, pz, gid_x, ray, fib, time_seed, o, d, inv);
    const MarchConsts mc = {(float)s.wx * 16.0f, (float)s.wy * 16.0f, (float)s.wz * 16.0f};
    const uint3 wsbits = {__float_as_uint(mc.wsx), __float_as_uint(mc.wsy), __float_as_uint(mc.wsz)};
    const i3 cm = {0x4B000000 + (d.x > 0.0f ? 16 : 0), 0x4B000000 + (d.y > 0.0f ? 16 : 0), 0x4B000000 + (d.z > 0.0f ? 16 : 0)};
    float t = 0.0f, t_pre = 0.0f;
    int samples = 0;
    bool hit = false, done;
    do { done = march_step_fast(s, wsbits, o, d, inv, cm, t, t_pre, hit); samples++; } while (!done);
    if (!hit && resolve_pending(s, mc, o, d, t_pre)) { hit = true; t = t_pre; }
    RayResult res;
    int nr;
    res.light = finish_ray<true>(s, o, d, inv, hit ? t : t_pre, hit, sun, nr);
    res.hit = hit;
    res.samples = samples + nr;
    return res;
}

extern "C" int device_on_host_update(int wx, int wy, int wz, int capacity, int brick0_empty, int R, float reactiveness,
                                     const int16_t *world, const uint8_t *bricks, const float *materials,
                                     const uint32_t *field_old, uint32_t *field_new,
                                     const int16_t *requests, long long nreq, int time_seed, int size, int shift,
                                     const float *table, const float *sun3, int mode,
                                     float *pre_blend, uint32_t *ray_samples, uint8_t *ray_hits) {
    if (R != 16 && R != 32 && R != 64 && R != 128 && R != 256) return -1;
    if (size < 1 || shift < 0 || shift >= size) return -1;
    std::vector<float4> mat4(256);
    for (int i = 0; i < 256; i++) mat4[i] = make_float4(materials[6 * i], materials[6 * i + 1], materials[6 * i + 2], materials[6 * i + 4]);
    uint32_t flags = 0;
    const SceneView s{world, bricks, materials, mat4.data(), field_old, wx, wy, wz, capacity, brick0_empty, &flags};
    const f3 sun = {sun3[0], sun3[1], sun3[2]};
    const long long cells = (long long)wx * wy * wz;
    for (long long i = 0; i < cells; i++) field_new[i] = field_old[i];
    const int SUB = R < 32 ? R : 32, NSUB = R / SUB;
    const long long nwg = nreq / size;
#pragma omp parallel for schedule(dynamic, 64)      // workgroups are independent (Jacobi update); built with -fopenmp
    for (long long w = 0; w < nwg; w++) {
        std::vector<f3> light(R);
        const int16_t *rq = requests + 4 * (w * size + shift);
        for (int ray = 0; ray < R; ray++) {
            const uint32_t gid_x = (uint32_t)(w * R + ray);
            const f3 fib = {table[3 * ray], table[3 * ray + 1], table[3 * ray + 2]};
            RayResult r;
            if (mode == 0) r = trace_probe_ray<false>(s, rq[0], rq[1], rq[2], gid_x, ray, fib, time_seed, sun);
            else if (mode == 1) r = trace_probe_ray<true>(s, rq[0], rq[1], rq[2], gid_x, ray, fib, time_seed, sun);
            else if (mode == 2) r = trace_fused(s, rq[0], rq[1], rq[2], gid_x, ray, fib, time_seed, sun);
            else r = trace_single(s, rq[0], rq[1], rq[2], gid_x, ray, fib, time_seed, sun);
            light[ray] = r.light;
            if (ray_samples) ray_samples[w * R + ray] = (uint32_t)r.samples;
            if (ray_hits) ray_hits[w * R + ray] = r.hit ? 1 : 0;
        }
        float fr = 0.0f, fg = 0.0f, fb = 0.0f;
        for (int g = 0; g < NSUB; g++) {
            float a[3][32];
            for (int l = 0; l < SUB; l++) { a[0][l] = light[g * SUB + l].x; a[1][l] = light[g * SUB + l].y; a[2][l] = light[g * SUB + l].z; }
            for (int k = SUB / 2; k >= 1; k >>= 1)
                for (int c = 0; c < 3; c++) {
                    float b[32];
                    for (int l = 0; l < SUB; l++) b[l] = a[c][l] + a[c][l ^ k];     // __shfl_xor_sync
                    for (int l = 0; l < SUB; l++) a[c][l] = b[l];
                }
            fr += a[0][0] / (float)SUB; fg += a[1][0] / (float)SUB; fb += a[2][0] / (float)SUB;
        }
        fr /= (float)NSUB; fg /= (float)NSUB; fb /= (float)NSUB;
        if (pre_blend) { pre_blend[3 * w] = fr; pre_blend[3 * w + 1] = fg; pre_blend[3 * w + 2] = fb; }
        if ((unsigned)rq[0] < (unsigned)wx && (unsigned)rq[1] < (unsigned)wy && (unsigned)rq[2] < (unsigned)wz) {
            const long long idx = rq[0] + (long long)wx * (rq[1] + (long long)wy * rq[2]);
            f3 old = unpack10(field_old[idx]);
            old.x = gmin(gmax(old.x, 0.0f), 1.0f); old.y = gmin(gmax(old.y, 0.0f), 1.0f); old.z = gmin(gmax(old.z, 0.0f), 1.0f);
            const float t = reactiveness, omt = 1.0f - t;
            field_new[idx] = pack10(old.x * omt + fr * t, old.y * omt + fg * t, old.z * omt + fb * t);
        }
    }
    return 0;
}

// glossy.frag for every fragment of a frame: glossy_fragment() is the body of glossy_kernel itself.
// view = LumcuGlossyView as 17 floats: campos, camdir, horizline_scaled, vertiline_scaled, global_light_dir, frame_size.
extern "C" int device_on_host_glossy(int wx, int wy, int wz, int capacity, int brick0_empty,
                                     const int16_t *world, const uint8_t *bricks, const float *materials,
                                     const uint32_t *field, const float *view, int width, int height,
                                     const uint8_t *mat_norm, const float *depth, float *frame, uint8_t *covered) {
    std::vector<float4> mat4(256);
    for (int i = 0; i < 256; i++) mat4[i] = make_float4(materials[6 * i], materials[6 * i + 1], materials[6 * i + 2], materials[6 * i + 4]);
    uint32_t flags = 0;
    for (long long i = 0; i < (long long)wx * wy * wz; i++) if ((unsigned)world[i] >= (unsigned)capacity) flags = 1;   // validate_world_kernel
    const SceneView s{world, bricks, materials, mat4.data(), field, wx, wy, wz, capacity, brick0_empty, &flags};
    GlossyArgs a{};
    for (int k = 0; k < 3; k++) {
        a.campos[k] = view[k]; a.camdir[k] = view[3 + k]; a.hs[k] = view[6 + k]; a.vs[k] = view[9 + k]; a.gld[k] = view[12 + k];
    }
    a.fw = view[15]; a.fh = view[16];
    a.width = width; a.height = height;
    a.mat_norm = reinterpret_cast<const uchar4 *>(mat_norm);
    a.depth = depth;
    a.frame = reinterpret_cast<float4 *>(frame);
    a.covered = covered;
    for (int py = 0; py < height; py++)
        for (int px = 0; px < width; px++) glossy_fragment(s, a, px, py);
    return 0;
}

// map.comp for one model: the host half of lumcu_map_model (invocations per axis = 4*ceil(area/4), z fastest) around
// map_model_invocation(), the body of map_model_kernel.  bricks is modified in place.
extern "C" int device_on_host_map_model(int wx, int wy, int wz, int capacity, const int16_t *world, uint8_t *bricks,
                                        const uint8_t *model, int sx, int sy, int sz, const float *inv16,
                                        const int32_t *shift, const int32_t *area, int static_count) {
    MapArgs a;
    for (int i = 0; i < 16; i++) a.m[i] = inv16[i];
    for (int k = 0; k < 3; k++) {
        a.shift[k] = shift[k];
        const int groups = (area[k] + 3) / 4;
        a.ext[k] = groups > 0 ? groups * 4 : 0;
    }
    a.msx = sx; a.msy = sy; a.msz = sz;
    a.static_count = static_count;
    for (int gx = 0; gx < a.ext[0]; gx++)
        for (int gy = 0; gy < a.ext[1]; gy++)
            for (int gz = 0; gz < a.ext[2]; gz++)
                map_model_invocation(world, bricks, model, wx, wy, wz, capacity, a, gx, gy, gz);
    return 0;
}
