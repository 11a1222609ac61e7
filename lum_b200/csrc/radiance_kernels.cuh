// radiance_kernels.cuh -- kernels of the radiance stage (sm_100a).
#pragma once
#include "radiance_device.cuh"

namespace lumcu {

struct UpdateArgs {
    const short4 *__restrict__ requests;   // x,y,z,w per request
    const float *__restrict__ dir_table;   // [R][3]
    uint32_t *__restrict__ staged;         // [nwg] packed texel per workgroup (committed afterwards)
    float *__restrict__ dbg_pre_blend;     // [nwg][3] or null
    uint32_t *__restrict__ dbg_samples;    // [nwg*R] or null
    uint8_t *__restrict__ dbg_hits;        // [nwg*R] or null
    unsigned long long *__restrict__ dbg_totals;  // S, H
    long long wg_begin;                    // first workgroup of this launch
    long long wg_end;
    int time_seed, size, shift;
    int z_begin, z_end;                    // only probes with z in [z_begin,z_end) are processed
    float reactiveness;
    float sun_x, sun_y, sun_z;
};

// One CTA = one workgroup of radiance.comp (radiance.comp:360-414): R threads, one ray each.
// The subgroup reduction is the xor butterfly the oracle pins; _light[] lives in shared memory.
template <int R, bool DEBUG>
__global__ void __launch_bounds__(R) radiance_update_wg_kernel(SceneView s, UpdateArgs a) {
    constexpr int SUB = R < 32 ? R : 32;
    constexpr int NSUB = R / SUB;
    __shared__ float sh_light[NSUB][3];

    const long long w = a.wg_begin + (long long)blockIdx.x;
    const int ray = threadIdx.x;
    const short4 rq = a.requests[w * (long long)a.size + a.shift];
    if (rq.z < a.z_begin || rq.z >= a.z_end) return;          // uniform per CTA
    const uint32_t gid_x = (uint32_t)(w * R + ray);
    const f3 fib = {__ldg(a.dir_table + 3 * ray), __ldg(a.dir_table + 3 * ray + 1), __ldg(a.dir_table + 3 * ray + 2)};
    const f3 sun = {a.sun_x, a.sun_y, a.sun_z};

    RayResult r = trace_probe_ray(s, rq.x, rq.y, rq.z, gid_x, ray, fib, a.time_seed, sun);

    float lx = r.light.x, ly = r.light.y, lz = r.light.z;
    constexpr unsigned mask = SUB == 32 ? 0xffffffffu : ((1u << SUB) - 1u);
#pragma unroll
    for (int k = SUB / 2; k >= 1; k >>= 1) {
        lx += __shfl_xor_sync(mask, lx, k);
        ly += __shfl_xor_sync(mask, ly, k);
        lz += __shfl_xor_sync(mask, lz, k);
    }
    if constexpr (DEBUG) {
        if (a.dbg_samples) a.dbg_samples[w * R + ray] = (uint32_t)r.samples;
        if (a.dbg_hits) a.dbg_hits[w * R + ray] = (uint8_t)r.hit;
        unsigned long long sm = (unsigned long long)r.samples, hm = (unsigned long long)r.hit;
#pragma unroll
        for (int k = SUB / 2; k >= 1; k >>= 1) {
            sm += __shfl_xor_sync(mask, sm, k);
            hm += __shfl_xor_sync(mask, hm, k);
        }
        if ((ray % SUB) == 0) { atomicAdd(a.dbg_totals, sm); atomicAdd(a.dbg_totals + 1, hm); }
    }
    if ((ray % SUB) == 0) {
        sh_light[ray / SUB][0] = lx / (float)SUB;             // subgroupAdd(..)/clamp(R,0,subgroupSize) :386
        sh_light[ray / SUB][1] = ly / (float)SUB;
        sh_light[ray / SUB][2] = lz / (float)SUB;
    }
    __syncthreads();
    if (ray == 0) {
        float fr = 0.0f, fg = 0.0f, fb = 0.0f;                // :401-407
#pragma unroll
        for (int g = 0; g < NSUB; g++) { fr += sh_light[g][0]; fg += sh_light[g][1]; fb += sh_light[g][2]; }
        fr /= (float)NSUB; fg /= (float)NSUB; fb /= (float)NSUB;
        if constexpr (DEBUG) {
            if (a.dbg_pre_blend) {
                a.dbg_pre_blend[3 * w] = fr; a.dbg_pre_blend[3 * w + 1] = fg; a.dbg_pre_blend[3 * w + 2] = fb;
            }
        }
        // store_probe :349-354 (the store itself is deferred to the commit kernel: Jacobi)
        uint32_t word = 0;
        if ((unsigned)rq.x < (unsigned)s.wx && (unsigned)rq.y < (unsigned)s.wy && (unsigned)rq.z < (unsigned)s.wz) {
            f3 old = unpack10(__ldg(s.field + (rq.x + s.wx * (rq.y + s.wy * rq.z))));
            old.x = gmin(gmax(old.x, 0.0f), 1.0f); old.y = gmin(gmax(old.y, 0.0f), 1.0f); old.z = gmin(gmax(old.z, 0.0f), 1.0f);
            const float t = a.reactiveness, omt = 1.0f - t;
            word = pack10(old.x * omt + fr * t, old.y * omt + fg * t, old.z * omt + fb * t);
        }
        a.staged[w] = word;
    }
}

// Deferred imageStore (radiance.comp:353): scatter the staged texels of workgroups
// [wg_begin, wg_end) into the field once every read of the old field is done.
__global__ void commit_kernel(uint32_t *__restrict__ field, const uint32_t *__restrict__ staged,
                              const short4 *__restrict__ requests, long long wg_begin, long long wg_end,
                              int size, int shift, int wx, int wy, int wz, int z_begin, int z_end) {
    long long w = wg_begin + (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= wg_end) return;
    short4 rq = requests[w * (long long)size + shift];
    if (rq.z < z_begin || rq.z >= z_end) return;
    if ((unsigned)rq.x < (unsigned)wx && (unsigned)rq.y < (unsigned)wy && (unsigned)rq.z < (unsigned)wz)
        field[rq.x + wx * (rq.y + wy * rq.z)] = staged[w];
}

// shift_radiance (internal_render.cpp:1218-1335): field[p] = tmp[p + d] on the overlap box.
__global__ void shift_kernel(uint32_t *__restrict__ field, const uint32_t *__restrict__ tmp,
                             int wx, int wy, int wz, int dx, int dy, int dz) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long cells = (long long)wx * wy * wz;
    if (i >= cells) return;
    int x = (int)(i % wx), y = (int)((i / wx) % wy), z = (int)(i / ((long long)wx * wy));
    int sx = x + dx, sy = y + dy, sz = z + dz;
    if ((unsigned)sx < (unsigned)wx && (unsigned)sy < (unsigned)wy && (unsigned)sz < (unsigned)wz)
        field[i] = tmp[sx + wx * (sy + (long long)wy * sz)];
}

__global__ void unpack_table_kernel(float *__restrict__ out) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < 1024) out[c] = unorm10_to_float((uint32_t)c);
}

// ---- request-list builder (internal_render.cpp:1142-1169) ------------------------------
constexpr int BUILD_THREADS = 256;

__device__ __forceinline__ bool probe_selected(const int16_t *__restrict__ world, int wx, int wy, int wz,
                                               int x, int y, int z) {
    int sum = 0;
#pragma unroll
    for (int dz = -1; dz <= 1; dz++)
#pragma unroll
        for (int dy = -1; dy <= 1; dy++)
#pragma unroll
            for (int dx = -1; dx <= 1; dx++) {
                int cx = min(max(x + dx, 0), wx - 1), cy = min(max(y + dy, 0), wy - 1), cz = min(max(z + dz, 0), wz - 1);
                sum += __ldg(world + (cx + wx * (cy + (long long)wy * cz)));
            }
    return sum > 0;
}

// pass 1: per-block counts of selected cells (+ the flag byte per cell)
__global__ void build_count_kernel(const int16_t *__restrict__ world, int wx, int wy, int wz,
                                   uint8_t *__restrict__ flags, uint32_t *__restrict__ block_counts) {
    long long cells = (long long)wx * wy * wz;
    long long i = (long long)blockIdx.x * BUILD_THREADS + threadIdx.x;
    bool sel = false;
    if (i < cells) {
        int x = (int)(i % wx), y = (int)((i / wx) % wy), z = (int)(i / ((long long)wx * wy));
        sel = probe_selected(world, wx, wy, wz, x, y, z);
        flags[i] = sel ? 1 : 0;
    }
    int c = __syncthreads_count(sel);
    if (threadIdx.x == 0) block_counts[blockIdx.x] = (uint32_t)c;
}

// pass 2: exclusive scan of the block counts by ONE block (nblocks <= 2^24 handled by looping)
__global__ void build_scan_kernel(const uint32_t *__restrict__ block_counts, unsigned long long *__restrict__ block_offsets,
                                  long long nblocks, unsigned long long *__restrict__ total) {
    __shared__ unsigned long long warp_sums[32];
    __shared__ unsigned long long carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (long long base = 0; base < nblocks; base += blockDim.x) {
        long long i = base + threadIdx.x;
        unsigned long long v = i < nblocks ? block_counts[i] : 0ull, incl = v;
#pragma unroll
        for (int k = 1; k < 32; k <<= 1) {
            unsigned long long n = __shfl_up_sync(0xffffffffu, incl, k);
            if (lane >= k) incl += n;
        }
        if (lane == 31) warp_sums[wid] = incl;
        __syncthreads();
        if (wid == 0) {
            unsigned long long ws = lane < nw ? warp_sums[lane] : 0ull, wincl = ws;
#pragma unroll
            for (int k = 1; k < 32; k <<= 1) {
                unsigned long long n = __shfl_up_sync(0xffffffffu, wincl, k);
                if (lane >= k) wincl += n;
            }
            warp_sums[lane] = wincl - ws;   // exclusive
        }
        __syncthreads();
        unsigned long long excl = carry + warp_sums[wid] + incl - v;
        if (i < nblocks) block_offsets[i] = excl;
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry;
}

// pass 3: ordered scatter (z,y,x order == linear cell order)
__global__ void build_scatter_kernel(const uint8_t *__restrict__ flags, const unsigned long long *__restrict__ block_offsets,
                                     int wx, int wy, int wz, short4 *__restrict__ out, long long cap) {
    __shared__ int warp_counts[BUILD_THREADS / 32];
    long long cells = (long long)wx * wy * wz;
    long long i = (long long)blockIdx.x * BUILD_THREADS + threadIdx.x;
    bool sel = i < cells && flags[i];
    unsigned b = __ballot_sync(0xffffffffu, sel);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) warp_counts[wid] = __popc(b);
    __syncthreads();
    int before = 0;
    for (int k = 0; k < wid; k++) before += warp_counts[k];
    if (sel) {
        long long pos = (long long)block_offsets[blockIdx.x] + before + __popc(b & ((1u << lane) - 1u));
        if (pos < cap) {
            int x = (int)(i % wx), y = (int)((i / wx) % wy), z = (int)(i / ((long long)wx * wy));
            out[pos] = make_short4((short)x, (short)y, (short)z, 0);
        }
    }
}

// specials not already selected, in submission order (internal_render.cpp:1164-1168); ONE block.
__global__ void build_specials_kernel(const uint8_t *__restrict__ flags, int wx, int wy, int wz,
                                      const short4 *__restrict__ special, int nspecial,
                                      short4 *__restrict__ out, unsigned long long *__restrict__ total, long long cap) {
    __shared__ int warp_counts[32];
    __shared__ unsigned long long base;
    if (threadIdx.x == 0) base = *total;
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int start = 0; start < nspecial; start += blockDim.x) {
        int i = start + threadIdx.x;
        bool keep = false;
        short4 u = make_short4(0, 0, 0, 0);
        if (i < nspecial) {
            u = special[i];
            bool inside = (unsigned)u.x < (unsigned)wx && (unsigned)u.y < (unsigned)wy && (unsigned)u.z < (unsigned)wz;
            keep = inside && !flags[u.x + wx * (u.y + (long long)wy * u.z)];
        }
        unsigned b = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) warp_counts[wid] = __popc(b);
        __syncthreads();
        int before = 0, all = 0;
        for (int k = 0; k < nw; k++) { if (k < wid) before += warp_counts[k]; all += warp_counts[k]; }
        if (keep) {
            long long pos = (long long)base + before + __popc(b & ((1u << lane) - 1u));
            if (pos < cap) out[pos] = u;
        }
        __syncthreads();
        if (threadIdx.x == 0) base += all;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = base;
}

// first workgroup index whose probe z >= zb, over the z-sorted prefix [0, nwg_sorted)
__global__ void slab_bound_kernel(const short4 *__restrict__ requests, long long nwg_sorted, int size, int shift,
                                  int zb0, int zb1, long long *__restrict__ out2) {
    long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= nwg_sorted) return;
    int z = requests[w * (long long)size + shift].z;
    int zp = w == 0 ? -32768 : requests[(w - 1) * (long long)size + shift].z;
    if (zp < zb0 && z >= zb0) out2[0] = w;
    if (zp < zb1 && z >= zb1) out2[1] = w;
}

}  // namespace lumcu
