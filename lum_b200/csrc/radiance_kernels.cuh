// radiance_kernels.cuh -- kernels of the radiance stage (sm_100a).
#pragma once
#include "radiance_device.cuh"

namespace lumcu {

// Multi-GPU exchange (one process per GPU, NVLink peer memory): the new texel of workgroup w is stored into slot w of
// the "inbox" of EVERY rank -- the rank's own staged[] array and the peers' through IPC-mapped pointers -- straight
// from the trace kernel's last phase, so the transfer rides under the tracing instead of following it.  n <= 1: only
// UpdateArgs::staged is written.
constexpr int MAX_PEERS = 16;
struct PeerWords { uint32_t *p[MAX_PEERS]; int n; };

struct UpdateArgs {
    const short4 *__restrict__ requests;   // x,y,z,w per request
    const float *__restrict__ dir_table;   // [R][3]
    uint32_t *__restrict__ staged;         // [nwg] packed texel per workgroup (committed afterwards)
    PeerWords inbox;                       // n > 1: staged[] of every rank (own included), see above
    const unsigned long long *__restrict__ nreq_dev;   // non-null: the request count lives on the device (a list
                                           // built without a host sync); the dispatch is min(wg_end, *nreq_dev / size)
    float *__restrict__ dbg_pre_blend;     // [nwg][3] or null
    uint32_t *__restrict__ dbg_samples;    // [nwg*R] or null
    uint8_t *__restrict__ dbg_hits;        // [nwg*R] or null
    unsigned long long *__restrict__ dbg_totals;  // S, H
    long long wg_begin;                    // first workgroup of this launch
    long long wg_end;
    // Probe sharding (multi-GPU), il_n == 1: the whole range.  il_chunk == 0: CONTIGUOUS -- shard r of n takes
    // workgroups [T*r/n, T*(r+1)/n) of the range's T (equal probe counts whatever the scene looks like, and each rank
    // touches only its own part of the world and field: at 8 ranks 4 % faster than dealing chunks out, whose ranks
    // each pull the whole scene through a cold L2).  il_chunk > 0: INTERLEAVED -- the range is cut into chunks of
    // il_chunk workgroups and this launch takes chunks il_rank, il_rank + il_n, ... (every rank samples every part of
    // the world: balances ray LENGTHS too, for scenes whose cost per probe varies by region).
    int il_chunk, il_rank, il_n;
    int time_seed, size, shift;
    int z_begin, z_end;                    // only probes with z in [z_begin,z_end) are processed
    float reactiveness;
    float sun_x, sun_y, sun_z;
};

// i-th workgroup of this launch -> workgroup index of the dispatch (see UpdateArgs::il_*).  first = where this launch's
// window starts (shard_window).  il_chunk == 0: the window is the shard, contiguous.
__device__ __forceinline__ long long il_map(long long first, int il_chunk, int il_rank, int il_n, long long i) {
    if (il_n <= 1 || il_chunk == 0) return first + i;
    const unsigned c = (unsigned)i / (unsigned)il_chunk;           // launch indices fit 32 bits (a 64-bit division per CTA is ~2 % of a trace)
    return first + ((long long)c * il_n + il_rank) * il_chunk + (long long)((unsigned)i - c * (unsigned)il_chunk);
}

// last workgroup (exclusive) of this launch: the host's bound, cut down to the device-side request count if there is one
// (A 64-bit division here cost the C4 update 0.23 ms: it is a ~150-instruction subroutine and every warp of every CTA
// runs it.  Request counts fit 32 bits -- lumcu_create refuses larger lists -- and size is 1 for a full update.)
__device__ __forceinline__ long long device_workgroups(const unsigned long long *nreq_dev, int size) {
    const unsigned n = (unsigned)__ldg(nreq_dev);
    return (long long)(size == 1 ? n : n / (unsigned)size);
}
__device__ __forceinline__ long long dispatch_end(const UpdateArgs &a) {
    long long e = a.wg_end;
    if (a.nreq_dev) { const long long n = device_workgroups(a.nreq_dev, a.size); e = n < e ? n : e; }
    return e;
}
// Does where this launch's window STARTS depend on the device-side count?  Only a contiguous split does (its shard
// boundaries are fractions of the dispatch's length).  Everywhere else a CTA knows its first workgroup from the launch
// arguments alone and can fetch its requests while the count is still on its way: the count is a global load at the very
// start of every CTA, and waiting for it before anything else cost the C4 update 0.13 ms.
__device__ __forceinline__ bool window_needs_count(const UpdateArgs &a) { return a.nreq_dev && a.il_n > 1 && a.il_chunk == 0; }
// The window of the dispatch this launch works in: [wg_begin, dispatch_end), or -- contiguous sharding -- the shard's
// own part of it, computed here because only the device may know where the dispatch ends.
struct ShardWindow { long long first, end; };
// floor(T * r / n) for T < 2^32, r <= n <= 16, in 32-bit arithmetic (every thread of every CTA evaluates this, and a
// 64-bit division is a ~120-instruction subroutine): T = q*n + m  =>  T*r/n = q*r + m*r/n.
__device__ __forceinline__ long long split_point(unsigned T, unsigned r, unsigned n) {
    const unsigned q = T / n, m = T - q * n;
    return (long long)q * r + (m * r) / n;
}
__device__ __forceinline__ ShardWindow shard_window(long long wg_begin, long long end, int il_chunk, int il_rank, int il_n) {
    if (il_n > 1 && il_chunk == 0) {
        const unsigned T = end > wg_begin ? (unsigned)(end - wg_begin) : 0u;
        return {wg_begin + split_point(T, (unsigned)il_rank, (unsigned)il_n), wg_begin + split_point(T, (unsigned)il_rank + 1u, (unsigned)il_n)};
    }
    return {wg_begin, end};
}
// a12's store, deferred: the packed texel of workgroup w goes to slot w of every rank's inbox
__device__ __forceinline__ void stage_word(const UpdateArgs &a, long long w, uint32_t word) {
    if (a.inbox.n > 1) {
#pragma unroll 4
        for (int p = 0; p < a.inbox.n; p++) a.inbox.p[p][w] = word;
    } else {
        a.staged[w] = word;
    }
}

// One CTA = one workgroup of radiance.comp (radiance.comp:360-414): R threads, one ray each.
// The subgroup reduction is the xor butterfly the oracle pins; _light[] lives in shared memory.
template <int R, bool DEBUG>
__global__ void __launch_bounds__(R) radiance_update_wg_kernel(SceneView s, UpdateArgs a) {
    constexpr int SUB = R < 32 ? R : 32;
    constexpr int NSUB = R / SUB;
    __shared__ float sh_light[NSUB][3];

    // The grid is capped (a launch must not depend on a request count that only the device knows): CTAs stride.
    const int ray = threadIdx.x;
    // the window from the host's bound first (no wait for the device-side count unless its start depends on it) ...
    ShardWindow sw = shard_window(a.wg_begin, window_needs_count(a) ? dispatch_end(a) : a.wg_end, a.il_chunk, a.il_rank, a.il_n);
    const bool valid = s.brick0_empty && !(__ldg(s.flags) & 1u);        // uniform
    for (long long i = blockIdx.x;; i += gridDim.x) {
    const long long w = il_map(sw.first, a.il_chunk, a.il_rank, a.il_n, i);
    if (w >= sw.end) return;                                     // uniform per CTA (end of the shard)
    const short4 rq = a.requests[w * (long long)a.size + a.shift];     // (inside max_requests even beyond the real count)
    // ... and the real end once the request is on its way
    if (i == (long long)blockIdx.x && !window_needs_count(a)) { const long long e = dispatch_end(a); sw.end = e < sw.end ? e : sw.end; if (w >= sw.end) return; }
    if (rq.z < a.z_begin || rq.z >= a.z_end) continue;        // uniform per CTA
    const uint32_t gid_x = (uint32_t)(w * R + ray);
    const f3 fib = {__ldg(a.dir_table + 3 * ray), __ldg(a.dir_table + 3 * ray + 1), __ldg(a.dir_table + 3 * ray + 2)};
    const f3 sun = {a.sun_x, a.sun_y, a.sun_z};

    RayResult r = valid ? trace_probe_ray<true>(s, rq.x, rq.y, rq.z, gid_x, ray, fib, a.time_seed, sun)
                        : trace_probe_ray<false>(s, rq.x, rq.y, rq.z, gid_x, ray, fib, a.time_seed, sun);

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
        stage_word(a, w, word);
    }
    __syncthreads();                                             // sh_light is reused by the next workgroup
    }
}

// ---- pooled kernel: the production path ---------------------------------------------------
// One CTA owns a batch of NR = RPT*THREADS rays (NR/R probes) staged in shared memory:
//   A  ray setup, one ray per thread per round, fully convergent.
//   B  coarse march with lane-level dynamic refill from the CTA's ray pool: a lane whose ray
//      ended immediately takes the next ray, so warps stay ~80 % full although ray lengths vary
//      1..67 samples.  When the pool is empty and a warp is down to < PARK live lanes it parks its unfinished
//      rays (their t goes back to the slot) and leaves.
//   B' the parked stragglers (~10 % of the rays, the long ones) are compacted into a few "tail"
//      warps, which march them in further passes among themselves (named barrier), while
//   C  every other warp already runs refinement + shading on the finished rays, 32 slots per
//      claim, convergent again.  The tail warps join C when they are done and finally shade the
//      stragglers.  So nobody idles behind the longest ray of the batch.
//   D  the workgroup reduction of radiance.comp:386-413 (xor butterfly per 32-ray subgroup, the
//      subgroups spread over the warps) and store_probe, one thread per probe.
// The result does not depend on the schedule: each ray's sample sequence is a function of its
// own (o, d, t) only, and the reduction tree is fixed.
struct PoolRay {
    f3 o, d, inv; float t; int ns; i3 cm;
};

// Ray slots, packed for 16-byte shared-memory accesses: a = (o.xyz, t), b = (d.xyz, inv.x),
// c = (inv.yz).  A refill is 2 x LDS.128 + 1 x LDS.64 instead of ten scalar loads; consecutive slots
// keep every quarter-warp conflict-free.  After shading, a.xyz holds the ray's light.
template <int NR, bool DEBUG>
struct PoolShared {
    float4 a[NR];
    float4 b[NR];
    float2 c[NR];
    unsigned char done[NR];               // 1 = terminated during pass 0 (shaded by the chunk loop); 0 = was parked
    unsigned short ns[DEBUG ? NR : 1];    // marching samples so far (instrumented variant only)
    __device__ __forceinline__ void load(int r, PoolRay &q) const {
        const float4 va = a[r], vb = b[r];
        const float2 vc = c[r];
        q.o = {va.x, va.y, va.z};
        q.d = {vb.x, vb.y, vb.z};
        q.inv = {vb.w, vc.x, vc.y};
        q.t = va.w;
        if constexpr (DEBUG) q.ns = ns[r];
        // march_step_fast's per-ray corner constants: float bits of 2^23 + (dir > 0 ? 16 : 0)
        q.cm = {0x4B000000 + (q.d.x > 0.0f ? 16 : 0), 0x4B000000 + (q.d.y > 0.0f ? 16 : 0),
                0x4B000000 + (q.d.z > 0.0f ? 16 : 0)};
    }
};

// March rays taken from pool positions [0, limit) (slot = src ? src[pos] : pos) until the pool is
// empty; then either finish the warp's remaining rays (can_park == false) or park them into
// dst[] (count in *dst_count) once fewer than PARK lanes are live.
// PRIVATE: the warp owns positions [first, limit) outright (pass 0: each warp marches its own slice of
// the batch, so a refill needs no shared counter, no atomic and no broadcast -- measured 2 % faster
// than the CTA-wide pool; the stragglers are balanced across warps by the tail passes anyway).
// FAST: march_step_fast (validated scene); a ray that ends without a hit stores -t_pre ("pending", see
// resolve_pending), otherwise -t.
template <int NR, int PARK, bool DEBUG, bool PRIVATE, bool FAST>
__device__ __forceinline__ void pool_march(PoolShared<NR, DEBUG> &sm, const SceneView &s, const MarchConsts &mc,
                                           int *next, int first, const unsigned short *src, int limit,
                                           unsigned short *dst, int *dst_count, bool can_park, int lane) {
    constexpr unsigned FULL = 0xffffffffu;
    const uint3 wsbits = {__float_as_uint(mc.wsx), __float_as_uint(mc.wsy), __float_as_uint(mc.wsz)};
    int cur = -1;
    PoolRay q = {};
    bool exhausted = false;
    int mine = first;                                              // PRIVATE: next unclaimed position
    for (;;) {
        unsigned idle = __ballot_sync(FULL, cur < 0);
        if (idle) {
            if (!exhausted) {
                const int cnt = __popc(idle);
                int base = mine;
                if constexpr (PRIVATE) {
                    mine += cnt;
                } else {
                    if (lane == 0) base = atomicAdd(next, cnt);
                    base = __shfl_sync(FULL, base, 0);
                }
                if (cur < 0) {
                    const int pos = base + __popc(idle & ((1u << lane) - 1u));
                    if (pos < limit) {
                        cur = src ? (int)src[pos] : pos;
                        sm.load(cur, q);
                    }
                }
                exhausted = base + cnt >= limit;
                if (exhausted) idle = __ballot_sync(FULL, cur < 0);   // otherwise every idle lane got a ray
            }
            if (exhausted) {
                const int act = 32 - __popc(idle);
                if (act == 0) break;
                if (can_park && act < PARK) {
                    int base = 0;
                    if (lane == 0) base = atomicAdd(dst_count, act);
                    base = __shfl_sync(FULL, base, 0);
                    if (cur >= 0) {
                        dst[base + __popc(~idle & ((1u << lane) - 1u))] = (unsigned short)cur;
                        sm.a[cur].w = q.t;
                        if constexpr (PRIVATE) sm.done[cur] = 0;   // pass 0 only: in the tail passes it already is 0, and the
                                                                    // other warps are reading done[] by then (phase C)
                        if constexpr (DEBUG) sm.ns[cur] = (unsigned short)q.ns;
                    }
                    break;
                }
            }
        }
        // Two march steps per refill check: with the 62-instruction step the ballot / refill / exit logic above is a
        // quarter of an iteration, and a lane that ends in the first step idles for one step only (B200, C4 world:
        // 1 step 20.55 ms, 2 steps 19.84, 3 steps 20.34, 4 steps 21.60; the pair fused, march_two_steps_fast: 18.96).
        if constexpr (FAST) {
            if (cur >= 0) {
                bool hit;
                float t_end;
                int steps;
                const bool done = march_two_steps_fast(s, wsbits, q.o, q.d, q.inv, q.cm, q.t, t_end, hit, steps);
                q.ns += steps;
                if (done) {
                    sm.a[cur].w = hit ? q.t : -t_end;              // what a ray that ends without a hit stores: -t_pre
                    if constexpr (DEBUG) sm.ns[cur] = (unsigned short)q.ns;
                    cur = -1;
                }
            }
        } else
#pragma unroll
        for (int rep = 0; rep < 2; rep++)
        if (cur >= 0) {
            const int st = march_step(s, mc, q.o, q.d, q.inv, q.t);
            q.ns++;
            if (st != 0) {
                sm.a[cur].w = st == 1 ? q.t : -q.t;                // sign = no hit (t >= 0.5 always)
                if constexpr (DEBUG) sm.ns[cur] = (unsigned short)q.ns;
                cur = -1;
            }
        }
    }
}

// Refinement + shading of slot r; the slot's o[] is replaced by the ray's light.
template <int NR, int R, bool DEBUG, bool FAST>
__device__ __forceinline__ void pool_shade(PoolShared<NR, DEBUG> &sm, const SceneView &s, const MarchConsts &mc,
                                           const UpdateArgs &a, long long wg0, int r, f3 sun,
                                           unsigned long long &my_s, unsigned long long &my_h) {
    PoolRay q;
    sm.load(r, q);
    bool hit = q.t > 0.0f;
    if constexpr (FAST) { if (!hit) hit = resolve_pending(s, mc, q.o, q.d, -q.t); }
    int nr;
    const f3 l = finish_ray<FAST>(s, q.o, q.d, q.inv, fabsf(q.t), hit, sun, nr);
    sm.a[r].x = l.x; sm.a[r].y = l.y; sm.a[r].z = l.z;
    if constexpr (DEBUG) {
        const unsigned nsamp = (unsigned)q.ns + (unsigned)nr;
        if (a.dbg_samples) a.dbg_samples[wg0 * R + r] = nsamp;
        if (a.dbg_hits) a.dbg_hits[wg0 * R + r] = hit ? 1 : 0;
        my_s += nsamp; my_h += hit ? 1 : 0;
    }
}

// Named barrier 1: the tail warps of a CTA among themselves (count = 32 x tail warps).
__device__ __forceinline__ void tail_barrier_sync(int count) {
#if defined(__CUDA_ARCH__)
    asm volatile("bar.sync 1, %0;" ::"r"(count) : "memory");
#elif defined(LUMCU_HOST_EMU)     // tests/device_host/simt_emu.hpp: the kernels run thread by thread on the host (tests only)
    lumcu_host_named_barrier(1, count);
#endif
}

#ifndef LUMCU_TAILDIV
#define LUMCU_TAILDIV 48       // stragglers per tail warp
#endif
#ifndef LUMCU_PARK
#define LUMCU_PARK 8          // tuned on B200 (C4 world): 8 beats 1/4/16/24/32
#endif
template <int R, int THREADS, int RPT, bool DEBUG>
struct PoolCta {
    static constexpr int NR = RPT * THREADS;        // rays per batch (RPT rays per thread)
    static constexpr int PPC = NR / R;              // probes per CTA
    static constexpr int NW = THREADS / 32;
    static constexpr int PARK = LUMCU_PARK;
    static constexpr int LIST = NW * PARK;
    static_assert(PPC >= 1, "batch smaller than one probe");
    PoolShared<NR, DEBUG> sm;
    unsigned short list[3][LIST];   // [0] = stragglers of pass 0, [1],[2] ping-pong afterwards
    short4 req[PPC];
    float sub[PPC * (R < 32 ? 1 : R / 32)][3];   // subgroup means (phase D)
    int next, park[3], chunk;
    long long wg0;                  // first workgroup of the batch (kept here, not in two registers across the march)
};

template <int R, int THREADS, int RPT, bool DEBUG, bool FAST>
__device__ __forceinline__ void pool_cta_body(PoolCta<R, THREADS, RPT, DEBUG> &cta, const SceneView &s, const UpdateArgs &a) {
    using Cta = PoolCta<R, THREADS, RPT, DEBUG>;
    constexpr int NR = Cta::NR, PPC = Cta::PPC, NW = Cta::NW, PARK = Cta::PARK;
    constexpr int SUB = R < 32 ? R : 32;
    constexpr int NSUB = R / SUB;
    constexpr unsigned FULL = 0xffffffffu;
    PoolShared<NR, DEBUG> &sm = cta.sm;
    unsigned short (&sm_list)[3][Cta::LIST] = cta.list;
    short4 (&sm_req)[PPC] = cta.req;
    int &sm_next = cta.next, &sm_chunk = cta.chunk;
    int (&sm_park)[3] = cta.park;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    static_assert(PPC <= THREADS, "one request per thread");
    int nprobes;
    {
        // the window from the host's bound first (no wait for the device-side count unless its start depends on it),
        // the batch's requests fetched, and only then the real end
        const bool early = window_needs_count(a);
        ShardWindow sw = shard_window(a.wg_begin, early ? dispatch_end(a) : a.wg_end, a.il_chunk, a.il_rank, a.il_n);
        const long long wg0 = il_map(sw.first, a.il_chunk, a.il_rank, a.il_n, (long long)blockIdx.x * PPC);   // il_chunk % PPC == 0
        long long left = sw.end - wg0;
        if (left <= 0) return;                                   // uniform per CTA (end of the shard)
        short4 rq = make_short4(0, 0, 0, 0);
        if (tid < PPC && tid < left) rq = a.requests[(wg0 + tid) * (long long)a.size + a.shift];    // (inside max_requests even beyond the real count)
        if (!early) { const long long e = dispatch_end(a); left = (e < sw.end ? e : sw.end) - wg0; if (left <= 0) return; }
        nprobes = left < PPC ? (int)left : PPC;
        if (tid < nprobes) sm_req[tid] = rq;
        if (tid == 0) { sm_next = 0; sm_park[0] = 0; sm_park[1] = 0; sm_park[2] = 0; sm_chunk = 0; cta.wg0 = wg0; }
    }
    const int nrays = nprobes * R;
    __syncthreads();
    const uint32_t gid0 = (uint32_t)(cta.wg0 * R);               // gl_GlobalInvocationID.x of the batch's first ray (mod 2^32, as the shader's uint)

    // ---- phase A: setup ----
    for (int r = tid; r < nrays; r += THREADS) {
        const int pr = r / R, ray = r % R;
        const short4 rq = sm_req[pr];
        const uint32_t gid_x = gid0 + (uint32_t)r;               // = (wg0 + pr) * R + ray
        const f3 fib = {__ldg(a.dir_table + 3 * ray), __ldg(a.dir_table + 3 * ray + 1), __ldg(a.dir_table + 3 * ray + 2)};
        f3 o, d, inv;
        setup_ray(rq.x, rq.y, rq.z, gid_x, ray, fib, a.time_seed, o, d, inv);
        sm.a[r] = make_float4(o.x, o.y, o.z, 0.0f);
        sm.b[r] = make_float4(d.x, d.y, d.z, inv.x);
        sm.c[r] = make_float2(inv.y, inv.z);
        sm.done[r] = 1;
        if constexpr (DEBUG) sm.ns[r] = 0;
    }
    __syncthreads();

    // ---- phase B, pass 0: every warp marches ----
    const MarchConsts mc = {(float)s.wx * 16.0f, (float)s.wy * 16.0f, (float)s.wz * 16.0f};
    {   // warp w owns slots [w*NR/NW, (w+1)*NR/NW) of the batch
        const int first = warp * (NR / NW), last = min(nrays, first + NR / NW);
        pool_march<NR, PARK, DEBUG, true, FAST>(sm, s, mc, nullptr, first, nullptr, last, sm_list[0], &sm_park[0], NW > 1, lane);
    }
    __syncthreads();
    const int parked0 = sm_park[0];
    const int tail_warps = parked0 == 0 ? 0 : max(1, min(NW, (parked0 + LUMCU_TAILDIV - 1) / LUMCU_TAILDIV));
    const int tail_threads = tail_warps * 32;

    // ---- phase B': the tail warps march the stragglers among themselves ----
    if (warp < tail_warps) {
        int limit = parked0, rd = 0;                       // list index read in this pass
        for (;;) {
            const int wr = rd == 1 ? 2 : 1;
            if (tid == 0) { sm_next = 0; sm_park[wr] = 0; }
            tail_barrier_sync(tail_threads);
            const int parts = max(1, min(tail_warps, (limit + LUMCU_TAILDIV - 1) / LUMCU_TAILDIV));
            if (warp < parts)
                pool_march<NR, PARK, DEBUG, false, FAST>(sm, s, mc, &sm_next, 0, sm_list[rd], limit, sm_list[wr], &sm_park[wr], parts > 1, lane);
            tail_barrier_sync(tail_threads);
            const int parked = sm_park[wr];
            if (parked == 0) break;
            limit = parked;
            rd = wr;
        }
    }

    // ---- phase C: refine + shade; 32 slots per claim ----
    const f3 sun = {a.sun_x, a.sun_y, a.sun_z};
    unsigned long long my_s = 0, my_h = 0;
    for (;;) {
        int c = 0;
        if (lane == 0) c = atomicAdd(&sm_chunk, 32);
        c = __shfl_sync(FULL, c, 0);
        if (c >= nrays) break;
        const int r = c + lane;
        if (r < nrays && sm.done[r]) pool_shade<NR, R, DEBUG, FAST>(sm, s, mc, a, cta.wg0, r, sun, my_s, my_h);
    }
    if (warp < tail_warps)                                 // the stragglers, now finished too
        for (int i = tid; i < parked0; i += tail_threads)
            pool_shade<NR, R, DEBUG, FAST>(sm, s, mc, a, cta.wg0, sm_list[0][i], sun, my_s, my_h);
    if constexpr (DEBUG) {
#pragma unroll
        for (int k = 16; k >= 1; k >>= 1) {
            my_s += __shfl_xor_sync(FULL, my_s, k);
            my_h += __shfl_xor_sync(FULL, my_h, k);
        }
        if (lane == 0) { atomicAdd(a.dbg_totals, my_s); atomicAdd(a.dbg_totals + 1, my_h); }
    }
    __syncthreads();

    // ---- phase D: reduce + blend ----
    // The PPC*NSUB subgroup sums are spread over the warps (xor butterfly, the oracle's pinned tree); then one
    // thread per probe forms the mean of its subgroups and blends, all probes of the batch side by side.
    for (int sg = warp; sg < nprobes * NSUB; sg += NW) {
        const int r = sg * SUB + lane;
        const float4 lv = lane < SUB ? sm.a[r] : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        float lx = lv.x, ly = lv.y, lz = lv.z;
#pragma unroll
        for (int k = SUB / 2; k >= 1; k >>= 1) {
            lx += __shfl_xor_sync(FULL, lx, k);
            ly += __shfl_xor_sync(FULL, ly, k);
            lz += __shfl_xor_sync(FULL, lz, k);
        }
        if (lane == 0) { cta.sub[sg][0] = lx / (float)SUB; cta.sub[sg][1] = ly / (float)SUB; cta.sub[sg][2] = lz / (float)SUB; }
    }
    __syncthreads();
    if (tid < nprobes) {
        const int pr = tid;
        float fr = 0.0f, fg = 0.0f, fb = 0.0f;
#pragma unroll
        for (int g = 0; g < NSUB; g++) { fr += cta.sub[pr * NSUB + g][0]; fg += cta.sub[pr * NSUB + g][1]; fb += cta.sub[pr * NSUB + g][2]; }
        fr /= (float)NSUB; fg /= (float)NSUB; fb /= (float)NSUB;
        const long long w = cta.wg0 + pr;
        const short4 rq = sm_req[pr];
        if constexpr (DEBUG) {
            if (a.dbg_pre_blend) {
                a.dbg_pre_blend[3 * w] = fr; a.dbg_pre_blend[3 * w + 1] = fg; a.dbg_pre_blend[3 * w + 2] = fb;
            }
        }
        uint32_t word = 0;
        if ((unsigned)rq.x < (unsigned)s.wx && (unsigned)rq.y < (unsigned)s.wy && (unsigned)rq.z < (unsigned)s.wz) {
            f3 old = unpack10(__ldg(s.field + (rq.x + s.wx * (rq.y + s.wy * rq.z))));
            old.x = gmin(gmax(old.x, 0.0f), 1.0f); old.y = gmin(gmax(old.y, 0.0f), 1.0f); old.z = gmin(gmax(old.z, 0.0f), 1.0f);
            const float t = a.reactiveness, omt = 1.0f - t;
            word = pack10(old.x * omt + fr * t, old.y * omt + fg * t, old.z * omt + fb * t);
        }
        stage_word(a, w, word);
    }
}

#ifdef LUMCU_MINB              // tuning builds (tools/final_sweep.sh): a floor of resident CTAs per SM, i.e. a register cap
#define LUMCU_POOL_BOUNDS(T) __launch_bounds__(T, LUMCU_MINB)
#else
#define LUMCU_POOL_BOUNDS(T) __launch_bounds__(T)
#endif
template <int R, int THREADS, int RPT, bool DEBUG>
__global__ void LUMCU_POOL_BOUNDS(THREADS) radiance_update_pool_kernel(SceneView s, UpdateArgs a) {
    __shared__ PoolCta<R, THREADS, RPT, DEBUG> cta;
    // The fast march needs brick 0 empty and every world id inside the palette; the second is a property of
    // the uploaded data, checked on the device (validate_world_kernel) so that no host sync is needed.
    const bool general = !s.brick0_empty || (__ldg(s.flags) & 1u);
    if (general) pool_cta_body<R, THREADS, RPT, DEBUG, false>(cta, s, a);
    else pool_cta_body<R, THREADS, RPT, DEBUG, true>(cta, s, a);
}

// bit 0 of *flags: some world id lies outside [0, capacity) (such a block marches as "non-empty, all voxels
// empty", radiance_device.cuh get_voxel).  *flags is zeroed before the launch.
__global__ void validate_world_kernel(const int16_t *__restrict__ world, long long cells, int capacity,
                                      uint32_t *__restrict__ flags) {
    bool bad = false;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < cells; i += (long long)gridDim.x * blockDim.x)
        bad |= (unsigned)world[i] >= (unsigned)capacity;
    if (__syncthreads_or(bad) && threadIdx.x == 0) atomicOr(flags, 1u);
}

// Barrier flags of the multi-GPU exchange: flags[r] of a rank holds the epoch rank r last arrived at (written by
// rank r's arrive_kernel over NVLink, only ever growing).
struct PeerFlags { unsigned *f[MAX_PEERS]; int n; int rank; };
struct ExchangeWait {
    const volatile unsigned *flags;        // this rank's flag array, or null: no exchange (single GPU)
    int n;
    unsigned epoch;
    long long timeout_cycles;
    volatile unsigned *status;             // set to 1 when a rank never arrived
};

// flag reads of the exchange: acquire at system scope (what follows cannot be read before the flag)
__device__ __forceinline__ unsigned load_acquire_sys(const volatile unsigned *p) {
#ifdef __CUDA_ARCH__
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
#else
    const unsigned v = *p;
    __threadfence_system();
    return v;
#endif
}

// Deferred imageStore (radiance.comp:353): scatter the staged texels of workgroups [wg_begin, wg_end) into the field
// once every read of the old field is done.  In a multi-GPU frame the staged array is this rank's inbox: every rank
// (this one included) stored the texels of its own workgroups there during its trace kernel, and xw names the epoch
// all of them must have arrived at before the inbox is complete -- thread 0 of each CTA waits for the flags (a rank
// that never arrives trips the timeout instead of hanging the GPU), the commit itself is local.
struct Interleave { int chunk, rank, n; };
__global__ void commit_kernel(uint32_t *__restrict__ field, const uint32_t *__restrict__ staged,
                              const short4 *__restrict__ requests, long long wg_begin, long long wg_end,
                              int size, int shift, int wx, int wy, int wz, int z_begin, int z_end, Interleave il,
                              const unsigned long long *__restrict__ nreq_dev, ExchangeWait xw) {
    if (xw.flags) {
        if (threadIdx.x == 0) {
            const long long t0 = clock64();
            for (int p = 0; p < xw.n; p++)
                while ((int)(load_acquire_sys(xw.flags + p) - xw.epoch) < 0) {
                    __nanosleep(100);
                    if (clock64() - t0 > xw.timeout_cycles) { *xw.status = 1u; break; }
                }
        }
        __syncthreads();
    }
    if (nreq_dev) { const long long n = device_workgroups(nreq_dev, size); wg_end = n < wg_end ? n : wg_end; }
    const ShardWindow sw = shard_window(wg_begin, wg_end, il.chunk, il.rank, il.n);
    long long w = il_map(sw.first, il.chunk, il.rank, il.n, (long long)blockIdx.x * blockDim.x + threadIdx.x);
    if (w >= sw.end) return;
    short4 rq = requests[w * (long long)size + shift];
    if (rq.z < z_begin || rq.z >= z_end) return;
    if ((unsigned)rq.x < (unsigned)wx && (unsigned)rq.y < (unsigned)wy && (unsigned)rq.z < (unsigned)wz)
        field[rq.x + wx * (rq.y + wy * rq.z)] = __ldcv(staged + w);      // written by peers: not through a stale L1 line
}

// The same scatter for the plain single-GPU update of a full dispatch (size 1, no shard, no z filter), four workgroups
// per thread: the request rule emits probes in z,y,x order, so four consecutive requests are almost always four
// x-adjacent texels -- one 16-byte store where they are, and aligned; element by element where not (row ends, specials,
// probes outside the world).  (north_star: "coalesced vectorised stores into the radiance field".)
__global__ void commit_vec4_kernel(uint32_t *__restrict__ field, const uint32_t *__restrict__ staged,
                                   const short4 *__restrict__ requests, long long wg_end, int wx, int wy, int wz,
                                   const unsigned long long *__restrict__ nreq_dev) {
    if (nreq_dev) { const long long n = device_workgroups(nreq_dev, 1); wg_end = n < wg_end ? n : wg_end; }
    const long long w0 = 4 * ((long long)blockIdx.x * blockDim.x + threadIdx.x);
    if (w0 >= wg_end) return;
    if (w0 + 3 < wg_end) {
        const uint4 q01 = reinterpret_cast<const uint4 *>(requests + w0)[0], q23 = reinterpret_cast<const uint4 *>(requests + w0)[1];
        const uint4 v = *reinterpret_cast<const uint4 *>(staged + w0);
        // a request = (x, y, z, w) as four int16: x | y << 16 in the first word, z in the low half of the second
        const int x0 = (short)(q01.x & 0xffffu), y0 = (short)(q01.x >> 16), z0 = (short)(q01.y & 0xffffu);
        const bool row = q01.z == q01.x + 1u && q23.x == q01.x + 2u && q23.z == q01.x + 3u &&            // x, x+1, x+2, x+3, same y
                         (q01.w & 0xffffu) == (q01.y & 0xffffu) && (q23.y & 0xffffu) == (q01.y & 0xffffu) && (q23.w & 0xffffu) == (q01.y & 0xffffu);
        if (row && x0 >= 0 && x0 + 3 < wx && (unsigned)y0 < (unsigned)wy && (unsigned)z0 < (unsigned)wz) {
            const long long c0 = x0 + (long long)wx * (y0 + (long long)wy * z0);
            if ((c0 & 3) == 0) { *reinterpret_cast<uint4 *>(field + c0) = v; return; }
        }
    }
    for (long long w = w0; w < wg_end && w < w0 + 4; w++) {
        const short4 rq = requests[w];
        if ((unsigned)rq.x < (unsigned)wx && (unsigned)rq.y < (unsigned)wy && (unsigned)rq.z < (unsigned)wz)
            field[rq.x + wx * (rq.y + wy * rq.z)] = staged[w];
    }
}

// ---- the exchange as a PULL: all-gather and commit in one kernel ------------------------------------------------
// lumcu_update_shard_exchange: every rank traces its shard into its OWN staged[] (nothing remote in the trace kernel:
// fine-grained peer stores from its last phase cost the trace 2 % at 8 ranks), publishes its arrival, and then this
// kernel -- behind the wait for all flags -- walks the whole dispatch: the texel of workgroup w is read from the
// staged[] of the rank that owns w (over NVLink unless it is this rank; consecutive workgroups have one owner, so the
// remote reads are full 128-byte lines) and stored into the local field.  A persistent grid: each CTA waits for the
// flags once.
struct ShardSources { PeerWords staged; int chunk, n; };
__global__ void __launch_bounds__(256) gather_commit_kernel(uint32_t *__restrict__ field, ShardSources src,
                                                            const short4 *__restrict__ requests, long long wg_begin, long long wg_end,
                                                            int size, int shift, int wx, int wy, int wz,
                                                            const unsigned long long *__restrict__ nreq_dev, ExchangeWait xw) {
    __shared__ long long bound[MAX_PEERS + 1];
    if (threadIdx.x == 0) {
        const long long t0 = clock64();
        for (int p = 0; p < xw.n; p++)
            while ((int)(load_acquire_sys(xw.flags + p) - xw.epoch) < 0) {
                __nanosleep(100);
                if (clock64() - t0 > xw.timeout_cycles) { *xw.status = 1u; break; }
            }
    }
    if (nreq_dev) { const long long n = device_workgroups(nreq_dev, size); wg_end = n < wg_end ? n : wg_end; }
    if ((int)threadIdx.x <= src.n) {                               // where shard r of the contiguous split starts
        const ShardWindow sw = shard_window(wg_begin, wg_end, 0, (int)threadIdx.x, src.n);
        bound[threadIdx.x] = src.n > 1 ? sw.first : (threadIdx.x == 0 ? wg_begin : wg_end);
    }
    __syncthreads();
    for (long long w = wg_begin + (long long)blockIdx.x * blockDim.x + threadIdx.x; w < wg_end; w += (long long)gridDim.x * blockDim.x) {
        int r = 0;
        if (src.chunk == 0) {
            for (int k = 1; k < src.n; k++) r += w >= bound[k] ? 1 : 0;
        } else {
            r = (int)(((unsigned)(w - wg_begin) / (unsigned)src.chunk) % (unsigned)src.n);
        }
        const uint32_t word = __ldcv(src.staged.p[r] + w);
        const short4 rq = requests[w * (long long)size + shift];
        if ((unsigned)rq.x < (unsigned)wx && (unsigned)rq.y < (unsigned)wy && (unsigned)rq.z < (unsigned)wz)
            field[rq.x + wx * (rq.y + wy * rq.z)] = word;
    }
}

// "This rank has finished tracing epoch `epoch`, and everything it stored into the other ranks' inboxes is visible":
// runs right after the trace kernel on the same stream (its stores are complete when this starts; the system-scope
// fence orders them before the flag for any observer), lane p publishes the epoch in rank p's flag array.
__global__ void arrive_kernel(PeerFlags pf, unsigned epoch) {
    const int p = threadIdx.x;
    __threadfence_system();
    if (p < pf.n) {
        volatile unsigned *dst = pf.f[p] + pf.rank;
        *dst = epoch;
    }
    __threadfence_system();
}

// ---- glossy reflections: shaders/glossy.frag as a compute pass ---------------------
// One thread per fragment; a warp covers an 8 x 4 pixel tile (neighbouring pixels see the same surface, so
// their mirror rays stay together), a CTA of 128 threads a 32 x 4 strip.
// glossy.frag:439-491 main(); the stencil test of the pass (fillStencilGlossy.frag: discard if roughness > 0.5)
// is evaluated here instead of by the ROP.
__global__ void __launch_bounds__(128) glossy_kernel(SceneView s, GlossyArgs a) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int px = blockIdx.x * 32 + warp * 8 + (lane & 7), py = blockIdx.y * 4 + (lane >> 3);
    if (px >= a.width || py >= a.height) return;
    glossy_fragment(s, a, px, py);
}

// shift_radiance (internal_render.cpp:1218-1335): field[p] = tmp[p + d] on the overlap box.
__global__ void shift_kernel(uint32_t *__restrict__ field, const uint32_t *__restrict__ tmp,
                             int wx, int wy, int wz, int dx, int dy, int dz) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long cells = (long long)wx * wy * wz;
    if (i >= cells) return;
    const unsigned row = (unsigned)i / (unsigned)wx;                // cells fit 31 bits
    int x = (int)((unsigned)i - row * (unsigned)wx), z = (int)(row / (unsigned)wy), y = (int)(row - (unsigned)z * (unsigned)wy);
    int sx = x + dx, sy = y + dy, sz = z + dz;
    if ((unsigned)sx < (unsigned)wx && (unsigned)sy < (unsigned)wy && (unsigned)sz < (unsigned)wz)
        field[i] = tmp[sx + wx * (sy + (long long)wy * sz)];
}

__global__ void unpack_table_kernel(float *__restrict__ out) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < 1024) out[c] = unorm10_to_float((uint32_t)c);
}


// ---- feeders: blockify brick copies + map.comp ------------------------------------------
// exec_copies' per-block copies (internal_render.cpp:1402-1446): one CTA per brick, 256 threads x 16 B.
__global__ void copy_bricks_kernel(uint8_t *__restrict__ bricks, const int2 *__restrict__ pairs, int n, int capacity) {
    const int i = blockIdx.x;
    if (i >= n) return;
    const int2 pr = pairs[i];                                 // x = src, y = dst
    if ((unsigned)pr.y >= (unsigned)capacity) return;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (pr.x > 0 && pr.x < capacity) v = reinterpret_cast<const uint4 *>(bricks + (size_t)pr.x * 4096u)[threadIdx.x];
    reinterpret_cast<uint4 *>(bricks + (size_t)pr.y * 4096u)[threadIdx.x] = v;
}

// shaders/map.comp:32-67.  One thread per invocation; z is the fastest thread index so that the
// conditional stores of a warp fall into one 16-byte z-row of a brick where possible.
__global__ void map_model_kernel(const int16_t *__restrict__ world, uint8_t *__restrict__ bricks,
                                 const uint8_t *__restrict__ model, int wx, int wy, int wz, int capacity, MapArgs a) {
    const long long total = (long long)a.ext[0] * a.ext[1] * a.ext[2];
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int gz = (int)(i % a.ext[2]), gy = (int)((i / a.ext[2]) % a.ext[1]), gx = (int)(i / ((long long)a.ext[2] * a.ext[1]));
    map_model_invocation(world, bricks, model, wx, wy, wz, capacity, a, gx, gy, gz);
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

// ---- the rule-built part of the list in ONE pass ---------------------------------------------------------------
// Count -> scan -> scatter as three kernels (round 1) spent most of their 114 us (C4 world) between kernels and in a
// single-CTA scan; one kernel does all three with a decoupled look-back (Merrill & Garland): a CTA takes the next tile of
// BUILD_TILE consecutive cells (a ticket, so that a tile's predecessors are always already running or done), evaluates
// the 3x3x3 rule for them, publishes its count, looks back over its predecessors' published counts / prefixes until it
// knows how many requests precede its tile, publishes its own inclusive prefix and writes its requests -- in cell
// order, so the list is z,y,x-sorted exactly like the reference's loop produces it.
// tile_state[t]: bits 63..62 = 0 nothing yet, 1 the tile's own count, 2 the inclusive prefix; bits 61..0 = the number.
constexpr int BUILD_ITEMS = 4;                                      // x-consecutive cells per thread and group
constexpr int BUILD_GROUPS = 4;                                     // groups of BUILD_THREADS * BUILD_ITEMS cells per tile
constexpr int BUILD_GROUP_CELLS = BUILD_THREADS * BUILD_ITEMS;
constexpr int BUILD_TILE = BUILD_GROUP_CELLS * BUILD_GROUPS;         // 4096 cells: the ticket, the barriers and the look-back
                                                                    // are paid once per tile, and a CTA lives ~8 us whatever it holds
// cell index -> (x, y, z); cell indices fit 31 bits (lumcu_create), so 32-bit divisions: the 64-bit ones are ~120-instruction
// subroutines, and fifteen of them per thread were most of this kernel's first version
struct Cell3 { int x, y, z; };
__device__ __forceinline__ Cell3 cell_of(unsigned i, unsigned wx, unsigned wy) {
    const unsigned row = i / wx, z = row / wy;
    return {(int)(i - row * wx), (int)(row - z * wy), (int)z};
}
__global__ void __launch_bounds__(BUILD_THREADS) build_fused_kernel(const int16_t *__restrict__ world, int wx, int wy, int wz,
                                                                    uint8_t *__restrict__ flags, unsigned long long *tile_state,
                                                                    unsigned int *ticket, short4 *__restrict__ out, long long cap,
                                                                    unsigned long long *__restrict__ total) {
    __shared__ unsigned int sm_tile;
    __shared__ int warp_counts[BUILD_GROUPS][BUILD_THREADS / 32];
    __shared__ unsigned long long sm_part[BUILD_THREADS / 32];
    __shared__ int sm_stop[BUILD_THREADS / 32];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) sm_tile = atomicAdd(ticket, 1u);
    __syncthreads();
    const long long tile = sm_tile;
    const long long cells = (long long)wx * wy * wz, ntiles = (cells + BUILD_TILE - 1) / BUILD_TILE;
    if (tile >= ntiles) return;
    // A thread owns, in each of the tile's BUILD_GROUPS groups, BUILD_ITEMS = 4 x-consecutive cells.  Where rows are a
    // multiple of 4 long the four lie in one row and each of the 9 neighbour rows is one 8-byte load plus the two cells
    // either side: 27 loads for 4 cells instead of 108, and the sums of the three x-neighbours are shared.  Other
    // worlds take the cell-by-cell rule.
    const unsigned uwx = (unsigned)wx, uwy = (unsigned)wy, ucells = (unsigned)cells;
    const unsigned t0 = (unsigned)tile * BUILD_TILE + (unsigned)tid * BUILD_ITEMS;      // + g * BUILD_GROUP_CELLS
    unsigned mine = 0;                 // bit 4g + k: cell t0 + g * BUILD_GROUP_CELLS + k is selected
    int before[BUILD_GROUPS];          // selected cells of the group in front of this thread's (inside its warp, for now)
#pragma unroll
    for (int g = 0; g < BUILD_GROUPS; g++) {
        const unsigned i0 = t0 + g * BUILD_GROUP_CELLS;
        unsigned m = 0;
        if (i0 < ucells) {
            if ((wx & 3) == 0) {
                const Cell3 c = cell_of(i0, uwx, uwy);
                int s0 = 0, s1 = 0, s2 = 0, s3 = 0;
                const int xl = c.x > 0 ? c.x - 1 : 0, xr = c.x + 4 < wx ? c.x + 4 : wx - 1;
#pragma unroll
                for (int dz = -1; dz <= 1; dz++)
#pragma unroll
                    for (int dy = -1; dy <= 1; dy++) {
                        const int cy = min(max(c.y + dy, 0), wy - 1), cz = min(max(c.z + dz, 0), wz - 1);
                        const int16_t *row = world + uwx * ((unsigned)cy + uwy * (unsigned)cz);
                        const short4 v = __ldg(reinterpret_cast<const short4 *>(row + c.x));
                        const int l = __ldg(row + xl), r = __ldg(row + xr);
                        s0 += l + v.x + v.y; s1 += v.x + v.y + v.z; s2 += v.y + v.z + v.w; s3 += v.z + v.w + r;
                    }
                m = (s0 > 0 ? 1u : 0u) | (s1 > 0 ? 2u : 0u) | (s2 > 0 ? 4u : 0u) | (s3 > 0 ? 8u : 0u);
                *reinterpret_cast<uchar4 *>(flags + i0) = make_uchar4(m & 1u, (m >> 1) & 1u, (m >> 2) & 1u, (m >> 3) & 1u);
            } else {
#pragma unroll
                for (int k = 0; k < BUILD_ITEMS; k++) {
                    const unsigned i = i0 + k;
                    if (i < ucells) {
                        const Cell3 c = cell_of(i, uwx, uwy);
                        const bool sel = probe_selected(world, wx, wy, wz, c.x, c.y, c.z);
                        flags[i] = sel ? 1 : 0;
                        m |= sel ? (1u << k) : 0u;
                    }
                }
            }
        }
        mine |= m << (4 * g);
        const int cnt = __popc(m);
        int incl = cnt;                                             // inclusive scan of the group's counts inside the warp
#pragma unroll
        for (int k = 1; k < 32; k <<= 1) {
            const int n = __shfl_up_sync(0xffffffffu, incl, k);
            if (lane >= k) incl += n;
        }
        if (lane == 31) warp_counts[g][wid] = incl;
        before[g] = incl - cnt;
    }
    __syncthreads();
    // rank inside the tile: groups in order, threads in order inside a group, a thread's cells in order
    int count = 0;
#pragma unroll
    for (int g = 0; g < BUILD_GROUPS; g++) {
        before[g] += count;
#pragma unroll
        for (int w = 0; w < BUILD_THREADS / 32; w++) { if (w < wid) before[g] += warp_counts[g][w]; count += warp_counts[g][w]; }
    }
    // The whole CTA looks back, BUILD_THREADS tiles per round: thread j reads tile - 1 - j (- BUILD_THREADS per round),
    // each warp sums the counts of the tiles in front of the nearest one that already knows its inclusive prefix, and
    // that prefix; the warps' results are combined in order.
    unsigned long long prefix = 0;
    {
        volatile unsigned long long *st = tile_state;
        const unsigned long long VAL = (1ull << 62) - 1ull;
        if (tile > 0) {
            if (tid == 0) st[tile] = (1ull << 62) | (unsigned long long)count;
            for (long long p0 = tile - 1;; p0 -= BUILD_THREADS) {
                const long long p = p0 - tid;
                unsigned long long v = 2ull << 62;                   // in front of tile 0: an inclusive prefix of zero
                if (p >= 0) while (((v = st[p]) >> 62) == 0) {}
                const unsigned have = __ballot_sync(0xffffffffu, (v >> 62) == 2);
                const int stop = have ? __ffs(have) - 1 : 32;          // the nearest tile with a prefix ends the walk
                unsigned long long part = lane <= stop ? (v & VAL) : 0ull;
#pragma unroll
                for (int k = 16; k >= 1; k >>= 1) part += __shfl_xor_sync(0xffffffffu, part, k);
                __syncthreads();                                     // (sm_part / sm_stop of the previous round have been read)
                if (lane == 0) { sm_part[wid] = part; sm_stop[wid] = stop; }
                __syncthreads();
                bool found = false;
#pragma unroll
                for (int w = 0; w < BUILD_THREADS / 32; w++)
                    if (!found) { prefix += sm_part[w]; found = sm_stop[w] < 32; }
                if (found) break;                                    // uniform: every thread read the same arrays
            }
        }
        if (tid == 0) {
            st[tile] = (2ull << 62) | (prefix + (unsigned long long)count);
            if (tile == ntiles - 1) { total[0] = prefix + (unsigned long long)count; total[1] = total[0]; }
        }
    }
#pragma unroll
    for (int g = 0; g < BUILD_GROUPS; g++) {
        const unsigned m = (mine >> (4 * g)) & 15u;
        if (m) {
            long long pos = (long long)prefix + before[g];
            Cell3 c = cell_of(t0 + g * BUILD_GROUP_CELLS, uwx, uwy);
#pragma unroll
            for (int k = 0; k < BUILD_ITEMS; k++) {
                if ((m >> k) & 1u) {
                    if (pos < cap) out[pos] = make_short4((short)c.x, (short)c.y, (short)c.z, 0);
                    pos++;
                }
                if (++c.x >= wx) { c.x = 0; if (++c.y >= wy) { c.y = 0; c.z++; } }      // the next cell: carry into y, z
            }
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
