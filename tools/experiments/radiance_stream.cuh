// SHELVED EXPERIMENT -- not part of the build (lum_b200/build.py does not see this file).
// Result (B200, C4 world, profiles/r01_ncu_streaming_kernel_rejected.txt): bit-exact in the whole parity
// suite, but 28.2 ms against the pooled kernel's 21.0 ms.  March lanes are 96 % full and there are no barriers,
// as intended; the price is ~100 instructions of warp-uniform queue control in EVERY iteration (with 4.5 rays
// ending per step some lane is always idle, so the control path never gets skipped): 2,626 warp-instructions
// per 32 rays against 2,221, at 56 registers (9 CTAs/SM; capped at 48 it spills the ray into local memory:
// 37 ms).  Even an ideal 60-instruction control path would only break even.  Kept for the record; to try it
// again: #include it from lumcu.cu and launch radiance_update_stream_kernel on a persistent grid with a
// [grid * 4 warps][8][64] float4 light ring.
//
// radiance_stream.cuh -- the streaming radiance kernel (sm_100a): every WARP is an autonomous ray engine.
//
// The pooled kernel (radiance_kernels.cuh) gives each CTA a fixed batch of rays; when a warp's slice runs
// dry its last long rays are parked, compacted and finished by tail warps behind CTA barriers.  ncu put a
// price on that: ~490 of 2,220 warp-instructions per 32 rays go to refill/park/tail control, march lanes
// are 82 % full, and 16 % of all warp stall time is barrier wait.  Here nothing ever runs dry:
//
//   * a persistent grid; warp gw of TW walks the probes gw, gw + TW, gw + 2 TW, ... of the dispatch;
//   * per warp, in shared memory: S ray slots and three byte queues -- READY (set up, waiting for a lane),
//     FINISHED (march ended, waiting to be shaded), FREE;
//   * each loop iteration the warp, with warp-uniform decisions only,
//       - shades a batch of 32 finished rays when there are 32 (refinement + gather, fully convergent),
//       - sets up the next 32 rays (half a probe) when READY runs low and 32 slots are free (convergent),
//       - refills idle lanes from READY and takes one march step;
//   * a shaded ray's light goes to a per-warp ring in global memory (L2-resident, 16 B per ray); when the
//     64th ray of a probe is shaded the warp reduces it with the oracle's pinned butterfly and stages the
//     blended texel.  Long rays simply stay in their lanes while newer probes stream past them.
//
// No __syncthreads, no atomics, no parking.  Results cannot depend on the schedule for the same reason as
// in the pooled kernel: a ray's samples are a function of its own (o, d, t), and the reduction tree is fixed.
#pragma once
#include "radiance_kernels.cuh"

namespace lumcu {

constexpr int STREAM_S = 96;      // ray slots per warp: 32 marching + < 32 ready + < 32 finished
constexpr int STREAM_P = 8;       // probes in flight per warp (ring of light buffers)
constexpr int STREAM_WARPS = 4;   // warps per CTA (they share nothing but the CTA)

template <bool DEBUG>
struct StreamWarp {
    float4 a[STREAM_S];                    // o.xyz, t         (PoolShared's slot layout)
    float4 b[STREAM_S];                    // d.xyz, inv.x
    float2 c[STREAM_S];                    // inv.yz
    unsigned short pid[STREAM_S];          // ring entry << 8 | ray index within the probe
    unsigned short ns[DEBUG ? STREAM_S : 1];
    unsigned char rq[128], fq[128];        // READY / FINISHED: FIFO rings of slot numbers (indices mod 128)
    unsigned char fl[STREAM_S];            // FREE: stack of slot numbers
    short4 req[STREAM_P];                  // per ring entry: the probe request,
    long long wg[STREAM_P];                //   its workgroup index,
    int cnt[STREAM_P];                     //   rays shaded so far
    __device__ __forceinline__ void load(int r, PoolRay &q) const {
        const float4 va = a[r], vb = b[r];
        const float2 vc = c[r];
        q.o = {va.x, va.y, va.z};
        q.d = {vb.x, vb.y, vb.z};
        q.inv = {vb.w, vc.x, vc.y};
        q.t = va.w;
        if constexpr (DEBUG) q.ns = ns[r];
        q.cm = {0x4B000000 + (q.d.x > 0.0f ? 16 : 0), 0x4B000000 + (q.d.y > 0.0f ? 16 : 0),
                0x4B000000 + (q.d.z > 0.0f ? 16 : 0)};
    }
};

struct StreamArgs {
    float4 *__restrict__ lights;           // [total warps][STREAM_P][64] light per shaded ray (x,y,z,-)
};

// Warp-uniform loop state.  Registers are what this kernel is short of (the ray in flight takes 15, and 48 is
// the budget for 10 CTAs per SM), so the queue cursors travel packed in two registers and everything that is
// only needed when a batch is shaded or set up (ring pointer, probe stride, scene constants) is recomputed
// there from kernel parameters and special registers.
//   qs: rhead | rtail << 8 | fhead << 16 | ftail << 24   (cursors mod 256; queue rings are 128 entries)
//   ms: free_n | half << 8 | cur_e << 9 | ring_used << 16 (STREAM_P <= 8)
struct StreamState { unsigned qs, ms; unsigned k; };   // k: probes this warp has set up completely
__device__ __forceinline__ int st_ready(const StreamState &u) { return ((u.qs >> 8) - u.qs) & 255; }
__device__ __forceinline__ int st_finished(const StreamState &u) { return ((u.qs >> 24) - (u.qs >> 16)) & 255; }
__device__ __forceinline__ int st_free(const StreamState &u) { return u.ms & 255; }
__device__ __forceinline__ long long st_gw() { return (long long)blockIdx.x * STREAM_WARPS + (threadIdx.x >> 5); }
__device__ __forceinline__ long long st_next_w(const StreamState &u, const UpdateArgs &a) {
    return a.wg_begin + st_gw() + (long long)u.k * ((long long)gridDim.x * STREAM_WARPS);
}

// ---- shade n (<= 32) rays from the head of FINISHED ---------------------------------------------------
template <bool DEBUG, bool FAST>
__device__ __forceinline__ StreamState stream_shade(StreamWarp<DEBUG> &sw, const SceneView &s, const UpdateArgs &a,
                                                 const StreamArgs &sa, StreamState u, int n) {
    constexpr int R = 64, P = STREAM_P;
    constexpr unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    float4 *const ring = sa.lights + st_gw() * (P * R);
    const MarchConsts mc = {(float)s.wx * 16.0f, (float)s.wy * 16.0f, (float)s.wz * 16.0f};
    const f3 sun = {a.sun_x, a.sun_y, a.sun_z};
    const int fhead = (u.qs >> 16) & 255, free_n = st_free(u);
    int e = -1;
    unsigned long long my_s = 0, my_h = 0;
    if (lane < n) {
        const int slot = sw.fq[(fhead + lane) & 127];
        PoolRay r;
        sw.load(slot, r);
        bool hit = r.t > 0.0f;
        if constexpr (FAST) { if (!hit) hit = resolve_pending(s, mc, r.o, r.d, -r.t); }
        int nr;
        const f3 l = finish_ray<FAST>(s, r.o, r.d, r.inv, fabsf(r.t), hit, sun, nr);
        const int id = sw.pid[slot];
        e = id >> 8;
        const int ray = id & 255;
        ring[e * R + ray] = make_float4(l.x, l.y, l.z, 0.0f);
        if constexpr (DEBUG) {
            const unsigned nsamp = (unsigned)r.ns + (unsigned)nr;
            const long long w = sw.wg[e];
            if (a.dbg_samples) a.dbg_samples[w * R + ray] = nsamp;
            if (a.dbg_hits) a.dbg_hits[w * R + ray] = hit ? 1 : 0;
            my_s += nsamp; my_h += hit ? 1 : 0;
        }
        sw.fl[free_n + lane] = (unsigned char)slot;
    }
    u.qs = (u.qs & 0xff00ffffu) | (((unsigned)(fhead + n) & 255u) << 16);
    u.ms += (unsigned)n;                                     // free_n += n (never carries: free_n <= STREAM_S)
    // per-probe completion counts: one writer per ring entry
    const unsigned peers = __match_any_sync(FULL, e);
    const bool leader = e >= 0 && (peers & lt) == 0;
    bool complete = false;
    if (leader) {
        const int c = sw.cnt[e] + __popc(peers);
        sw.cnt[e] = c;
        complete = c == R;
    }
    __syncwarp();                                            // lights and counts visible to the whole warp
    unsigned todo = __ballot_sync(FULL, complete);
    while (todo) {
        const int src = __ffs(todo) - 1;
        todo &= todo - 1;
        const int pe = __shfl_sync(FULL, e, src);
        // reduction (radiance.comp:386-407): xor butterfly per 32-ray subgroup, mean of the two subgroups
        float fr = 0.0f, fg = 0.0f, fb = 0.0f;
#pragma unroll
        for (int g = 0; g < 2; g++) {
            const float4 lv = __ldcg(ring + pe * R + g * 32 + lane);
            float lx = lv.x, ly = lv.y, lz = lv.z;
#pragma unroll
            for (int k = 16; k >= 1; k >>= 1) {
                lx += __shfl_xor_sync(FULL, lx, k);
                ly += __shfl_xor_sync(FULL, ly, k);
                lz += __shfl_xor_sync(FULL, lz, k);
            }
            fr += lx / 32.0f; fg += ly / 32.0f; fb += lz / 32.0f;
        }
        fr /= 2.0f; fg /= 2.0f; fb /= 2.0f;
        if (lane == 0) {
            const long long w = sw.wg[pe];
            const short4 rq = sw.req[pe];
            if constexpr (DEBUG) {
                if (a.dbg_pre_blend) {
                    a.dbg_pre_blend[3 * w] = fr; a.dbg_pre_blend[3 * w + 1] = fg; a.dbg_pre_blend[3 * w + 2] = fb;
                }
            }
            uint32_t word = 0;                                 // store_probe :349-354 (deferred: Jacobi)
            if ((unsigned)rq.x < (unsigned)s.wx && (unsigned)rq.y < (unsigned)s.wy && (unsigned)rq.z < (unsigned)s.wz) {
                f3 old = unpack10(__ldg(s.field + (rq.x + s.wx * (rq.y + s.wy * rq.z))));
                old.x = gmin(gmax(old.x, 0.0f), 1.0f); old.y = gmin(gmax(old.y, 0.0f), 1.0f); old.z = gmin(gmax(old.z, 0.0f), 1.0f);
                const float t = a.reactiveness, omt = 1.0f - t;
                word = pack10(old.x * omt + fr * t, old.y * omt + fg * t, old.z * omt + fb * t);
            }
            a.staged[w] = word;
        }
        u.ms &= ~(1u << (16 + pe));                            // ring entry free again
    }
    if constexpr (DEBUG) {
#pragma unroll
        for (int k = 16; k >= 1; k >>= 1) {
            my_s += __shfl_xor_sync(FULL, my_s, k);
            my_h += __shfl_xor_sync(FULL, my_h, k);
        }
        if (lane == 0) { atomicAdd(a.dbg_totals, my_s); atomicAdd(a.dbg_totals + 1, my_h); }
    }
    __syncwarp();
    return u;
}

// ---- set up the next 32 rays (one half of the warp's next probe) into 32 free slots ---------------------
template <bool DEBUG>
__device__ __forceinline__ StreamState stream_setup(StreamWarp<DEBUG> &sw, const UpdateArgs &a, StreamState u) {
    constexpr int R = 64;
    const int lane = threadIdx.x & 31;
    const long long w = st_next_w(u, a);
    const int half = (u.ms >> 8) & 1;
    int cur_e = (u.ms >> 9) & 7;
    if (half == 0) {
        cur_e = __ffs(~(u.ms >> 16)) - 1;
        u.ms = (u.ms & ~(7u << 9)) | ((unsigned)cur_e << 9) | (1u << (16 + cur_e));
        const short4 rq = a.requests[w * (long long)a.size + a.shift];
        if (lane == 0) { sw.req[cur_e] = rq; sw.wg[cur_e] = w; sw.cnt[cur_e] = 0; }
        __syncwarp();
    }
    const short4 rq = sw.req[cur_e];
    const int ray = half * 32 + lane;
    const uint32_t gid_x = (uint32_t)(w * R + ray);
    const f3 fib = {__ldg(a.dir_table + 3 * ray), __ldg(a.dir_table + 3 * ray + 1), __ldg(a.dir_table + 3 * ray + 2)};
    f3 o, d, inv;
    setup_ray(rq.x, rq.y, rq.z, gid_x, ray, fib, a.time_seed, o, d, inv);
    const int free_n = st_free(u), rtail = (u.qs >> 8) & 255;
    const int slot = sw.fl[free_n - 32 + lane];
    sw.a[slot] = make_float4(o.x, o.y, o.z, 0.0f);
    sw.b[slot] = make_float4(d.x, d.y, d.z, inv.x);
    sw.c[slot] = make_float2(inv.y, inv.z);
    sw.pid[slot] = (unsigned short)((cur_e << 8) | ray);
    if constexpr (DEBUG) sw.ns[slot] = 0;
    sw.rq[(rtail + lane) & 127] = (unsigned char)slot;
    u.ms -= 32u;                                             // free_n -= 32
    u.qs = (u.qs & 0xffff00ffu) | (((unsigned)(rtail + 32) & 255u) << 8);
    u.ms ^= 1u << 8;                                         // half
    if (half == 1) u.k++;
    __syncwarp();
    return u;
}

template <bool DEBUG, bool FAST>
__device__ __forceinline__ void stream_warp_body(StreamWarp<DEBUG> &sw, const SceneView &s, const UpdateArgs &a,
                                                 const StreamArgs &sa) {
    constexpr int S = STREAM_S, P = STREAM_P;
    constexpr unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    const MarchConsts mc = {(float)s.wx * 16.0f, (float)s.wy * 16.0f, (float)s.wz * 16.0f};
    const uint3 wsbits = {__float_as_uint(mc.wsx), __float_as_uint(mc.wsy), __float_as_uint(mc.wsz)};
    StreamState u = {0u, (unsigned)S, 0u};

    for (int i = lane; i < S; i += 32) sw.fl[i] = (unsigned char)i;
    __syncwarp();

    // per-lane state: cur >= 0: slot of the ray being marched; -1: idle; <= -2: ray of slot -2 - cur just ended
    int cur = -1;
    PoolRay q = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}, 0.0f, 0, {0, 0, 0}};

    for (;;) {
        const unsigned idle = __ballot_sync(FULL, cur < 0);
        if (idle) {
            // lanes whose ray ended in the previous step hand their slot to FINISHED
            const unsigned finm = __ballot_sync(FULL, cur < -1);
            if (finm) {
                const int ftail = u.qs >> 24;
                if (cur < -1) { sw.fq[(ftail + __popc(finm & lt)) & 127] = (unsigned char)(-2 - cur); cur = -1; }
                u.qs += (unsigned)__popc(finm) << 24;
                __syncwarp();
            }
            // Shading and setup need the registers the marching state occupies: a lane's ray goes back to its
            // slot (only t changed) and is reloaded afterwards, so nothing of it stays live across the calls.
            const bool do_shade = st_finished(u) >= 32;
            const bool more = st_next_w(u, a) < a.wg_end;
            const bool ring_room = ((u.ms >> 8) & 1) || do_shade || (u.ms >> 16) != (1u << P) - 1u;
            const bool do_setup = st_ready(u) < 32 && st_free(u) + (do_shade ? 32 : 0) >= 32 && more && ring_room;
            if (do_shade || do_setup) {
                if (cur >= 0) {
                    sw.a[cur].w = q.t;
                    if constexpr (DEBUG) sw.ns[cur] = (unsigned short)q.ns;
                }
                if (do_shade) u = stream_shade<DEBUG, FAST>(sw, s, a, sa, u, 32);
                if (do_setup && st_free(u) >= 32 && (((u.ms >> 8) & 1) || (u.ms >> 16) != (1u << P) - 1u))
                    u = stream_setup<DEBUG>(sw, a, u);
                sw.load(cur >= 0 ? cur : 0, q);
            }
            const int avail = st_ready(u);
            if (avail > 0) {
                const int rank = __popc(idle & lt);
                if (cur < 0 && rank < avail) {
                    cur = sw.rq[((u.qs & 255) + rank) & 127];
                    sw.load(cur, q);
                }
                u.qs = (u.qs & 0xffffff00u) | (((u.qs & 255u) + (unsigned)min(__popc(idle), avail)) & 255u);
            }
            if (__ballot_sync(FULL, cur < 0) == FULL) {          // nothing to march
                if (st_ready(u) == 0) {
                    // READY is empty: shade what is pending (frees slots and ring entries); done when nothing is left
                    const int nf = st_finished(u);
                    if (nf > 0) { u = stream_shade<DEBUG, FAST>(sw, s, a, sa, u, min(32, nf)); continue; }
                    if (!(st_next_w(u, a) < a.wg_end)) break;
                }
                continue;
            }
        }
        if (cur >= 0) {
            bool done, hit;
            float t_end;
            if constexpr (FAST) {
                done = march_step_fast(s, wsbits, q.o, q.d, q.inv, q.cm, q.t, t_end, hit);
            } else {
                const int st = march_step(s, mc, q.o, q.d, q.inv, q.t);
                done = st != 0; hit = st == 1;
                t_end = q.t;
            }
            q.ns++;
            if (done) {
                sw.a[cur].w = hit ? q.t : -t_end;
                if constexpr (DEBUG) sw.ns[cur] = (unsigned short)q.ns;
                cur = -2 - cur;
            }
        }
    }
}

template <bool DEBUG>
__global__ void __launch_bounds__(STREAM_WARPS * 32, 9) radiance_update_stream_kernel(SceneView s, UpdateArgs a, StreamArgs sa) {
    __shared__ StreamWarp<DEBUG> warps[STREAM_WARPS];
    StreamWarp<DEBUG> &sw = warps[threadIdx.x >> 5];
    const bool general = !s.brick0_empty || (__ldg(s.flags) & 1u);
    if (general) stream_warp_body<DEBUG, false>(sw, s, a, sa);
    else stream_warp_body<DEBUG, true>(sw, s, a, sa);
}

}  // namespace lumcu
