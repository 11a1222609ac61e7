// SHELVED EXPERIMENTS -- not part of the build.  Two alternatives to the pooled kernel that were built on top of
// the fast march step, passed the whole bit-exact parity suite (kernel variants 5 and 4) and were measured on
// B200 with the C4 world (pooled kernel 20.8 ms, workgroup-shaped kernel 21.0 ms):
//   * radiance_update_seq_kernel  (one warp per probe, each lane marches its own R/32 rays one after the other;
//     no ballot, no barrier, 2.5 KB of slots per warp): 21.6 ms.  17.9 lock-step iterations per 32 rays against
//     22.7 for the workgroup shape and 8.6 pooled -- private per-lane queues balance too little
//     (profiles/r01_ncu_warp_per_probe_kernel_rejected.txt).
//   * radiance_update_fold_kernel (workgroup shape; a warp marches until <= 16 rays live, then two warps'
//     survivors are packed into one, log2(warps) levels): 27.7 ms (27.8 with one probe per CTA, 31.3 with four).
//     Fewer instructions than the workgroup kernel, but every fold is a CTA barrier at which the faster warps
//     wait, and the kernel turns latency-bound.
// What the three results say together: with the fast step every shape ends at ~21 ms because every one of them is
// issue-bound at IPC 0.77-0.84; lane utilisation (32 % ... 75 %) trades one for one against bookkeeping
// instructions.  To use them again: paste into radiance_kernels.cuh before the commit kernels and launch
// <<<ceil(nwg / WPC), 32 * WPC>>> resp. <<<ceil(nwg / PPC), 64 * PPC>>>.
//
// ---- one warp per probe ---------------------------------------------------------------------------
// Each warp owns a whole probe: its R rays are set up convergently into the warp's shared-memory slots (R/32
// rounds), then every lane marches ITS OWN R/32 rays one after the other (slot lane, lane + 32, ...): taking the
// next ray needs no ballot, no prefix sum, no atomic, and nothing is shared between warps -- no barrier anywhere.
// A warp is busy for the longest per-lane SUM of ray lengths instead of the longest single ray per 32 (C4, R = 64:
// ~31 lock-step iterations per probe instead of 45 for the workgroup-shaped kernel), shading runs in R/32 fully
// convergent rounds whose xor butterflies ARE the subgroup sums of radiance.comp:386, and the slots cost 2.5 KB
// per warp, so the SM keeps most of its L1.
template <int R, bool DEBUG>
struct SeqWarp {
    float4 a[R];                          // o.xyz, t (after the march: t if hit, -t_pre if not)
    float4 b[R];                          // d.xyz, inv.x
    float2 c[R];                          // inv.yz
    unsigned short ns[DEBUG ? R : 1];
};

template <int R, bool DEBUG, bool VALID>
__device__ __forceinline__ void seq_warp_body(SeqWarp<R, DEBUG> &sw, const SceneView &s, const UpdateArgs &a, long long w) {
    constexpr int RPL = R / 32;           // rays per lane
    constexpr unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const MarchConsts mc = {(float)s.wx * 16.0f, (float)s.wy * 16.0f, (float)s.wz * 16.0f};
    const uint3 wsbits = {__float_as_uint(mc.wsx), __float_as_uint(mc.wsy), __float_as_uint(mc.wsz)};
    const short4 rq = a.requests[w * (long long)a.size + a.shift];

    // ---- setup ----
#pragma unroll
    for (int k = 0; k < RPL; k++) {
        const int ray = k * 32 + lane;
        const uint32_t gid_x = (uint32_t)(w * R + ray);
        const f3 fib = {__ldg(a.dir_table + 3 * ray), __ldg(a.dir_table + 3 * ray + 1), __ldg(a.dir_table + 3 * ray + 2)};
        f3 o, d, inv;
        setup_ray(rq.x, rq.y, rq.z, gid_x, ray, fib, a.time_seed, o, d, inv);
        sw.a[ray] = make_float4(o.x, o.y, o.z, 0.0f);
        sw.b[ray] = make_float4(d.x, d.y, d.z, inv.x);
        sw.c[ray] = make_float2(inv.y, inv.z);
    }
    __syncwarp();

    // ---- march: a lane's rays one after the other ----
    {
        int cur = lane;
        PoolRay q;
        auto load = [&](int r) {
            const float4 va = sw.a[r], vb = sw.b[r];
            const float2 vc = sw.c[r];
            q.o = {va.x, va.y, va.z}; q.d = {vb.x, vb.y, vb.z}; q.inv = {vb.w, vc.x, vc.y}; q.t = 0.0f; q.ns = 0;
            q.cm = {0x4B000000 + (q.d.x > 0.0f ? 16 : 0), 0x4B000000 + (q.d.y > 0.0f ? 16 : 0), 0x4B000000 + (q.d.z > 0.0f ? 16 : 0)};
        };
        load(cur);
        while (__any_sync(FULL, cur >= 0)) {
            if (cur >= 0) {
                bool done, hit;
                float t_end;
                if constexpr (VALID) {
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
                    cur += 32;
                    if (cur < R) load(cur); else cur = -1;
                }
            }
        }
    }
    __syncwarp();

    // ---- refinement + shading + reduction, one subgroup per round ----
    const f3 sun = {a.sun_x, a.sun_y, a.sun_z};
    float fr = 0.0f, fg = 0.0f, fb = 0.0f;
    unsigned long long my_s = 0, my_h = 0;
#pragma unroll 1
    for (int k = 0; k < RPL; k++) {
        const int ray = k * 32 + lane;
        const float4 va = sw.a[ray], vb = sw.b[ray];
        const float2 vc = sw.c[ray];
        const f3 o = {va.x, va.y, va.z}, d = {vb.x, vb.y, vb.z}, inv = {vb.w, vc.x, vc.y};
        bool hit = va.w > 0.0f;
        if constexpr (VALID) { if (!hit) hit = resolve_pending(s, mc, o, d, -va.w); }
        int nr;
        const f3 l = finish_ray<VALID>(s, o, d, inv, fabsf(va.w), hit, sun, nr);
        float lx = l.x, ly = l.y, lz = l.z;
        if constexpr (DEBUG) {
            const unsigned nsamp = (unsigned)sw.ns[ray] + (unsigned)nr;
            if (a.dbg_samples) a.dbg_samples[w * R + ray] = nsamp;
            if (a.dbg_hits) a.dbg_hits[w * R + ray] = hit ? 1 : 0;
            my_s += nsamp; my_h += hit ? 1 : 0;
        }
#pragma unroll
        for (int j = 16; j >= 1; j >>= 1) {
            lx += __shfl_xor_sync(FULL, lx, j);
            ly += __shfl_xor_sync(FULL, ly, j);
            lz += __shfl_xor_sync(FULL, lz, j);
        }
        fr += lx / 32.0f; fg += ly / 32.0f; fb += lz / 32.0f;          // :386,398,404-406
    }
    fr /= (float)RPL; fg /= (float)RPL; fb /= (float)RPL;              // :407
    if constexpr (DEBUG) {
#pragma unroll
        for (int j = 16; j >= 1; j >>= 1) { my_s += __shfl_xor_sync(FULL, my_s, j); my_h += __shfl_xor_sync(FULL, my_h, j); }
        if (lane == 0) { atomicAdd(a.dbg_totals, my_s); atomicAdd(a.dbg_totals + 1, my_h); }
    }
    if (lane == 0) {
        if constexpr (DEBUG) {
            if (a.dbg_pre_blend) {
                a.dbg_pre_blend[3 * w] = fr; a.dbg_pre_blend[3 * w + 1] = fg; a.dbg_pre_blend[3 * w + 2] = fb;
            }
        }
        uint32_t word = 0;                                             // store_probe :349-354 (deferred: Jacobi)
        if ((unsigned)rq.x < (unsigned)s.wx && (unsigned)rq.y < (unsigned)s.wy && (unsigned)rq.z < (unsigned)s.wz) {
            f3 old = unpack10(__ldg(s.field + (rq.x + s.wx * (rq.y + s.wy * rq.z))));
            old.x = gmin(gmax(old.x, 0.0f), 1.0f); old.y = gmin(gmax(old.y, 0.0f), 1.0f); old.z = gmin(gmax(old.z, 0.0f), 1.0f);
            const float t = a.reactiveness, omt = 1.0f - t;
            word = pack10(old.x * omt + fr * t, old.y * omt + fg * t, old.z * omt + fb * t);
        }
        a.staged[w] = word;
    }
}

template <int R, int WPC, bool DEBUG>
__global__ void __launch_bounds__(WPC * 32) radiance_update_seq_kernel(SceneView s, UpdateArgs a) {
    __shared__ SeqWarp<R, DEBUG> warps[WPC];
    const long long w = il_map(a.wg_begin, a.il_chunk, a.il_rank, a.il_n, (long long)blockIdx.x * WPC + (threadIdx.x >> 5));   // il_chunk % WPC == 0
    if (w >= a.wg_end) return;                                         // uniform per warp
    SeqWarp<R, DEBUG> &sw = warps[threadIdx.x >> 5];
    const bool general = !s.brick0_empty || (__ldg(s.flags) & 1u);
    if (general) seq_warp_body<R, DEBUG, false>(sw, s, a, w);
    else seq_warp_body<R, DEBUG, true>(sw, s, a, w);
}

// ---- folding kernel ------------------------------------------------------------------------------
// The workgroup's shape again (one thread sets up one ray, PPC probes per CTA), but the marching warps FOLD: a
// warp marches its 32 rays in lock-step only until at most FOLD_AT of them are still alive, then the survivors of
// two warps are packed into one (ray state lives in shared-memory slots indexed by the ray's home thread, so a
// fold moves 1-byte ids, not rays) and the march goes on with half the warps; log2(warps) levels, the last one
// runs to completion.  Ray lengths are 1..67 samples with a mean of 7 (C4): un-folded, a warp needs 22.7
// lock-step iterations for its 32 rays (32 % of the lanes busy); with one fold 14.5 per warp.  Compared with the
// pooled kernel there is no refill, no parking and 2.6 KB instead of 16 KB of shared memory per CTA, i.e. the
// whole L1 and up to 64 warps per SM.
constexpr int FOLD_AT = 16;

template <int NT, bool DEBUG>
struct FoldShared {
    float4 a[NT];                         // o.xyz, t (after the march: t if hit, -t_pre if not)
    float4 b[NT];                         // d.xyz, inv.x
    float2 c[NT];                         // inv.yz
    unsigned short ns[DEBUG ? NT : 1];    // marching samples (instrumented variant only)
    unsigned char list[NT / 32][FOLD_AT]; // per warp: home threads of the rays it hands on
    int cnt[NT / 32];
    float sub[NT / 32][3];                // subgroup means (reduction)
};

template <int R, int PPC, bool DEBUG, bool VALID>
__device__ __forceinline__ void fold_cta_body(FoldShared<R * PPC, DEBUG> &sm, const SceneView &s, const UpdateArgs &a) {
    constexpr int NT = R * PPC, NW = NT / 32;
    constexpr int NSUB = R / 32;          // R is a multiple of 32 here
    constexpr unsigned FULL = 0xffffffffu;
    static_assert(R % 32 == 0 && (NW & (NW - 1)) == 0 && NW <= 8, "fold kernel: whole warps, power of two");
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned lt = (1u << lane) - 1u;
    const long long w0 = il_map(a.wg_begin, a.il_chunk, a.il_rank, a.il_n, (long long)blockIdx.x * PPC);
    const long long left = a.wg_end - w0;
    if (left <= 0) return;
    const int nprobes = left < PPC ? (int)left : PPC;
    const int pr = tid / R, ray = tid % R;
    const bool mine = pr < nprobes;
    const MarchConsts mc = {(float)s.wx * 16.0f, (float)s.wy * 16.0f, (float)s.wz * 16.0f};
    const uint3 wsbits = {__float_as_uint(mc.wsx), __float_as_uint(mc.wsy), __float_as_uint(mc.wsz)};

    // ---- setup: every thread its own ray, into its home slot ----
    short4 rq = make_short4(0, 0, 0, 0);
    if (mine) {
        rq = a.requests[(w0 + pr) * (long long)a.size + a.shift];
        const uint32_t gid_x = (uint32_t)((w0 + pr) * R + ray);
        const f3 fib = {__ldg(a.dir_table + 3 * ray), __ldg(a.dir_table + 3 * ray + 1), __ldg(a.dir_table + 3 * ray + 2)};
        f3 o, d, inv;
        setup_ray(rq.x, rq.y, rq.z, gid_x, ray, fib, a.time_seed, o, d, inv);
        sm.a[tid] = make_float4(o.x, o.y, o.z, 0.0f);
        sm.b[tid] = make_float4(d.x, d.y, d.z, inv.x);
        sm.c[tid] = make_float2(inv.y, inv.z);
        if constexpr (DEBUG) sm.ns[tid] = 0;
    }
    __syncwarp();

    // ---- march with folding ----
    int cur = mine ? tid : -1;            // home thread of the ray this lane marches
    for (int nwarps = NW;; nwarps >>= 1) {
        if (warp < nwarps) {
            PoolRay q = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}, 0.0f, 0, {0, 0, 0}};
            if (cur >= 0) {
                const float4 va = sm.a[cur], vb = sm.b[cur];
                const float2 vc = sm.c[cur];
                q.o = {va.x, va.y, va.z}; q.d = {vb.x, vb.y, vb.z}; q.inv = {vb.w, vc.x, vc.y}; q.t = va.w;
                if constexpr (DEBUG) q.ns = sm.ns[cur];
                q.cm = {0x4B000000 + (q.d.x > 0.0f ? 16 : 0), 0x4B000000 + (q.d.y > 0.0f ? 16 : 0), 0x4B000000 + (q.d.z > 0.0f ? 16 : 0)};
            }
            const int stop_at = nwarps > 1 ? FOLD_AT : 0;
            while (__popc(__ballot_sync(FULL, cur >= 0)) > stop_at) {
                if (cur >= 0) {
                    bool done, hit;
                    float t_end;
                    if constexpr (VALID) {
                        done = march_step_fast(s, wsbits, q.o, q.d, q.inv, q.cm, q.t, t_end, hit);
                    } else {
                        const int st = march_step(s, mc, q.o, q.d, q.inv, q.t);
                        done = st != 0; hit = st == 1;
                        t_end = q.t;
                    }
                    q.ns++;
                    if (done) {
                        sm.a[cur].w = hit ? q.t : -t_end;
                        if constexpr (DEBUG) sm.ns[cur] = (unsigned short)q.ns;
                        cur = -1;
                    }
                }
            }
            if (nwarps > 1) {             // hand the survivors on
                const unsigned alive = __ballot_sync(FULL, cur >= 0);
                if (cur >= 0) {
                    sm.list[warp][__popc(alive & lt)] = (unsigned char)cur;
                    sm.a[cur].w = q.t;
                    if constexpr (DEBUG) sm.ns[cur] = (unsigned short)q.ns;
                }
                if (lane == 0) sm.cnt[warp] = __popc(alive);
            }
        }
        if (nwarps == 1) break;
        __syncthreads();
        cur = -1;
        if (warp < (nwarps >> 1)) {       // lanes 0-15 take warp 2w's survivors, lanes 16-31 warp 2w+1's
            const int src = 2 * warp + (lane >> 4), j = lane & 15;
            if (j < sm.cnt[src]) cur = sm.list[src][j];
        }
        __syncthreads();                  // lists are rewritten at the next level
    }
    __syncthreads();

    // ---- refinement + shading: every thread its own ray again ----
    float lx = 0.0f, ly = 0.0f, lz = 0.0f;
    if (mine) {
        const float4 va = sm.a[tid], vb = sm.b[tid];
        const float2 vc = sm.c[tid];
        const f3 o = {va.x, va.y, va.z}, d = {vb.x, vb.y, vb.z}, inv = {vb.w, vc.x, vc.y};
        bool hit = va.w > 0.0f;
        if constexpr (VALID) { if (!hit) hit = resolve_pending(s, mc, o, d, -va.w); }
        int nr;
        const f3 sun = {a.sun_x, a.sun_y, a.sun_z};
        const f3 l = finish_ray<VALID>(s, o, d, inv, fabsf(va.w), hit, sun, nr);
        lx = l.x; ly = l.y; lz = l.z;
        if constexpr (DEBUG) {
            const unsigned nsamp = (unsigned)sm.ns[tid] + (unsigned)nr;
            if (a.dbg_samples) a.dbg_samples[(w0 + pr) * R + ray] = nsamp;
            if (a.dbg_hits) a.dbg_hits[(w0 + pr) * R + ray] = hit ? 1 : 0;
            unsigned long long sv = nsamp, hv = hit ? 1 : 0;
#pragma unroll
            for (int k = 16; k >= 1; k >>= 1) { sv += __shfl_xor_sync(FULL, sv, k); hv += __shfl_xor_sync(FULL, hv, k); }
            if (lane == 0) { atomicAdd(a.dbg_totals, sv); atomicAdd(a.dbg_totals + 1, hv); }
        }
    }
    // ---- reduction (radiance.comp:386-407) + store_probe ----
#pragma unroll
    for (int k = 16; k >= 1; k >>= 1) {
        lx += __shfl_xor_sync(FULL, lx, k);
        ly += __shfl_xor_sync(FULL, ly, k);
        lz += __shfl_xor_sync(FULL, lz, k);
    }
    if (lane == 0) { sm.sub[warp][0] = lx / 32.0f; sm.sub[warp][1] = ly / 32.0f; sm.sub[warp][2] = lz / 32.0f; }
    __syncthreads();
    if (tid < nprobes) {
        float fr = 0.0f, fg = 0.0f, fb = 0.0f;
#pragma unroll
        for (int g = 0; g < NSUB; g++) { fr += sm.sub[tid * NSUB + g][0]; fg += sm.sub[tid * NSUB + g][1]; fb += sm.sub[tid * NSUB + g][2]; }
        fr /= (float)NSUB; fg /= (float)NSUB; fb /= (float)NSUB;
        const long long w = w0 + tid;
        const short4 q = a.requests[w * (long long)a.size + a.shift];
        if constexpr (DEBUG) {
            if (a.dbg_pre_blend) {
                a.dbg_pre_blend[3 * w] = fr; a.dbg_pre_blend[3 * w + 1] = fg; a.dbg_pre_blend[3 * w + 2] = fb;
            }
        }
        uint32_t word = 0;
        if ((unsigned)q.x < (unsigned)s.wx && (unsigned)q.y < (unsigned)s.wy && (unsigned)q.z < (unsigned)s.wz) {
            f3 old = unpack10(__ldg(s.field + (q.x + s.wx * (q.y + s.wy * q.z))));
            old.x = gmin(gmax(old.x, 0.0f), 1.0f); old.y = gmin(gmax(old.y, 0.0f), 1.0f); old.z = gmin(gmax(old.z, 0.0f), 1.0f);
            const float t = a.reactiveness, omt = 1.0f - t;
            word = pack10(old.x * omt + fr * t, old.y * omt + fg * t, old.z * omt + fb * t);
        }
        a.staged[w] = word;
    }
}

template <int R, int PPC, bool DEBUG>
__global__ void __launch_bounds__(R * PPC) radiance_update_fold_kernel(SceneView s, UpdateArgs a) {
    __shared__ FoldShared<R * PPC, DEBUG> sm;
    const bool general = !s.brick0_empty || (__ldg(s.flags) & 1u);
    if (general) fold_cta_body<R, PPC, DEBUG, false>(sm, s, a);
    else fold_cta_body<R, PPC, DEBUG, true>(sm, s, a);
}

