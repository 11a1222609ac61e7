// kernels_on_host.cpp -- TEST INFRASTRUCTURE ONLY (see simt_emu.hpp).
// The radiance update as lumcu.cu's do_update() launches it -- trace kernel (pooled or one workgroup per CTA), then the
// commit kernel -- with the kernels of lum_b200/csrc/radiance_kernels.cuh themselves running on the host emulator.
#include <cstring>
#include "simt_emu.hpp"
#include "../../lum_b200/csrc/radiance_kernels.cuh"

using namespace lumcu;

#ifndef LUMCU_RPT
#define LUMCU_RPT 3
#endif

template <int R, int T>
static void launch_pool(const SceneView &s, const UpdateArgs &a, long long nwg) {
    constexpr int ppc = LUMCU_RPT * T / R;
    emu_launch((unsigned)((nwg + ppc - 1) / ppc), 1, T, [&] { radiance_update_pool_kernel<R, T, LUMCU_RPT, true>(s, a); });
}
template <int R>
static void launch_wg(const SceneView &s, const UpdateArgs &a, long long nwg) {
    // fewer CTAs than workgroups, as lumcu.cu caps the grid of this kernel: every CTA strides over ~3 workgroups
    emu_launch((unsigned)((nwg + 2) / 3), 1, R, [&] { radiance_update_wg_kernel<R, true>(s, a); });
}

// Trace workgroups [0, nreq/size) of shard il into staged[] (validate_world_kernel first, like ensure_validated()).
// variant: 0 = one workgroup per CTA, 1 = pooled.  Returns the workgroups traced, < 0 on bad arguments.
static long long trace_on_host(int wx, int wy, int wz, int capacity, int brick0_empty, int R, float reactiveness,
                               const int16_t *world, const uint8_t *bricks, const float *materials, const uint32_t *field_old,
                               const int16_t *requests, long long nreq, int time_seed, int size, int shift,
                               const float *table, const float *sun3, int variant, Interleave il, uint32_t *staged,
                               float *pre_blend, uint32_t *ray_samples, uint8_t *ray_hits, unsigned long long *totals,
                               const PeerWords *inbox = nullptr, long long slack = 0) {
    if (size < 1 || shift < 0 || shift >= size || il.n < 1 || il.rank < 0 || il.rank >= il.n || il.chunk < 0) return -1;
    std::vector<float4> mat4(256);
    for (int i = 0; i < 256; i++) mat4[i] = make_float4(materials[6 * i], materials[6 * i + 1], materials[6 * i + 2], materials[6 * i + 4]);
    const long long cells = (long long)wx * wy * wz;
    uint32_t flags = 0;
    emu_launch(2, 1, 64, [&] { validate_world_kernel(world, cells, capacity, &flags); });
    const SceneView s{world, bricks, materials, mat4.data(), field_old, wx, wy, wz, capacity, brick0_empty, &flags};
    // slack > 0: the host only knows an upper bound of the request count (lumcu_build_requests without a sync); the
    // launch is sized for the bound and the kernels cut the dispatch down to the device-side count
    const unsigned long long count = (unsigned long long)nreq;
    const long long nwg_all = (nreq + slack) / size;
    UpdateArgs a{};
    a.requests = reinterpret_cast<const short4 *>(requests);
    a.dir_table = table;
    a.staged = staged;
    if (inbox) a.inbox = *inbox;
    if (slack > 0) a.nreq_dev = &count;
    a.dbg_pre_blend = pre_blend; a.dbg_samples = ray_samples; a.dbg_hits = ray_hits; a.dbg_totals = totals;
    a.wg_begin = 0; a.wg_end = nwg_all;
    a.il_chunk = il.chunk; a.il_rank = il.rank; a.il_n = il.n;
    a.time_seed = time_seed; a.size = size; a.shift = shift;
    a.z_begin = -32768; a.z_end = 32768;
    a.reactiveness = reactiveness;
    a.sun_x = sun3[0]; a.sun_y = sun3[1]; a.sun_z = sun3[2];
    totals[0] = totals[1] = 0;
    // workgroups this launch traces (lumcu.cu: il_count)
    long long nwg = nwg_all;
    if (il.n > 1 && il.chunk == 0) {
        nwg = (nwg_all + il.n - 1) / il.n;         // contiguous shards: sized for the largest
    } else if (il.n > 1) {
        const long long full = nwg_all / il.chunk, rem = nwg_all % il.chunk;
        nwg = (full / il.n) * il.chunk + ((full % il.n) > il.rank ? il.chunk : 0) + ((full % il.n) == il.rank ? rem : 0);
    }
    if (nwg <= 0) return 0;
    if (variant == 1) {
        switch (R) {
            case 16: launch_pool<16, 128>(s, a, nwg); break;
            case 32: launch_pool<32, 128>(s, a, nwg); break;
            case 64: launch_pool<64, 128>(s, a, nwg); break;
            case 128: launch_pool<128, 128>(s, a, nwg); break;
            case 256: launch_pool<256, 256>(s, a, nwg); break;
            default: return -1;
        }
    } else {
        switch (R) {
            case 16: launch_wg<16>(s, a, nwg); break;
            case 32: launch_wg<32>(s, a, nwg); break;
            case 64: launch_wg<64>(s, a, nwg); break;
            case 128: launch_wg<128>(s, a, nwg); break;
            case 256: launch_wg<256>(s, a, nwg); break;
            default: return -1;
        }
    }
    return nwg;
}

constexpr int IL_CHUNK = 3072;     // lumcu.cu

// One update on one device, as lumcu_update launches it: trace, then commit_kernel.
extern "C" int kernels_on_host_update(int wx, int wy, int wz, int capacity, int brick0_empty, int R, float reactiveness,
                                      const int16_t *world, const uint8_t *bricks, const float *materials,
                                      const uint32_t *field_old, uint32_t *field_new,
                                      const int16_t *requests, long long nreq, int time_seed, int size, int shift,
                                      const float *table, const float *sun3, int variant, int il_rank, int il_n,
                                      float *pre_blend, uint32_t *ray_samples, uint8_t *ray_hits, unsigned long long *totals,
                                      long long slack) {
    const long long cells = (long long)wx * wy * wz, nwg_all = size > 0 ? (nreq + slack) / size : 0;
    const unsigned long long count = (unsigned long long)nreq;
    std::vector<uint32_t> staged((size_t)std::max<long long>(nwg_all, 1));
    const Interleave il{IL_CHUNK, il_rank, il_n};
    const long long nwg = trace_on_host(wx, wy, wz, capacity, brick0_empty, R, reactiveness, world, bricks, materials, field_old,
                                        requests, nreq, time_seed, size, shift, table, sun3, variant, il, staged.data(),
                                        pre_blend, ray_samples, ray_hits, totals, nullptr, slack);
    if (nwg < 0) return -1;
    for (long long i = 0; i < cells; i++) field_new[i] = field_old[i];
    if (nwg == 0) return 0;
    const int ct = 64;
    if (size == 1 && shift == 0 && il_n == 1) {
        // the plain full update commits through commit_vec4_kernel, as lumcu.cu's commit_local does (16-byte accesses:
        // the list goes into a 16-byte-aligned copy, a numpy slice need not be)
        std::vector<uint4> aligned((size_t)(nwg_all + 1) / 2 + 1);
        std::memcpy(aligned.data(), requests, (size_t)nwg_all * sizeof(short4));
        std::vector<uint4> field_al((size_t)(cells + 3) / 4);
        std::memcpy(field_al.data(), field_new, (size_t)cells * 4);
        const long long threads = (nwg_all + 3) / 4;
        emu_launch((unsigned)((threads + ct - 1) / ct), 1, ct, [&] {
            commit_vec4_kernel(reinterpret_cast<uint32_t *>(field_al.data()), staged.data(), reinterpret_cast<const short4 *>(aligned.data()),
                               nwg_all, wx, wy, wz, slack > 0 ? &count : nullptr);
        });
        std::memcpy(field_new, field_al.data(), (size_t)cells * 4);
        return 0;
    }
    emu_launch((unsigned)((nwg + ct - 1) / ct), 1, ct, [&] {
        commit_kernel(field_new, staged.data(), reinterpret_cast<const short4 *>(requests), 0, nwg_all, size, shift, wx, wy, wz, -32768, 32768, il,
                      slack > 0 ? &count : nullptr, ExchangeWait{});
    });
    return 0;
}

// The store side alone: staged[w] -> field at request w, through commit_vec4_kernel (vec != 0) or commit_kernel, for a
// hand-made list.  count_slack > 0: the launch is sized for nreq + count_slack and cut down to the device-side count.
extern "C" int kernels_on_host_commit(int wx, int wy, int wz, const int16_t *requests, long long nreq, const uint32_t *staged_in,
                                      uint32_t *field, int vec, long long count_slack) {
    const long long cells = (long long)wx * wy * wz, bound = nreq + count_slack;
    const unsigned long long count = (unsigned long long)nreq;
    std::vector<uint4> req_al((size_t)(bound + 1) / 2 + 1, make_uint4(0u, 0u, 0u, 0u));       // 16-byte aligned, as cudaMalloc gives
    std::memcpy(req_al.data(), requests, (size_t)nreq * sizeof(short4));
    std::vector<uint4> staged_al((size_t)(bound + 3) / 4 + 1, make_uint4(0xdeadbeefu, 0xdeadbeefu, 0xdeadbeefu, 0xdeadbeefu));
    std::memcpy(staged_al.data(), staged_in, (size_t)nreq * 4);
    std::vector<uint4> field_al((size_t)(cells + 3) / 4);
    std::memcpy(field_al.data(), field, (size_t)cells * 4);
    const int ct = 64;
    if (vec) {
        const long long threads = (bound + 3) / 4;
        emu_launch((unsigned)((threads + ct - 1) / ct), 1, ct, [&] {
            commit_vec4_kernel(reinterpret_cast<uint32_t *>(field_al.data()), reinterpret_cast<const uint32_t *>(staged_al.data()),
                               reinterpret_cast<const short4 *>(req_al.data()), bound, wx, wy, wz, count_slack > 0 ? &count : nullptr);
        });
    } else {
        emu_launch((unsigned)((bound + ct - 1) / ct), 1, ct, [&] {
            commit_kernel(reinterpret_cast<uint32_t *>(field_al.data()), reinterpret_cast<const uint32_t *>(staged_al.data()),
                          reinterpret_cast<const short4 *>(req_al.data()), 0, bound, 1, 0, wx, wy, wz, -32768, 32768,
                          Interleave{IL_CHUNK, 0, 1}, count_slack > 0 ? &count : nullptr, ExchangeWait{});
        });
    }
    std::memcpy(field, field_al.data(), (size_t)cells * 4);
    return 0;
}

// `frames` frames on `nranks` devices as the exchange runs them.  In epoch e every rank traces its shard reading its own
// replica of the field; arrive_kernel publishes e in every rank's flag array; then
//   push == 0 (lumcu_update_shard_exchange): the trace wrote half e & 1 of the rank's OWN staged[], and each rank's
//             gather_commit_kernel -- which waits for all flags -- reads every workgroup's texel from its owner's staged[]
//             and stores it into its own replica;
//   push == 1 (lumcu_update_slab_exchange's transfer): the trace stored the new texels into half e & 1 of EVERY rank's
//             inbox from its last phase, and each rank's commit_kernel scatters the whole dispatch from its own inbox.
// (Ranks run one after another here, so every flag is already there when a kernel looks; what is checked is the index
// arithmetic -- shard windows, owners, parity halves -- and that every replica ends up equal to the single-device field.)
// replicas: uint32 [nranks][cells], in = old field, out = field after the last frame.  il_chunk: 0 = contiguous shards,
// else a multiple of the pooled kernel's probes per CTA (lumcu.cu uses 3072; small here so that a small dispatch is
// really dealt out).
extern "C" int kernels_on_host_update_sharded(int wx, int wy, int wz, int capacity, int brick0_empty, int R, float reactiveness,
                                              const int16_t *world, const uint8_t *bricks, const float *materials,
                                              uint32_t *replicas, int nranks, int il_chunk,
                                              const int16_t *requests, long long nreq, int time_seed, int size, int shift,
                                              const float *table, const float *sun3, int variant, int frames, int push, long long slack) {
    if (nranks < 1 || nranks > MAX_PEERS) return -1;
    // slack > 0: every rank only knows an upper bound of the request count (a list built without a sync): launches and
    // shard chunks follow the bound, the kernels cut the dispatch down to the count they read from "device" memory
    const unsigned long long count = (unsigned long long)nreq;
    const long long cells = (long long)wx * wy * wz, nwg_all = size > 0 ? (nreq + slack) / size : 0;
    const size_t half = (size_t)std::max<long long>(nwg_all, 1);
    std::vector<std::vector<uint32_t>> staged((size_t)nranks, std::vector<uint32_t>(2 * half, 0xDEADBEEFu));
    std::vector<std::vector<unsigned>> flags((size_t)nranks, std::vector<unsigned>(MAX_PEERS, 0u));
    unsigned status = 0;
    unsigned long long tot[2];
    for (int f = 0; f < frames; f++) {
        const unsigned epoch = (unsigned)f + 1u;
        const int parity = (int)(epoch & 1u);
        PeerWords inbox{};
        inbox.n = nranks;
        for (int p = 0; p < nranks; p++) inbox.p[p] = staged[p].data() + parity * half;
        for (int r = 0; r < nranks; r++) {
            const long long traced = trace_on_host(wx, wy, wz, capacity, brick0_empty, R, reactiveness, world, bricks, materials,
                                                   replicas + (size_t)r * cells, requests, nreq, time_seed + f, size, shift, table, sun3,
                                                   variant, Interleave{il_chunk, r, nranks}, staged[r].data() + parity * half,
                                                   nullptr, nullptr, nullptr, tot, push ? &inbox : nullptr, slack);
            if (traced < 0) return -1;
            PeerFlags pf{};
            pf.n = nranks; pf.rank = r;
            for (int p = 0; p < nranks; p++) pf.f[p] = flags[p].data();
            emu_launch(1, 1, 32, [&] { arrive_kernel(pf, epoch); });
        }
        const int ct = 64;
        for (int r = 0; r < nranks; r++) {
            if (nwg_all <= 0) continue;
            const ExchangeWait xw{flags[r].data(), nranks, epoch, 1ll << 40, &status};
            if (push)
                emu_launch((unsigned)((nwg_all + ct - 1) / ct), 1, ct, [&] {
                    commit_kernel(replicas + (size_t)r * cells, staged[r].data() + parity * half, reinterpret_cast<const short4 *>(requests),
                                  0, nwg_all, size, shift, wx, wy, wz, -32768, 32768, Interleave{il_chunk, 0, 1}, slack > 0 ? &count : nullptr, xw);
                });
            else
                emu_launch(3, 1, ct, [&] {                          // a persistent grid: 3 CTAs stride over the dispatch
                    gather_commit_kernel(replicas + (size_t)r * cells, ShardSources{inbox, il_chunk, nranks},
                                         reinterpret_cast<const short4 *>(requests), 0, nwg_all, size, shift, wx, wy, wz,
                                         slack > 0 ? &count : nullptr, xw);
                });
        }
    }
    return status ? -2 : 0;
}

// The request-list builder as lumcu_build_requests launches it: build_fused_kernel (rule + look-back scan + ordered
// scatter in one pass; the emulator runs CTAs one after another, so a tile's predecessors have always published) and
// build_specials_kernel.  out: int16 [cap][4]; returns the request count, *nbuilt = the z-sorted part (without the
// special tail).  (The single-CTA specials kernel runs with 256 instead of 1024 threads here -- it is written for any
// multiple of 32 -- to keep the emulator at 256 OS threads.)
extern "C" long long kernels_on_host_build_requests(int wx, int wy, int wz, const int16_t *world, const int16_t *special,
                                                    int nspecial, int16_t *out, long long cap, long long *nbuilt) {
    const long long cells = (long long)wx * wy * wz;
    const long long ntiles = (cells + BUILD_TILE - 1) / BUILD_TILE;
    std::vector<uint8_t> flags((size_t)cells);
    std::vector<unsigned long long> state((size_t)ntiles + 1, 0ull);
    unsigned long long total[2] = {0, 0};      // [0] the list, [1] its rule-built prefix
    short4 *o = reinterpret_cast<short4 *>(out);
    emu_launch((unsigned)ntiles, 1, BUILD_THREADS, [&] {
        build_fused_kernel(world, wx, wy, wz, flags.data(), state.data(), reinterpret_cast<unsigned int *>(state.data() + ntiles), o, cap, total);
    });
    if (nspecial > 0)
        emu_launch(1, 1, 256, [&] {
            build_specials_kernel(flags.data(), wx, wy, wz, reinterpret_cast<const short4 *>(special), nspecial, o, total, cap);
        });
    *nbuilt = (long long)total[1];
    return (long long)total[0];
}

// shift_radiance as lumcu_shift launches it (tmp = copy of the field).
extern "C" int kernels_on_host_shift(int wx, int wy, int wz, uint32_t *field, int dx, int dy, int dz) {
    const long long cells = (long long)wx * wy * wz;
    if ((dx == 0 && dy == 0 && dz == 0) || abs(dx) >= wx || abs(dy) >= wy || abs(dz) >= wz) return 0;
    std::vector<uint32_t> tmp(field, field + cells);
    const int t = 256;
    emu_launch((unsigned)((cells + t - 1) / t), 1, t, [&] { shift_kernel(field, tmp.data(), wx, wy, wz, dx, dy, dz); });
    return 0;
}

// exec_copies as lumcu_copy_bricks launches it: one CTA of 256 threads per (src, dst) pair, 16 bytes per thread.
extern "C" int kernels_on_host_copy_bricks(uint8_t *bricks, int capacity, const int32_t *src, const int32_t *dst, int n) {
    std::vector<int2> pairs((size_t)std::max(n, 1));
    for (int i = 0; i < n; i++) { pairs[i].x = src[i]; pairs[i].y = dst[i]; }
    if (n > 0) emu_launch((unsigned)n, 1, 256, [&] { copy_bricks_kernel(bricks, pairs.data(), n, capacity); });
    return 0;
}

// lumcu_slab_range: first workgroup with z >= z_begin / z >= z_end in the z-sorted part of the request list.
extern "C" int kernels_on_host_slab_range(const int16_t *requests, long long nwg_sorted, int size, int shift,
                                          int z_begin, int z_end, long long *out2) {
    out2[0] = nwg_sorted; out2[1] = nwg_sorted;
    const int t = 256;
    if (nwg_sorted > 0)
        emu_launch((unsigned)((nwg_sorted + t - 1) / t), 1, t, [&] {
            slab_bound_kernel(reinterpret_cast<const short4 *>(requests), nwg_sorted, size, shift, z_begin, z_end, out2);
        });
    return 0;
}
