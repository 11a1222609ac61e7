// lumcu.cu -- C-ABI of the B200 radiance stage (see include/lumcu.h).
//
// Host side of the stage: device mirror of Lum's world / block palette / material palette,
// the request buffer, the radiance field, and the launch logic that replaces
// LumInternalRenderer::update_radiance()'s Vulkan half (internal_render.cpp:1170-1213).
// There is no CPU path: every entry point needs a CUDA device.
#include "../../include/lumcu.h"
#include "radiance_kernels.cuh"

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

using namespace lumcu;

namespace {

thread_local std::string g_err;

int fail(int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CU(call)                                                                               \
    do {                                                                                       \
        cudaError_t e_ = (call);                                                               \
        if (e_ != cudaSuccess)                                                                 \
            return fail(LUMCU_E_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

// Pinned staging ring: the reference's stagingWorld / stagingRadianceUpdates ring<Buffer>
// (internal_render.cpp:198-207).  Two slots so the host can fill one while the other drains.
struct Staging {
    void *host[2] = {nullptr, nullptr};
    cudaEvent_t done[2] = {nullptr, nullptr};
    size_t cap[2] = {0, 0};
    int next = 0;
    int acquire(size_t bytes, void **out, int *slot) {
        int s = next;
        next ^= 1;
        if (done[s]) CU(cudaEventSynchronize(done[s]));
        else CU(cudaEventCreateWithFlags(&done[s], cudaEventDisableTiming));
        if (cap[s] < bytes) {
            if (host[s]) CU(cudaFreeHost(host[s]));
            host[s] = nullptr; cap[s] = 0;
            CU(cudaMallocHost(&host[s], bytes));
            cap[s] = bytes;
        }
        *out = host[s];
        *slot = s;
        return 0;
    }
    void release() {
        for (int s = 0; s < 2; s++) {
            if (host[s]) cudaFreeHost(host[s]);
            if (done[s]) cudaEventDestroy(done[s]);
            host[s] = nullptr; done[s] = nullptr; cap[s] = 0;
        }
    }
};

}  // namespace

// Interleaved sharding (UpdateArgs::il_*): chunks of IL_CHUNK workgroups, halved for smaller dispatches until every rank
// gets at least 64 of them (a rank's share can be off by one chunk: 65 K probes dealt to 8 ranks in chunks of 3,072 are
// two chunks for some ranks and three for others), never below IL_CHUNK_MIN -- every size is a multiple of every
// pooled kernel's probes per CTA.
constexpr int IL_CHUNK = 3072;
constexpr int IL_CHUNK_MIN = 48;

struct LumcuCtx {
    LumcuDesc desc{};
    int wx = 0, wy = 0, wz = 0, R = 64;
    long long cells = 0;
    long long max_requests = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;    // uploads from caller-pinned memory (no staging copy)
    cudaStream_t dl_stream = nullptr;      // lumcu_download_field_async: D2H beside the next frame's trace and upload
    cudaEvent_t ev_dl_ready = nullptr, ev_dl_done = nullptr;
    bool dl_pending = false;               // a D2H of the field is in flight on dl_stream: the next writer of the field waits for it
    cudaEvent_t ev_compute = nullptr, ev_copy = nullptr;
    cudaEvent_t ev_begin = nullptr, ev_end = nullptr, ev_k0 = nullptr, ev_k1 = nullptr;
    cudaEvent_t fences[2 * LUMCU_FENCES] = {};  // lumcu_record_fence / lumcu_wait_fence: [2*slot] stream, [2*slot+1] dl_stream
    bool timed = false, profiling = false, ktimed = false;

    int16_t *d_world = nullptr;            // = d_world_buf[world_cur]: the world the next launches read
    int16_t *d_world_buf[2] = {nullptr, nullptr};   // two buffers: lumcu_upload_world_async fills the other one while
    int world_cur = 0;                              // the frame in flight still reads this one, then flips
    cudaEvent_t ev_world_retired[2] = {nullptr, nullptr};   // "everything queued before the flip away from buffer b has run"
    bool world_retired[2] = {false, false};
    uint8_t *d_bricks = nullptr;
    float *d_materials = nullptr;
    float4 *d_mat4 = nullptr;              // {r,g,b,emmit} per id, repacked at upload (the shading phase's one load)
    uint32_t *d_field = nullptr;
    uint32_t *d_field_tmp = nullptr;       // shift scratch (lazy)
    short4 *d_requests = nullptr;          // = d_requests_buf[req_cur]: the list the next launches read
    short4 *d_requests_buf[2] = {nullptr, nullptr};     // two lists: a list built ahead goes into the other one while the
    unsigned long long *d_total_buf[2] = {nullptr, nullptr};   // frame in flight still reads this one (with its own count)
    int req_cur = 0;
    cudaEvent_t ev_list_retired[2] = {nullptr, nullptr};       // "everything queued before the flip away from list b has run"
    bool list_retired[2] = {false, false};
    cudaStream_t prep_stream = nullptr;    // high priority: next frame's request build + world validation, beside this frame's trace
    cudaEvent_t ev_prep = nullptr, ev_world_ready = nullptr, ev_scratch = nullptr;
    bool world_ready_pending = false;      // ev_world_ready marks an upload the prep stream has not waited for yet
    bool scratch_on_stream = false;        // the builder's scratch was last used on `stream` (ev_scratch), not on prep_stream

    uint32_t *d_staged = nullptr;
    float *d_table = nullptr;
    std::vector<float> h_table;
    float sun[3] = {0, 0, 0};

    long long nreq = 0;          // current request count -- an upper bound while count_on_device
    long long nsorted = 0;       // leading requests known to be z-sorted
    long long shard_basis = 0;      // what the interleave chunk size is derived from: the request count as lumcu_set_requests
                                    // gave it, or a built list's upper bound -- NOT the count learnt later, so that ranks
                                    // agree on the chunk size whether or not one of them has asked for its count
    long long count_hint = -1;      // the last request count the host has seen (a hint for kernel choice while count_on_device)
    unsigned long long *h_count = nullptr;   // pinned: every build without a sync also queues a 16-byte copy of its count here
    cudaEvent_t ev_count = nullptr;
    bool count_copy_pending = false;
    bool count_on_device = false;   // the list was built without a host sync: the true count is *d_total (+ d_nbuilt),
                                    // the kernels cut their dispatch down to it, and the host learns it when asked
    bool have_world = false, have_bricks = false, have_materials = false;
    int sm_count = 148;          // cudaDevAttrMultiProcessorCount (grid sizing of the grid-stride kernels)
    int pool_threads = 128;      // tuning knobs (env LUMCU_POOL_THREADS / LUMCU_CARVEOUT, read at create)
    int carveout = -1;
    bool vec_commit = true;      // env LUMCU_VEC_COMMIT=0: the element-wise commit kernel for the plain update too (A/B runs)
    int shard_chunk = IL_CHUNK;  // lumcu_set_shard_mode: IL_CHUNK = interleaved chunks (default), 0 = contiguous shards
    int kernel_variant = 2;      // 0 = one workgroup per CTA (radiance.comp's own shape), 1 = pooled, 2 = auto
    bool brick0_empty = true;    // device bricks start zeroed; tracked by lumcu_upload_bricks
    uint32_t *d_scene_flags = nullptr;     // [2], one word per world buffer; bit 0: world ids outside the palette (validate_world_kernel)
    bool world_dirty = true;     // the world changed since the last validation
    bool world_shared = false;   // its device address was handed out: validate before every update

    // builder scratch (lazy)
    uint8_t *d_flags = nullptr;
    unsigned long long *d_tile_state = nullptr;   // build_fused_kernel's look-back states + its ticket counter
    unsigned long long *d_total = nullptr;      // [0] = requests in the list, [1] = of which built by the 3x3x3 rule (z-sorted)
    short4 *d_special = nullptr;
    long long special_cap = 0;
    long long *d_bounds = nullptr;
    uint64_t req_version = 1, slab_key_version = 0;
    int slab_key[4] = {0, 0, 0, 0};
    long long slab_val[2] = {0, 0};

    // debug
    int debug_flags = 0;
    float *d_pre_blend = nullptr;
    uint32_t *d_ray_samples = nullptr;
    uint8_t *d_ray_hits = nullptr;
    unsigned long long *d_totals = nullptr;
    long long dbg_nwg = 0;       // workgroups of the last instrumented update
    long long dbg_wg_begin = 0;
    long long dbg_traced = 0;    // workgroups the last instrumented update traced (totals[2])

    // models (map.comp inputs) and blockify copy pairs
    struct Model { uint8_t *d_vox = nullptr; int sx = 0, sy = 0, sz = 0; };
    std::vector<Model> models;
    int2 *d_pairs = nullptr;
    long long pairs_cap = 0;

    // peer exchange (lumcu_peers_open)
    bool have_peers = false;
    int peer_rank = 0, peer_n = 1;
    PeerWords peer_inbox[2]{};             // [parity]: that half of every rank's staged[] (own included)
    PeerFlags peer_flags{};
    void *peer_opened[2 * MAX_PEERS] = {};
    int n_peer_opened = 0;
    unsigned *d_sync_flags = nullptr;      // [MAX_PEERS] epoch each rank last arrived at (written by the peers)
    unsigned *h_status = nullptr;          // pinned + mapped: barrier timeout flag
    unsigned *d_status = nullptr;
    unsigned epoch = 0;

    // glossy pass scratch for the host-buffer entry point (lazy)
    uchar4 *d_g_matnorm = nullptr; float *d_g_depth = nullptr; float4 *d_g_frame = nullptr; uint8_t *d_g_cov = nullptr;
    long long g_cap = 0;

    Staging stage_world, stage_req, stage_misc, stage_pairs;
    uint64_t launches = 0;
};

namespace {

int use(LumcuCtx *c) {
    if (!c) return fail(LUMCU_E_ARG, "null context");
    CU(cudaSetDevice(c->desc.device));
    return 0;
}

void host_direction_table(int n, std::vector<float> &out) {
    // sphericalFibonacci (radiance.comp:293-314) in f32, one rounding per operation
    out.resize((size_t)n * 3);
    const float PI = 3.1415926535897932384626433832795f;
    const float PHI = sqrtf(5.0f) * 0.5f + 0.5f;
    for (int k = 0; k < n; k++) {
        float i = (float)k, fn = (float)n;
        float ab = i * (PHI - 1.0f);
        float mf = ab - floorf(ab);
        float phi = 2.0f * PI * mf;
        float ct = 1.0f - (2.0f * i + 1.0f) * (1.0f / fn);
        float q = 1.0f - ct * ct;
        q = q < 0.0f ? 0.0f : q;
        q = 1.0f < q ? 1.0f : q;
        float st = sqrtf(q);
        out[3 * k + 0] = cosf(phi) * st;
        out[3 * k + 1] = sinf(phi) * st;
        out[3 * k + 2] = ct;
    }
}

// production path: pooled kernel (no z filter); one CTA per RPT*threads rays
#ifndef LUMCU_RPT
#define LUMCU_RPT 3           // rays per thread of the reference-shaped (R = 64) production kernel: 384 rays = 6 probes
                              // per CTA, 16 KB of ray slots (B200, C4 world: 22.16 ms; RPT 4: 23.16, RPT 2: 24.15)
#endif
template <int R, int T, int RPT, bool DEBUG>
int launch_pool_one(LumcuCtx *c, const SceneView &sv, const UpdateArgs &a, long long nwg) {
    constexpr int ppc = RPT * T / R;
    static_assert(ppc >= 1 && IL_CHUNK_MIN % ppc == 0 && IL_CHUNK % IL_CHUNK_MIN == 0, "an interleaved chunk must hold whole CTA batches (il_map)");
    auto k = radiance_update_pool_kernel<R, T, RPT, DEBUG>;
    if (c->carveout >= 0) cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, c->carveout);
    dim3 grid((unsigned)((nwg + ppc - 1) / ppc));
    k<<<grid, T, 0, c->stream>>>(sv, a);
    c->launches++;
    CU(cudaGetLastError());
    return 0;
}

template <bool DEBUG>
int launch_update_pool(LumcuCtx *c, const SceneView &sv, const UpdateArgs &a, long long nwg) {
    // 3 ray slots per thread everywhere (what the R = 64 sweep found best: the slots compete with L1 for the SM's
    // 256 KB); a batch must hold whole probes, so R = 256 takes 256 threads (3 probes per CTA)
    switch (c->R) {
        case 16: return launch_pool_one<16, 128, LUMCU_RPT, DEBUG>(c, sv, a, nwg);
        case 32: return launch_pool_one<32, 128, LUMCU_RPT, DEBUG>(c, sv, a, nwg);
        case 64: return c->pool_threads == 128 ? launch_pool_one<64, 128, LUMCU_RPT, DEBUG>(c, sv, a, nwg)
                                               : launch_pool_one<64, 256, 4, DEBUG>(c, sv, a, nwg);
        case 128: return launch_pool_one<128, 128, LUMCU_RPT, DEBUG>(c, sv, a, nwg);
        case 256: return launch_pool_one<256, 256, LUMCU_RPT, DEBUG>(c, sv, a, nwg);
        default: return fail(LUMCU_E_ARG, "rays_per_probe %d unsupported", c->R);
    }
}

template <bool DEBUG>
int launch_update_R(LumcuCtx *c, const SceneView &sv, const UpdateArgs &a, long long nwg) {
    // one CTA per workgroup up to a full machine of resident CTAs (32 per SM), striding beyond: the grid of a small
    // dispatch does not depend on the request count, which may still be on the device (and lets a CUDA graph replay it)
    dim3 grid((unsigned)std::min<long long>(nwg, (long long)c->sm_count * 32));
    switch (c->R) {
        case 16: radiance_update_wg_kernel<16, DEBUG><<<grid, 16, 0, c->stream>>>(sv, a); break;
        case 32: radiance_update_wg_kernel<32, DEBUG><<<grid, 32, 0, c->stream>>>(sv, a); break;
        case 64: radiance_update_wg_kernel<64, DEBUG><<<grid, 64, 0, c->stream>>>(sv, a); break;
        case 128: radiance_update_wg_kernel<128, DEBUG><<<grid, 128, 0, c->stream>>>(sv, a); break;
        case 256: radiance_update_wg_kernel<256, DEBUG><<<grid, 256, 0, c->stream>>>(sv, a); break;
        default: return fail(LUMCU_E_ARG, "rays_per_probe %d unsupported", c->R);
    }
    c->launches++;
    CU(cudaGetLastError());
    return 0;
}

int ensure_debug(LumcuCtx *c) {
    if (!c->d_totals) {
        CU(cudaMalloc(&c->d_totals, 2 * sizeof(unsigned long long)));
    }
    if ((c->debug_flags & LUMCU_DEBUG_PRE_BLEND) && !c->d_pre_blend)
        CU(cudaMalloc(&c->d_pre_blend, (size_t)c->max_requests * 3 * sizeof(float)));
    if ((c->debug_flags & LUMCU_DEBUG_RAYS) && !c->d_ray_samples) {
        CU(cudaMalloc(&c->d_ray_samples, (size_t)c->max_requests * c->R * sizeof(uint32_t)));
        CU(cudaMalloc(&c->d_ray_hits, (size_t)c->max_requests * c->R));
    }
    return 0;
}

// Device-side check that every world id is inside the palette (what march_step_fast relies on); no host sync.
// One flag word per world buffer, so that the check of the next frame's world can run while this frame still traces.
int launch_validate(LumcuCtx *c, cudaStream_t st) {
    uint32_t *flag = c->d_scene_flags + c->world_cur;
    CU(cudaMemsetAsync(flag, 0, sizeof(uint32_t), st));
    const int vt = 256;
    const long long vb = std::min<long long>((c->cells + vt - 1) / vt, (long long)c->sm_count * 8);
    validate_world_kernel<<<(unsigned)vb, vt, 0, st>>>(c->d_world, c->cells, c->desc.palette_capacity, flag);
    c->launches++;
    CU(cudaGetLastError());
    return 0;
}
int ensure_validated(LumcuCtx *c) {
    if (c->world_dirty || c->world_shared) {
        if (int rc = launch_validate(c, c->stream)) return rc;
        c->world_dirty = false;
    }
    return 0;
}

SceneView scene_view(LumcuCtx *c) {
    return SceneView{c->d_world, c->d_bricks, c->d_materials, c->d_mat4, c->d_field, c->wx, c->wy, c->wz, c->desc.palette_capacity,
                     c->brick0_empty ? 1 : 0, c->d_scene_flags + c->world_cur
    };
}

// chunk size of an interleaved deal of `total` workgroups to n ranks (`total` = the host's bound: the same on every rank)
int interleave_chunk(long long total, int n) {
    int ch = IL_CHUNK;
    while (ch > IL_CHUNK_MIN && (long long)ch * n * 64 > total) ch /= 2;
    return ch;
}
// the shard of a dispatch of `total` workgroups under the context's shard mode
Interleave shard_of(const LumcuCtx *c, long long total, int rank, int n) {
    return Interleave{c->shard_chunk == 0 ? 0 : interleave_chunk(total, n), rank, n};
}

// number of workgroups of [0, total) that shard `il` owns
long long il_count(long long total, const Interleave &il) {
    if (il.n <= 1) return total;
    if (il.chunk == 0) return (total + il.n - 1) / il.n;      // contiguous: the largest shard (also covers a device-side count below the bound)
    const long long full = total / il.chunk, rem = total % il.chunk;
    return (full / il.n) * il.chunk + ((full % il.n) > il.rank ? il.chunk : 0) + ((full % il.n) == il.rank ? rem : 0);
}

// The request count as the host needs it for launch sizes: exact, or (count_on_device) an upper bound.
const unsigned long long *count_ptr(const LumcuCtx *c) { return c->count_on_device ? c->d_total : nullptr; }

// Has the count of an earlier sync-less build arrived in pinned memory?  Never waits.
void poll_count(LumcuCtx *c) {
    if (c->count_copy_pending && cudaEventQuery(c->ev_count) == cudaSuccess) {
        c->count_hint = (long long)c->h_count[0];
        c->count_copy_pending = false;
    }
}

// Learn the device-side count (one 16-byte readback + sync); every entry point that needs exact numbers calls this.
int resolve_count(LumcuCtx *c) {
    if (!c->count_on_device) return 0;
    unsigned long long t[2] = {0, 0};
    CU(cudaMemcpyAsync(t, c->d_total, sizeof t, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    c->count_on_device = false;
    if ((long long)t[0] > c->max_requests) {
        c->nreq = 0; c->nsorted = 0;
        return fail(LUMCU_E_ARG, "request list (%llu) exceeds max_requests %lld", t[0], c->max_requests);
    }
    c->nreq = (long long)t[0];
    c->nsorted = (long long)t[1];
    c->count_hint = c->nreq;
    return 0;
}

// A lumcu_download_field_async may still be reading the field on the download stream: whoever writes the field next
// (a commit, a shift, an upload) is queued behind it.
int before_field_write(LumcuCtx *c) {
    if (c->dl_pending) {
        CU(cudaStreamWaitEvent(c->stream, c->ev_dl_done, 0));
        c->dl_pending = false;
    }
    return 0;
}

uint32_t *staged_half(LumcuCtx *c, int parity) { return c->d_staged + (size_t)parity * (size_t)c->max_requests; }

// xw.flags != null: a multi-GPU frame -- the staged half is this rank's inbox and the kernel waits for every rank's arrival
int commit_local(LumcuCtx *c, int size, int shift, long long wg_begin, long long wg_end, int z_begin, int z_end,
                 Interleave il = Interleave{IL_CHUNK, 0, 1}, int parity = 0, ExchangeWait xw = ExchangeWait{}) {
    const long long nwg = il_count(wg_end - wg_begin, il);
    if (nwg <= 0) return 0;
    if (int rc = before_field_write(c)) return rc;
    const int ct = 256;
    if (c->vec_commit && il.n <= 1 && !xw.flags && parity == 0 && size == 1 && shift == 0 && wg_begin == 0 && z_begin <= -32768 && z_end >= 32768) {
        // the plain full update: four workgroups per thread, 16-byte stores where four requests are four adjacent texels
        const long long threads = (wg_end + 3) / 4;
        commit_vec4_kernel<<<(unsigned)((threads + ct - 1) / ct), ct, 0, c->stream>>>(c->d_field, staged_half(c, parity), c->d_requests, wg_end,
                                                                                   c->wx, c->wy, c->wz, count_ptr(c));
        c->launches++;
        CU(cudaGetLastError());
        return 0;
    }
    commit_kernel<<<(unsigned)((nwg + ct - 1) / ct), ct, 0, c->stream>>>(
        c->d_field, staged_half(c, parity), c->d_requests, wg_begin, wg_end, size, shift, c->wx, c->wy, c->wz, z_begin, z_end, il,
        count_ptr(c), xw);
    c->launches++;
    CU(cudaGetLastError());
    return 0;
}

// second_pass: this launch continues a frame another do_update() started (the special tail of a slab update): the
// instrumented totals keep accumulating and the debug range stays the first pass's.
int do_update(LumcuCtx *c, int time_seed, int size, int shift, long long wg_begin, long long wg_end,
              int z_begin, int z_end, bool commit = true, Interleave il = Interleave{IL_CHUNK, 0, 1}, bool second_pass = false,
              const PeerWords *inbox = nullptr, int parity = 0) {
    if (!c->have_world || !c->have_bricks || !c->have_materials)
        return fail(LUMCU_E_STATE, "lumcu_update before world/bricks/materials were uploaded");
    if (size < 1) return fail(LUMCU_E_ARG, "size must be >= 1");
    if (shift < 0 || shift >= size) return fail(LUMCU_E_ARG, "shift must be in [0,size)");
    const long long nwg_all = c->nreq / size;
    if (wg_begin < 0 || wg_end > nwg_all || wg_begin > wg_end)
        return fail(LUMCU_E_ARG, "workgroup range [%lld,%lld) outside [0,%lld)", wg_begin, wg_end, nwg_all);
    const long long nwg = il_count(wg_end - wg_begin, il);     // workgroups this launch traces
    if (nwg > 0) {
        if (int rc = ensure_validated(c)) return rc;
        const SceneView sv = scene_view(c);
        UpdateArgs a{};
        a.requests = c->d_requests;
        a.dir_table = c->d_table;
        a.staged = staged_half(c, parity);
        if (inbox) a.inbox = *inbox;
        a.nreq_dev = count_ptr(c);
        a.wg_begin = wg_begin; a.wg_end = wg_end;
        a.il_chunk = il.chunk; a.il_rank = il.rank; a.il_n = il.n;
        a.time_seed = time_seed; a.size = size; a.shift = shift;
        a.z_begin = z_begin; a.z_end = z_end;
        a.reactiveness = c->desc.reactiveness;
        a.sun_x = c->sun[0]; a.sun_y = c->sun[1]; a.sun_z = c->sun[2];
        // the z filter (unsorted special tail of a slab update) only exists in the per-workgroup kernel
        // and small dispatches (demo-sized worlds) are latency-bound: they want one workgroup per CTA
        const bool unfiltered = z_begin <= -32768 && z_end >= 32768;
        // Which kernel (variant 2 = choose): the pooled kernel once the dispatch gives it most of a wave of CTAs, i.e. from
        // ~6 K workgroups on (B200, tools/kernel_crossover.py: 16 K probes -- pooled 0.048 / 0.098 / 0.31 ms against
        // 0.064 / 0.115 / 0.36 at 16 / 64 / 256 rays per probe; the demo world's 2,689 probes -- 0.043 / 0.047 / 0.072
        // against 0.028 / 0.033 / 0.062).  While the real count is still on the device the last count the host has
        // seen stands in for it (poll_count); the launch itself stays sized for the bound.
        long long est = nwg;
        if (c->count_on_device) {
            poll_count(c);
            if (c->count_hint >= 0) est = std::min(nwg, il_count(c->count_hint / size, il));
        }
        const bool use_pool = unfiltered && (c->kernel_variant == 1 || (c->kernel_variant == 2 && est >= 6144));
        if (c->debug_flags) {
            if (c->count_on_device) return fail(LUMCU_E_STATE, "instrumented update with an unresolved request count");
            if (int rc = ensure_debug(c)) return rc;
            a.dbg_pre_blend = (c->debug_flags & LUMCU_DEBUG_PRE_BLEND) ? c->d_pre_blend : nullptr;
            a.dbg_samples = (c->debug_flags & LUMCU_DEBUG_RAYS) ? c->d_ray_samples : nullptr;
            a.dbg_hits = (c->debug_flags & LUMCU_DEBUG_RAYS) ? c->d_ray_hits : nullptr;
            a.dbg_totals = c->d_totals;
            if (!second_pass) CU(cudaMemsetAsync(c->d_totals, 0, 2 * sizeof(unsigned long long), c->stream));
            if (int rc = (use_pool ? launch_update_pool<true>(c, sv, a, nwg) : launch_update_R<true>(c, sv, a, nwg))) return rc;
            if (second_pass) {
                c->dbg_traced += nwg;
            } else {
                c->dbg_traced = nwg;
                c->dbg_nwg = il.n > 1 ? wg_end - wg_begin : nwg; c->dbg_wg_begin = wg_begin;   // (a shard leaves the other ranks' entries untouched)
            }
        } else {
            if (c->profiling) CU(cudaEventRecord(c->ev_k0, c->stream));
            if (int rc = (use_pool ? launch_update_pool<false>(c, sv, a, nwg) : launch_update_R<false>(c, sv, a, nwg))) return rc;
            if (c->profiling) { CU(cudaEventRecord(c->ev_k1, c->stream)); c->ktimed = true; }
        }
        if (commit)
            if (int rc = commit_local(c, size, shift, wg_begin, wg_end, z_begin, z_end, il, parity)) return rc;
    }
    return 0;
}

constexpr long long PEER_TIMEOUT_CYCLES = 40000000000ll;      // ~20 s at 2 GHz: a rank that died must not hang the others' GPUs

int arrive(LumcuCtx *c) {
    arrive_kernel<<<1, 32, 0, c->stream>>>(c->peer_flags, c->epoch);
    c->launches++;
    CU(cudaGetLastError());
    return 0;
}

ExchangeWait exchange_wait(LumcuCtx *c) {
    return ExchangeWait{c->d_sync_flags, c->peer_n, c->epoch, PEER_TIMEOUT_CYCLES, c->d_status};
}

}  // namespace

extern "C" {

const char *lumcu_last_error(void) { return g_err.c_str(); }
int lumcu_version(void) { return 100; }

int lumcu_create(const LumcuDesc *d, LumcuCtx **out) {
    if (!d || !out) return fail(LUMCU_E_ARG, "null argument");
    *out = nullptr;
    if (d->world_size[0] < 1 || d->world_size[1] < 1 || d->world_size[2] < 1 ||
        d->world_size[0] > 32767 || d->world_size[1] > 32767 || d->world_size[2] > 32767)
        return fail(LUMCU_E_ARG, "world_size out of range");
    if ((long long)d->world_size[0] * d->world_size[1] * d->world_size[2] > (1ll << 31) - 1)
        return fail(LUMCU_E_ARG, "world too large (cells must fit in int32)");
    if (d->palette_capacity < 1 || d->palette_capacity > 4096)
        return fail(LUMCU_E_ARG, "palette_capacity must be in [1,4096]");
    const int R = d->rays_per_probe;
    if (R != 16 && R != 32 && R != 64 && R != 128 && R != 256)
        return fail(LUMCU_E_ARG, "rays_per_probe must be one of 16,32,64,128,256");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(LUMCU_E_CUDA, "no CUDA device (%s); this stage has no CPU fallback",
                    e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    if (d->device < 0 || d->device >= ndev) return fail(LUMCU_E_ARG, "device %d out of range (%d devices)", d->device, ndev);
    LumcuCtx *c = new (std::nothrow) LumcuCtx();
    if (!c) return fail(LUMCU_E_ARG, "out of host memory");
    c->desc = *d;
    c->wx = d->world_size[0]; c->wy = d->world_size[1]; c->wz = d->world_size[2];
    c->R = R;
    c->cells = (long long)c->wx * c->wy * c->wz;
    c->max_requests = d->max_requests > 0 ? d->max_requests : c->cells + 4096;
    if (c->max_requests >= (1ll << 32)) { delete c; return fail(LUMCU_E_ARG, "max_requests must fit 32 bits"); }
    auto bail = [&](int rc) { lumcu_destroy(c); return rc; };
#define CUB(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return bail(fail(LUMCU_E_CUDA, "%s failed: %s", #call, cudaGetErrorString(e_))); } while (0)
    CUB(cudaSetDevice(d->device));
    CUB(cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, d->device));
    CUB(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    CUB(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    CUB(cudaStreamCreateWithFlags(&c->dl_stream, cudaStreamNonBlocking));
    CUB(cudaEventCreateWithFlags(&c->ev_dl_ready, cudaEventDisableTiming));
    CUB(cudaEventCreateWithFlags(&c->ev_dl_done, cudaEventDisableTiming));
    CUB(cudaEventCreateWithFlags(&c->ev_world_retired[0], cudaEventDisableTiming));
    CUB(cudaEventCreateWithFlags(&c->ev_world_retired[1], cudaEventDisableTiming));
    CUB(cudaEventCreateWithFlags(&c->ev_compute, cudaEventDisableTiming));
    CUB(cudaEventCreateWithFlags(&c->ev_copy, cudaEventDisableTiming));
    CUB(cudaEventCreate(&c->ev_begin));
    CUB(cudaEventCreate(&c->ev_end));
    CUB(cudaEventCreate(&c->ev_k0));
    CUB(cudaEventCreate(&c->ev_k1));
    CUB(cudaMalloc(&c->d_world_buf[0], (size_t)c->cells * sizeof(int16_t)));
    c->d_world = c->d_world_buf[0];        // (the second buffer is allocated by the first lumcu_upload_world_async)
    CUB(cudaMalloc(&c->d_scene_flags, 2 * sizeof(uint32_t)));
    CUB(cudaMemsetAsync(c->d_scene_flags, 0, 2 * sizeof(uint32_t), c->stream));
    {
        int lo = 0, hi = 0;
        CUB(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        CUB(cudaStreamCreateWithPriority(&c->prep_stream, cudaStreamNonBlocking, hi));
    }
    CUB(cudaEventCreateWithFlags(&c->ev_prep, cudaEventDisableTiming));
    CUB(cudaEventCreateWithFlags(&c->ev_world_ready, cudaEventDisableTiming));
    CUB(cudaEventCreateWithFlags(&c->ev_scratch, cudaEventDisableTiming));
    CUB(cudaEventCreateWithFlags(&c->ev_list_retired[0], cudaEventDisableTiming));
    CUB(cudaEventCreateWithFlags(&c->ev_list_retired[1], cudaEventDisableTiming));
    CUB(cudaMalloc(&c->d_bricks, (size_t)d->palette_capacity * 4096));
    CUB(cudaMalloc(&c->d_materials, 256 * 6 * sizeof(float)));
    CUB(cudaMalloc(&c->d_mat4, 256 * sizeof(float4)));
    CUB(cudaMalloc(&c->d_field, (size_t)c->cells * sizeof(uint32_t)));
    CUB(cudaMalloc(&c->d_requests_buf[0], (size_t)c->max_requests * sizeof(short4)));
    c->d_requests = c->d_requests_buf[0];  // (the second list and the builder's scratch are allocated on first use)
    CUB(cudaMalloc(&c->d_staged, 2 * (size_t)c->max_requests * sizeof(uint32_t)));   // two halves: the exchange alternates
    CUB(cudaMalloc(&c->d_table, (size_t)R * 3 * sizeof(float)));
    CUB(cudaMemsetAsync(c->d_world, 0, (size_t)c->cells * sizeof(int16_t), c->stream));
    CUB(cudaMemsetAsync(c->d_bricks, 0, (size_t)d->palette_capacity * 4096, c->stream));
    CUB(cudaMemsetAsync(c->d_materials, 0, 256 * 6 * sizeof(float), c->stream));
    CUB(cudaMemsetAsync(c->d_mat4, 0, 256 * sizeof(float4), c->stream));
    CUB(cudaMemsetAsync(c->d_field, 0, (size_t)c->cells * sizeof(uint32_t), c->stream));   // initial field = zero words
    if (const char *e = getenv("LUMCU_POOL_THREADS")) c->pool_threads = atoi(e) == 256 ? 256 : 128;
    if (const char *e = getenv("LUMCU_CARVEOUT")) c->carveout = atoi(e);
    if (const char *e = getenv("LUMCU_KERNEL_VARIANT")) c->kernel_variant = atoi(e);
    if (const char *e = getenv("LUMCU_VEC_COMMIT")) c->vec_commit = atoi(e) != 0;
    if (const char *e = getenv("LUMCU_SHARD_MODE")) c->shard_chunk = atoi(e) == LUMCU_SHARD_CONTIGUOUS ? 0 : IL_CHUNK;
    host_direction_table(R, c->h_table);
    CUB(cudaMemcpyAsync(c->d_table, c->h_table.data(), (size_t)R * 3 * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    {   // normalize(vec3(0.5,0.5,-0.9)) (radiance.comp:42), f32, x->y->z dot
        float x = 0.5f, y = 0.5f, z = -0.9f;
        float l = sqrtf(x * x + y * y + z * z);
        c->sun[0] = x / l; c->sun[1] = y / l; c->sun[2] = z / l;
    }
    CUB(cudaStreamSynchronize(c->stream));
#undef CUB
    *out = c;
    return 0;
}

int lumcu_destroy(LumcuCtx *c) {
    if (!c) return 0;
    cudaSetDevice(c->desc.device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->copy_stream) cudaStreamSynchronize(c->copy_stream);
    if (c->dl_stream) cudaStreamSynchronize(c->dl_stream);
    if (c->prep_stream) cudaStreamSynchronize(c->prep_stream);
    cudaFree(c->d_scene_flags);
    cudaFree(c->d_g_matnorm); cudaFree(c->d_g_depth); cudaFree(c->d_g_frame); cudaFree(c->d_g_cov);
    cudaFree(c->d_world_buf[0]); cudaFree(c->d_world_buf[1]); cudaFree(c->d_bricks); cudaFree(c->d_materials); cudaFree(c->d_mat4); cudaFree(c->d_field);
    cudaFree(c->d_field_tmp); cudaFree(c->d_requests_buf[0]); cudaFree(c->d_requests_buf[1]); cudaFree(c->d_staged); cudaFree(c->d_table);
    cudaFree(c->d_flags); cudaFree(c->d_tile_state); cudaFree(c->d_total_buf[0]); cudaFree(c->d_total_buf[1]);
    lumcu_peers_close(c);
    cudaFree(c->d_sync_flags);
    if (c->h_status) cudaFreeHost(c->h_status);
    cudaFree(c->d_special); cudaFree(c->d_bounds); cudaFree(c->d_pairs);
    for (auto &m : c->models) cudaFree(m.d_vox);
    cudaFree(c->d_pre_blend); cudaFree(c->d_ray_samples); cudaFree(c->d_ray_hits); cudaFree(c->d_totals);
    c->stage_world.release(); c->stage_req.release(); c->stage_misc.release(); c->stage_pairs.release();
    if (c->h_count) cudaFreeHost(c->h_count);
    if (c->ev_count) cudaEventDestroy(c->ev_count);
    for (auto &f : c->fences) if (f) cudaEventDestroy(f);
    if (c->ev_begin) cudaEventDestroy(c->ev_begin);
    if (c->ev_end) cudaEventDestroy(c->ev_end);
    if (c->ev_k0) cudaEventDestroy(c->ev_k0);
    if (c->ev_k1) cudaEventDestroy(c->ev_k1);
    if (c->ev_compute) cudaEventDestroy(c->ev_compute);
    if (c->ev_copy) cudaEventDestroy(c->ev_copy);
    if (c->ev_dl_ready) cudaEventDestroy(c->ev_dl_ready);
    if (c->ev_dl_done) cudaEventDestroy(c->ev_dl_done);
    for (auto &e : c->ev_world_retired) if (e) cudaEventDestroy(e);
    for (cudaEvent_t e : {c->ev_prep, c->ev_world_ready, c->ev_scratch, c->ev_list_retired[0], c->ev_list_retired[1]}) if (e) cudaEventDestroy(e);
    if (c->prep_stream) cudaStreamDestroy(c->prep_stream);
    if (c->dl_stream) cudaStreamDestroy(c->dl_stream);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
    return 0;
}

static int staged_upload(LumcuCtx *c, Staging &st, void *dst, const void *src, size_t bytes) {
    // Caller buffer already page-locked (cudaHostAlloc / cudaHostRegister): copy straight from it on
    // the copy stream -- ordered after the work already queued (the kernels still reading dst) and
    // before whatever is queued next -- and return once the copy itself is done, so the buffer is
    // consumed before the call returns without a host-side staging memcpy.
    cudaPointerAttributes attr{};
    if (bytes >= (64u << 10) && cudaPointerGetAttributes(&attr, src) == cudaSuccess && attr.type == cudaMemoryTypeHost) {
        CU(cudaEventRecord(c->ev_compute, c->stream));
        CU(cudaStreamWaitEvent(c->copy_stream, c->ev_compute, 0));
        CU(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, c->copy_stream));
        CU(cudaEventRecord(c->ev_copy, c->copy_stream));
        CU(cudaStreamWaitEvent(c->stream, c->ev_copy, 0));
        CU(cudaEventSynchronize(c->ev_copy));
        return 0;
    }
    cudaGetLastError();   // cudaPointerGetAttributes on plain pageable memory may leave an error state on old drivers
    void *h = nullptr; int slot = 0;
    if (int rc = st.acquire(bytes, &h, &slot)) return rc;
    memcpy(h, src, bytes);
    CU(cudaMemcpyAsync(dst, h, bytes, cudaMemcpyHostToDevice, c->stream));
    CU(cudaEventRecord(st.done[slot], c->stream));
    return 0;
}

int lumcu_upload_world(LumcuCtx *c, const int16_t *world) {
    if (int rc = use(c)) return rc;
    if (!world) return fail(LUMCU_E_ARG, "null world");
    if (int rc = staged_upload(c, c->stage_world, c->d_world, world, (size_t)c->cells * sizeof(int16_t))) return rc;
    CU(cudaEventRecord(c->ev_world_ready, c->stream));
    c->world_ready_pending = true;
    c->have_world = true;
    c->world_dirty = true;
    return 0;
}

int lumcu_upload_world_async(LumcuCtx *c, const int16_t *world) {
    if (int rc = use(c)) return rc;
    if (!world) return fail(LUMCU_E_ARG, "null world");
    const size_t bytes = (size_t)c->cells * sizeof(int16_t);
    cudaPointerAttributes attr{};
    if (cudaPointerGetAttributes(&attr, world) != cudaSuccess || attr.type != cudaMemoryTypeHost) {
        cudaGetLastError();
        return lumcu_upload_world(c, world);                       // pageable memory: the synchronous path
    }
    // The frame in flight still reads the current buffer, so the copy goes into the OTHER one, on the copy stream,
    // beside that frame's tracing.  It only has to wait for whoever last read the other buffer -- everything that was
    // queued before the previous flip (ev_world_retired) -- and the launches queued from here on wait for the copy.
    const int cur = c->world_cur, nxt = cur ^ 1;
    if (!c->d_world_buf[nxt]) CU(cudaMalloc(&c->d_world_buf[nxt], bytes));
    CU(cudaEventRecord(c->ev_world_retired[cur], c->stream));      // every reader of `cur` has been queued by now
    c->world_retired[cur] = true;
    if (c->world_retired[nxt]) CU(cudaStreamWaitEvent(c->copy_stream, c->ev_world_retired[nxt], 0));
    CU(cudaMemcpyAsync(c->d_world_buf[nxt], world, bytes, cudaMemcpyHostToDevice, c->copy_stream));
    CU(cudaEventRecord(c->ev_copy, c->copy_stream));
    CU(cudaEventRecord(c->ev_world_ready, c->copy_stream));
    c->world_ready_pending = true;
    CU(cudaStreamWaitEvent(c->stream, c->ev_copy, 0));
    c->world_cur = nxt;
    c->d_world = c->d_world_buf[nxt];
    c->have_world = true;
    c->world_dirty = true;
    return 0;
}

int lumcu_upload_bricks(LumcuCtx *c, int32_t first, int32_t count, const uint8_t *vox) {
    if (int rc = use(c)) return rc;
    if (!vox || count < 0 || first < 0 || first + count > c->desc.palette_capacity)
        return fail(LUMCU_E_ARG, "brick range [%d,%d) outside palette capacity %d", first, first + count, c->desc.palette_capacity);
    if (count > 0) {
        if (int rc = staged_upload(c, c->stage_misc, c->d_bricks + (size_t)first * 4096, vox, (size_t)count * 4096)) return rc;
        if (first == 0) {
            bool empty = true;
            for (int i = 0; i < 4096 && empty; i++) empty = vox[i] == 0;
            c->brick0_empty = empty;
        }
    }
    c->have_bricks = true;
    return 0;
}

int lumcu_upload_materials(LumcuCtx *c, const float *m) {
    if (int rc = use(c)) return rc;
    if (!m) return fail(LUMCU_E_ARG, "null materials");
    if (int rc = staged_upload(c, c->stage_misc, c->d_materials, m, 256 * 6 * sizeof(float))) return rc;
    float packed[256 * 4];     // device mirror for the tracer: the floats a hit uses (r, g, b, emmit) side by side
    for (int i = 0; i < 256; i++) {
        packed[4 * i + 0] = m[6 * i + 0]; packed[4 * i + 1] = m[6 * i + 1];
        packed[4 * i + 2] = m[6 * i + 2]; packed[4 * i + 3] = m[6 * i + 4];
    }
    if (int rc = staged_upload(c, c->stage_misc, c->d_mat4, packed, sizeof packed)) return rc;
    c->have_materials = true;
    return 0;
}


int lumcu_upload_model(LumcuCtx *c, const uint8_t *vox, int32_t sx, int32_t sy, int32_t sz, int32_t *out) {
    if (int rc = use(c)) return rc;
    if (!vox || !out || sx < 1 || sy < 1 || sz < 1 || (long long)sx * sy * sz > (1ll << 30))
        return fail(LUMCU_E_ARG, "bad model (%d x %d x %d)", sx, sy, sz);
    const size_t bytes = (size_t)sx * sy * sz;
    LumcuCtx::Model m;
    CU(cudaMalloc(&m.d_vox, bytes));
    m.sx = sx; m.sy = sy; m.sz = sz;
    if (int rc = staged_upload(c, c->stage_misc, m.d_vox, vox, bytes)) { cudaFree(m.d_vox); return rc; }
    int slot = -1;
    for (size_t i = 0; i < c->models.size(); i++) if (!c->models[i].d_vox) { slot = (int)i; break; }
    if (slot < 0) { c->models.push_back(m); slot = (int)c->models.size() - 1; } else c->models[slot] = m;
    *out = slot;
    return 0;
}

int lumcu_free_model(LumcuCtx *c, int32_t model) {
    if (int rc = use(c)) return rc;
    if (model < 0 || model >= (int)c->models.size() || !c->models[model].d_vox) return fail(LUMCU_E_ARG, "unknown model %d", model);
    CU(cudaStreamSynchronize(c->stream));
    cudaFree(c->models[model].d_vox);
    c->models[model] = LumcuCtx::Model();
    return 0;
}

int lumcu_copy_bricks(LumcuCtx *c, const int32_t *src, const int32_t *dst, int32_t n) {
    if (int rc = use(c)) return rc;
    if (n < 0 || (n > 0 && (!src || !dst))) return fail(LUMCU_E_ARG, "bad brick copy list");
    if (n == 0) return 0;
    for (int i = 0; i < n; i++)
        if (src[i] < 0 || src[i] >= c->desc.palette_capacity || dst[i] < 0 || dst[i] >= c->desc.palette_capacity)
            return fail(LUMCU_E_ARG, "brick copy %d: %d -> %d outside palette capacity %d", i, src[i], dst[i], c->desc.palette_capacity);
    {   // the copies run concurrently (one CTA per pair): a destination may be neither a source nor written twice
        std::vector<uint8_t> role((size_t)c->desc.palette_capacity, 0);      // 1 = source, 2 = destination
        for (int i = 0; i < n; i++) role[src[i]] |= 1;
        for (int i = 0; i < n; i++) {
            if (role[dst[i]] & 1) return fail(LUMCU_E_ARG, "brick copy %d: destination %d is also a source", i, dst[i]);
            if (role[dst[i]] & 2) return fail(LUMCU_E_ARG, "brick copy %d: destination %d appears twice", i, dst[i]);
            role[dst[i]] |= 2;
        }
    }
    if (c->pairs_cap < n) {
        cudaFree(c->d_pairs); c->d_pairs = nullptr; c->pairs_cap = 0;
        long long cap = std::max<long long>(n, 256);
        CU(cudaMalloc(&c->d_pairs, (size_t)cap * sizeof(int2)));
        c->pairs_cap = cap;
    }
    void *h = nullptr; int slot = 0;
    if (int rc = c->stage_pairs.acquire((size_t)n * sizeof(int2), &h, &slot)) return rc;
    int2 *hp = (int2 *)h;
    for (int i = 0; i < n; i++) { hp[i].x = src[i]; hp[i].y = dst[i]; if (dst[i] == 0) c->brick0_empty = false; }
    CU(cudaMemcpyAsync(c->d_pairs, hp, (size_t)n * sizeof(int2), cudaMemcpyHostToDevice, c->stream));
    CU(cudaEventRecord(c->stage_pairs.done[slot], c->stream));
    copy_bricks_kernel<<<(unsigned)n, 256, 0, c->stream>>>(c->d_bricks, c->d_pairs, n, c->desc.palette_capacity);
    c->launches++;
    CU(cudaGetLastError());
    c->have_bricks = true;
    return 0;
}

int lumcu_map_model(LumcuCtx *c, int32_t model, const float inv[16], const int32_t shift[3], const int32_t area[3], int32_t static_count) {
    if (int rc = use(c)) return rc;
    if (model < 0 || model >= (int)c->models.size() || !c->models[model].d_vox) return fail(LUMCU_E_ARG, "unknown model %d", model);
    if (!inv || !shift || !area) return fail(LUMCU_E_ARG, "null argument");
    if (!c->have_world) return fail(LUMCU_E_STATE, "lumcu_map_model before lumcu_upload_world");
    MapArgs a;
    for (int i = 0; i < 16; i++) a.m[i] = inv[i];
    long long total = 1;
    for (int k = 0; k < 3; k++) {
        a.shift[k] = shift[k];
        // vkCmdDispatch((area+3)/4) with a negative or zero area dispatches nothing (internal_render.cpp:1517)
        const int groups = (area[k] + 3) / 4;
        a.ext[k] = groups > 0 ? groups * 4 : 0;
        total *= a.ext[k];
    }
    if (total > (1ll << 40)) return fail(LUMCU_E_ARG, "map area too large");
    if (static_count <= 0) c->brick0_empty = false;    // brick 0 becomes a store target (map.comp:62 with a zero threshold)
    a.static_count = static_count;
    const LumcuCtx::Model &m = c->models[model];
    a.msx = m.sx; a.msy = m.sy; a.msz = m.sz;
    if (total > 0) {
        const int t = 256;
        map_model_kernel<<<(unsigned)((total + t - 1) / t), t, 0, c->stream>>>(c->d_world, c->d_bricks, m.d_vox, c->wx, c->wy, c->wz, c->desc.palette_capacity, a);
        c->launches++;
        CU(cudaGetLastError());
    }
    return 0;
}

int lumcu_download_bricks(LumcuCtx *c, int32_t first, int32_t count, uint8_t *vox) {
    if (int rc = use(c)) return rc;
    if (!vox || count < 0 || first < 0 || first + count > c->desc.palette_capacity)
        return fail(LUMCU_E_ARG, "brick range [%d,%d) outside palette capacity %d", first, first + count, c->desc.palette_capacity);
    if (count > 0) CU(cudaMemcpyAsync(vox, c->d_bricks + (size_t)first * 4096, (size_t)count * 4096, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return 0;
}

int lumcu_download_world(LumcuCtx *c, int16_t *world) {
    if (int rc = use(c)) return rc;
    if (!world) return fail(LUMCU_E_ARG, "null world");
    CU(cudaMemcpyAsync(world, c->d_world, (size_t)c->cells * sizeof(int16_t), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return 0;
}

int lumcu_set_requests(LumcuCtx *c, const int16_t *xyzw, int64_t n) {
    if (int rc = use(c)) return rc;
    if (n < 0 || n > c->max_requests) return fail(LUMCU_E_ARG, "request count %lld exceeds max_requests %lld", (long long)n, c->max_requests);
    if (n > 0 && !xyzw) return fail(LUMCU_E_ARG, "null requests");
    if (n > 0) {
        if (int rc = staged_upload(c, c->stage_req, c->d_requests, xyzw, (size_t)n * sizeof(short4))) return rc;
        long long s = 1;
        while (s < n && xyzw[4 * s + 2] >= xyzw[4 * (s - 1) + 2]) s++;
        c->nsorted = s;
    } else {
        c->nsorted = 0;
    }
    c->nreq = n;
    c->shard_basis = n;
    c->count_hint = n;
    c->count_on_device = false;
    c->req_version++;
    return 0;
}

int lumcu_build_requests(LumcuCtx *c, const int16_t *special, int32_t nspecial, int64_t *out_count) {
    if (int rc = use(c)) return rc;
    if (!c->have_world) return fail(LUMCU_E_STATE, "lumcu_build_requests before lumcu_upload_world");
    if (nspecial < 0 || (nspecial > 0 && !special)) return fail(LUMCU_E_ARG, "bad special requests");
    const long long ntiles = (c->cells + BUILD_TILE - 1) / BUILD_TILE;
    if (!c->d_flags) {
        CU(cudaMalloc(&c->d_flags, (size_t)c->cells));
        CU(cudaMalloc(&c->d_tile_state, ((size_t)ntiles + 1) * sizeof(unsigned long long)));     // [ntiles] = the ticket counter
    }
    // A list nobody waits for (no count wanted, no specials to upload first) is built AHEAD: into the other list buffer,
    // on the high-priority prep stream, as soon as the world it is built from has arrived -- beside the trace of the
    // frame in flight, which still reads the current list -- together with that world's validation.  The stage's
    // stream waits for it here, so everything queued from now on sees the new list.  Otherwise: in place, in order.
    const bool ahead = !out_count && nspecial == 0;
    const int tgt = ahead ? c->req_cur ^ 1 : c->req_cur;
    if (!c->d_requests_buf[tgt]) CU(cudaMalloc(&c->d_requests_buf[tgt], (size_t)c->max_requests * sizeof(short4)));
    if (!c->d_total_buf[tgt]) CU(cudaMalloc(&c->d_total_buf[tgt], 2 * sizeof(unsigned long long)));
    cudaStream_t st = ahead ? c->prep_stream : c->stream;
    if (ahead) {
        CU(cudaEventRecord(c->ev_list_retired[c->req_cur], c->stream));        // every reader of the current list has been queued by now
        c->list_retired[c->req_cur] = true;
        if (c->list_retired[tgt]) CU(cudaStreamWaitEvent(st, c->ev_list_retired[tgt], 0));
        if (c->world_ready_pending) { CU(cudaStreamWaitEvent(st, c->ev_world_ready, 0)); c->world_ready_pending = false; }
        if (c->scratch_on_stream) { CU(cudaStreamWaitEvent(st, c->ev_scratch, 0)); c->scratch_on_stream = false; }
        if (c->world_dirty && !c->world_shared) {
            if (int rc = launch_validate(c, st)) return rc;
            c->world_dirty = false;
        }
    }
    CU(cudaMemsetAsync(c->d_tile_state, 0, ((size_t)ntiles + 1) * sizeof(unsigned long long), st));
    build_fused_kernel<<<(unsigned)ntiles, BUILD_THREADS, 0, st>>>(c->d_world, c->wx, c->wy, c->wz, c->d_flags, c->d_tile_state,
                                                                  (unsigned int *)(c->d_tile_state + ntiles), c->d_requests_buf[tgt],
                                                                  c->max_requests, c->d_total_buf[tgt]);
    c->launches++;
    CU(cudaGetLastError());
    if (ahead) {
        CU(cudaEventRecord(c->ev_prep, st));
        CU(cudaStreamWaitEvent(c->stream, c->ev_prep, 0));
        c->req_cur = tgt;
    } else {
        if (nspecial > 0) {
            if (c->special_cap < nspecial) {
                cudaFree(c->d_special); c->d_special = nullptr; c->special_cap = 0;
                CU(cudaMalloc(&c->d_special, (size_t)nspecial * sizeof(short4)));
                c->special_cap = nspecial;
            }
            if (int rc = staged_upload(c, c->stage_misc, c->d_special, special, (size_t)nspecial * sizeof(short4))) return rc;
            build_specials_kernel<<<1, 1024, 0, c->stream>>>(c->d_flags, c->wx, c->wy, c->wz, c->d_special, nspecial, c->d_requests_buf[tgt], c->d_total_buf[tgt], c->max_requests);
            c->launches++;
            CU(cudaGetLastError());
        }
        CU(cudaEventRecord(c->ev_scratch, c->stream));
        c->scratch_on_stream = true;
    }
    c->d_requests = c->d_requests_buf[tgt];
    c->d_total = c->d_total_buf[tgt];
    {   // the count follows the host at its own pace: a 16-byte copy into pinned memory that the next frames poll
        if (!c->h_count) { CU(cudaHostAlloc(&c->h_count, 2 * sizeof(unsigned long long), cudaHostAllocDefault)); CU(cudaEventCreateWithFlags(&c->ev_count, cudaEventDisableTiming)); }
        poll_count(c);                                               // (do not lose an earlier count that has arrived)
        CU(cudaMemcpyAsync(c->h_count, c->d_total, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ahead ? c->prep_stream : c->stream));
        CU(cudaEventRecord(c->ev_count, ahead ? c->prep_stream : c->stream));
        c->count_copy_pending = true;
    }
    // The count stays on the device: the kernels of the frame cut their dispatch down to it, and nothing here waits
    // for the GPU.  Until somebody asks (out_count, lumcu_request_count, ...) the host works with an upper bound.
    c->count_on_device = true;
    c->nreq = std::min<long long>(c->max_requests, c->cells + nspecial);
    c->shard_basis = c->nreq;
    c->nsorted = 0;
    c->req_version++;
    if (out_count) {
        if (int rc = resolve_count(c)) return rc;
        *out_count = (int64_t)c->nreq;
    }
    return 0;
}

int lumcu_request_count(LumcuCtx *c, int64_t *out) {
    if (!c || !out) return fail(LUMCU_E_ARG, "null argument");
    if (c->count_on_device) { if (int rc = use(c)) return rc; if (int rc = resolve_count(c)) return rc; }
    *out = c->nreq;
    return 0;
}

int lumcu_download_requests(LumcuCtx *c, int16_t *out, int64_t cap) {
    if (int rc = use(c)) return rc;
    if (int rc = resolve_count(c)) return rc;
    if (!out || cap < c->nreq) return fail(LUMCU_E_ARG, "output too small for %lld requests", c->nreq);
    if (c->nreq > 0) CU(cudaMemcpyAsync(out, c->d_requests, (size_t)c->nreq * sizeof(short4), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return 0;
}

int lumcu_update(LumcuCtx *c, int32_t time_seed, int32_t size, int32_t shift) {
    if (int rc = use(c)) return rc;
    if (size < 1) return fail(LUMCU_E_ARG, "size must be >= 1");
    if (c->debug_flags) if (int rc = resolve_count(c)) return rc;      // the debug buffers are sized and read by exact counts
    CU(cudaEventRecord(c->ev_begin, c->stream));
    if (int rc = do_update(c, time_seed, size, shift, 0, c->nreq / size, -32768, 32768)) return rc;
    CU(cudaEventRecord(c->ev_end, c->stream));
    c->timed = true;
    return 0;
}

int lumcu_update_range(LumcuCtx *c, int32_t time_seed, int32_t size, int32_t shift, int64_t wg_begin, int64_t wg_end) {
    if (int rc = use(c)) return rc;
    if (int rc = resolve_count(c)) return rc;
    CU(cudaEventRecord(c->ev_begin, c->stream));
    if (int rc = do_update(c, time_seed, size, shift, wg_begin, wg_end, -32768, 32768)) return rc;
    CU(cudaEventRecord(c->ev_end, c->stream));
    c->timed = true;
    return 0;
}

static long long sorted_workgroups(const LumcuCtx *c, int size, int shift) {
    long long nwg_all = c->nreq / size;
    long long n = c->nsorted > shift ? (c->nsorted - shift + size - 1) / size : 0;
    return std::min(nwg_all, n);
}

int lumcu_slab_range(LumcuCtx *c, int32_t size, int32_t shift, int32_t z_begin, int32_t z_end, int64_t *wg_begin, int64_t *wg_end) {
    if (int rc = use(c)) return rc;
    if (size < 1 || shift < 0 || shift >= size || !wg_begin || !wg_end || z_begin > z_end)
        return fail(LUMCU_E_ARG, "bad slab arguments");
    if (int rc = resolve_count(c)) return rc;
    // request coordinates are int16: anything outside means "unbounded" (slab_bound_kernel's predecessor of the
    // first workgroup is z = -32768, which must compare below every bound that can select something)
    z_begin = std::min(std::max(z_begin, -32767), 32768);
    z_end = std::min(std::max(z_end, -32767), 32768);
    if (c->slab_key_version == c->req_version && c->slab_key[0] == size && c->slab_key[1] == shift &&
        c->slab_key[2] == z_begin && c->slab_key[3] == z_end) {
        *wg_begin = c->slab_val[0]; *wg_end = c->slab_val[1];
        return 0;
    }
    const long long nwg_sorted = sorted_workgroups(c, size, shift);
    if (!c->d_bounds) CU(cudaMalloc(&c->d_bounds, 2 * sizeof(long long)));
    long long init[2] = {nwg_sorted, nwg_sorted}, res[2];
    CU(cudaMemcpyAsync(c->d_bounds, init, sizeof init, cudaMemcpyHostToDevice, c->stream));
    if (nwg_sorted > 0) {
        const int t = 256;
        slab_bound_kernel<<<(unsigned)((nwg_sorted + t - 1) / t), t, 0, c->stream>>>(c->d_requests, nwg_sorted, size, shift, z_begin, z_end, c->d_bounds);
        c->launches++;
        CU(cudaGetLastError());
    }
    CU(cudaMemcpyAsync(res, c->d_bounds, sizeof res, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    c->slab_key_version = c->req_version;
    c->slab_key[0] = size; c->slab_key[1] = shift; c->slab_key[2] = z_begin; c->slab_key[3] = z_end;
    c->slab_val[0] = res[0]; c->slab_val[1] = res[1];
    *wg_begin = res[0];
    *wg_end = res[1];
    return 0;
}

int lumcu_update_slab(LumcuCtx *c, int32_t time_seed, int32_t size, int32_t shift, int32_t z_begin, int32_t z_end) {
    if (int rc = use(c)) return rc;
    int64_t b = 0, e = 0;
    if (int rc = lumcu_slab_range(c, size, shift, z_begin, z_end, &b, &e)) return rc;
    const long long nwg_all = c->nreq / size, tail = sorted_workgroups(c, size, shift);
    CU(cudaEventRecord(c->ev_begin, c->stream));
    // Jacobi: both passes trace against the field as it was before the frame (the tail's rays may land in texels the
    // sorted range is about to replace), so nothing is committed until both have run.
    if (int rc = do_update(c, time_seed, size, shift, b, e, -32768, 32768, false)) return rc;
    if (tail < nwg_all)   // unsorted tail (specials): every rank scans it, keeps its slab's probes
        if (int rc = do_update(c, time_seed, size, shift, tail, nwg_all, z_begin, z_end, false, Interleave{IL_CHUNK, 0, 1}, true)) return rc;
    if (int rc = commit_local(c, size, shift, b, e, -32768, 32768)) return rc;
    if (tail < nwg_all)
        if (int rc = commit_local(c, size, shift, tail, nwg_all, z_begin, z_end)) return rc;
    CU(cudaEventRecord(c->ev_end, c->stream));
    c->timed = true;
    return 0;
}

int lumcu_ipc_export(LumcuCtx *c, uint8_t *handle) {
    if (int rc = use(c)) return rc;
    if (!handle) return fail(LUMCU_E_ARG, "null argument");
    if (!c->d_sync_flags) {
        CU(cudaMalloc(&c->d_sync_flags, MAX_PEERS * sizeof(unsigned)));
        CU(cudaMemset(c->d_sync_flags, 0, MAX_PEERS * sizeof(unsigned)));
        CU(cudaHostAlloc(&c->h_status, sizeof(unsigned), cudaHostAllocMapped));
        *c->h_status = 0;
        CU(cudaHostGetDevicePointer(&c->d_status, c->h_status, 0));
    }
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaIpcMemHandle_t h;
    CU(cudaIpcGetMemHandle(&h, c->d_staged));          // the inbox: both halves of staged[]
    memcpy(handle, &h, 64);
    CU(cudaIpcGetMemHandle(&h, c->d_sync_flags));
    memcpy(handle + 64, &h, 64);
    return 0;
}

int lumcu_peers_close(LumcuCtx *c) {
    if (!c) return 0;
    if (c->have_peers || c->n_peer_opened > 0) {   // also after a lumcu_peers_open that failed half-way
        cudaSetDevice(c->desc.device);
        if (c->stream) cudaStreamSynchronize(c->stream);
        for (int i = 0; i < c->n_peer_opened; i++) cudaIpcCloseMemHandle(c->peer_opened[i]);
    }
    c->n_peer_opened = 0;
    c->have_peers = false;
    return 0;
}

int lumcu_peers_open(LumcuCtx *c, int32_t rank, int32_t nranks, const uint8_t *handles) {
    if (int rc = use(c)) return rc;
    if (nranks < 1 || nranks > MAX_PEERS || rank < 0 || rank >= nranks || !handles)
        return fail(LUMCU_E_ARG, "bad peer set (rank %d of %d, at most %d ranks)", rank, nranks, MAX_PEERS);
    if (!c->d_sync_flags) return fail(LUMCU_E_STATE, "lumcu_peers_open before lumcu_ipc_export");
    lumcu_peers_close(c);
    c->peer_inbox[0] = PeerWords{}; c->peer_inbox[1] = PeerWords{};
    c->peer_flags = PeerFlags{};
    // every rank must have created its stage with the same max_requests (the halves of staged[] are max_requests apart)
    for (int p = 0; p < nranks; p++) {
        if (p == rank) {
            c->peer_inbox[0].p[p] = staged_half(c, 0); c->peer_inbox[1].p[p] = staged_half(c, 1);
            c->peer_flags.f[p] = c->d_sync_flags;
            continue;
        }
        cudaIpcMemHandle_t h;
        void *ptr = nullptr;
        memcpy(&h, handles + 128 * p, 64);
        CU(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
        c->peer_opened[c->n_peer_opened++] = ptr;
        c->peer_inbox[0].p[p] = (uint32_t *)ptr; c->peer_inbox[1].p[p] = (uint32_t *)ptr + (size_t)c->max_requests;
        memcpy(&h, handles + 128 * p + 64, 64);
        CU(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
        c->peer_opened[c->n_peer_opened++] = ptr;
        c->peer_flags.f[p] = (unsigned *)ptr;
    }
    c->peer_inbox[0].n = nranks; c->peer_inbox[1].n = nranks;
    c->peer_flags.n = nranks;
    c->peer_flags.rank = rank;
    c->peer_rank = rank; c->peer_n = nranks;
    c->have_peers = true;
    return 0;
}

int lumcu_update_shard(LumcuCtx *c, int32_t time_seed, int32_t size, int32_t shift, int32_t rank, int32_t nranks) {
    if (int rc = use(c)) return rc;
    if (nranks < 1 || rank < 0 || rank >= nranks) return fail(LUMCU_E_ARG, "bad shard %d of %d", rank, nranks);
    if (size < 1) return fail(LUMCU_E_ARG, "size must be >= 1");
    if (c->debug_flags) if (int rc = resolve_count(c)) return rc;
    CU(cudaEventRecord(c->ev_begin, c->stream));
    if (int rc = do_update(c, time_seed, size, shift, 0, c->nreq / size, -32768, 32768, true, shard_of(c, c->shard_basis / size, rank, nranks))) return rc;
    CU(cudaEventRecord(c->ev_end, c->stream));
    c->timed = true;
    return 0;
}

// One multi-GPU frame (epoch e, parity = e & 1), workgroup shards:
//   trace     this rank's workgroups into half `parity` of its own staged[] -- nothing remote;
//   arrive    a one-warp kernel: system fence, then e into every rank's flag array;
//   gather + commit   one persistent kernel: waits (inside) until every rank has arrived at e, then reads the texel of
//             every workgroup of the dispatch from its owner's staged[] (NVLink, full lines) and stores it into the
//             local field.
// One barrier per frame, no collective call.  The next frame's trace reads only local memory (stream order suffices)
// and writes the other half of staged[]; a rank overwrites half p again (epoch e + 2) only after the barrier of epoch
// e + 1, which every peer reaches after its gather of epoch e -- the last reader of half p -- in stream order.
// (z-slabs, lumcu_update_slab_exchange: the same frame with the transfer as a PUSH from the trace kernel's last phase
// into every rank's inbox, and a local commit behind the wait.)
static int exchange_frame(LumcuCtx *c, int time_seed, int size, int shift, long long b, long long e, bool slab,
                          int z_begin, int z_end, Interleave il) {
    const long long nwg_all = c->nreq / size, tail = slab ? sorted_workgroups(c, size, shift) : nwg_all;
    c->epoch++;
    const int parity = (int)(c->epoch & 1u);
    // z-slabs PUSH (the trace stores into every rank's inbox: a slab's owner is not a function of the workgroup index);
    // workgroup shards PULL (gather_commit_kernel reads each texel from its owner's staged[])
    const PeerWords *inbox = slab ? &c->peer_inbox[parity] : nullptr;
    CU(cudaEventRecord(c->ev_begin, c->stream));
    if (int rc = do_update(c, time_seed, size, shift, b, e, -32768, 32768, false, il, false, inbox, parity)) return rc;
    if (slab && tail < nwg_all)      // unsorted tail (specials): every rank scans it, keeps its slab's probes
        if (int rc = do_update(c, time_seed, size, shift, tail, nwg_all, z_begin, z_end, false, Interleave{IL_CHUNK, 0, 1}, true, inbox, parity)) return rc;
    if (int rc = arrive(c)) return rc;
    if (slab) {
        if (int rc = commit_local(c, size, shift, 0, nwg_all, -32768, 32768, Interleave{IL_CHUNK, 0, 1}, parity, exchange_wait(c))) return rc;
    } else if (nwg_all > 0) {
        if (int rc = before_field_write(c)) return rc;
        const ShardSources src{c->peer_inbox[parity], il.chunk, il.n};
        const int ct = 256;
        const long long blocks = std::min<long long>((nwg_all + ct - 1) / ct, (long long)c->sm_count * 8);
        gather_commit_kernel<<<(unsigned)blocks, ct, 0, c->stream>>>(c->d_field, src, c->d_requests, 0, nwg_all, size, shift,
                                                                    c->wx, c->wy, c->wz, count_ptr(c), exchange_wait(c));
        c->launches++;
        CU(cudaGetLastError());
    }
    CU(cudaEventRecord(c->ev_end, c->stream));
    c->timed = true;
    return 0;
}

int lumcu_update_shard_exchange(LumcuCtx *c, int32_t time_seed, int32_t size, int32_t shift) {
    if (int rc = use(c)) return rc;
    if (!c->have_peers) return fail(LUMCU_E_STATE, "lumcu_update_shard_exchange before lumcu_peers_open");
    if (*c->h_status) return fail(LUMCU_E_STATE, "a peer barrier timed out earlier (a rank stopped participating)");
    if (size < 1) return fail(LUMCU_E_ARG, "size must be >= 1");
    if (c->debug_flags) if (int rc = resolve_count(c)) return rc;
    return exchange_frame(c, time_seed, size, shift, 0, c->nreq / size, false, -32768, 32768, shard_of(c, c->shard_basis / size, c->peer_rank, c->peer_n));
}

int lumcu_update_slab_exchange(LumcuCtx *c, int32_t time_seed, int32_t size, int32_t shift, int32_t z_begin, int32_t z_end) {
    if (int rc = use(c)) return rc;
    if (!c->have_peers) return fail(LUMCU_E_STATE, "lumcu_update_slab_exchange before lumcu_peers_open");
    if (*c->h_status) return fail(LUMCU_E_STATE, "a peer barrier timed out earlier (a rank stopped participating)");
    int64_t b = 0, e = 0;
    if (int rc = lumcu_slab_range(c, size, shift, z_begin, z_end, &b, &e)) return rc;      // (resolves the request count)
    return exchange_frame(c, time_seed, size, shift, b, e, true, z_begin, z_end, Interleave{IL_CHUNK, 0, 1});
}

int lumcu_shift(LumcuCtx *c, int32_t dx, int32_t dy, int32_t dz) {
    if (int rc = use(c)) return rc;
    if ((dx == 0 && dy == 0 && dz == 0) || abs(dx) >= c->wx || abs(dy) >= c->wy || abs(dz) >= c->wz) return 0;
    if (int rc = before_field_write(c)) return rc;
    if (!c->d_field_tmp) CU(cudaMalloc(&c->d_field_tmp, (size_t)c->cells * sizeof(uint32_t)));
    CU(cudaMemcpyAsync(c->d_field_tmp, c->d_field, (size_t)c->cells * sizeof(uint32_t), cudaMemcpyDeviceToDevice, c->stream));
    const int t = 256;
    shift_kernel<<<(unsigned)((c->cells + t - 1) / t), t, 0, c->stream>>>(c->d_field, c->d_field_tmp, c->wx, c->wy, c->wz, dx, dy, dz);
    c->launches++;
    CU(cudaGetLastError());
    return 0;
}

int lumcu_upload_field(LumcuCtx *c, const uint32_t *f) {
    if (int rc = use(c)) return rc;
    if (!f) return fail(LUMCU_E_ARG, "null field");
    if (int rc = before_field_write(c)) return rc;
    return staged_upload(c, c->stage_misc, c->d_field, f, (size_t)c->cells * sizeof(uint32_t));
}

int lumcu_download_field(LumcuCtx *c, uint32_t *f) {
    if (int rc = use(c)) return rc;
    if (!f) return fail(LUMCU_E_ARG, "null field");
    CU(cudaMemcpyAsync(f, c->d_field, (size_t)c->cells * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return 0;
}

int lumcu_download_field_async(LumcuCtx *c, uint32_t *out, int64_t first, int64_t count) {
    if (int rc = use(c)) return rc;
    if (!out || first < 0 || count < 0 || first + count > c->cells) return fail(LUMCU_E_ARG, "bad field range [%lld, +%lld)", (long long)first, (long long)count);
    if (count == 0) return 0;
    // on its own stream, behind what is queued so far (the commit that produced the field) and beside what comes next
    // (the next frame's upload, build and trace, none of which write the field; its commit waits: before_field_write)
    CU(cudaEventRecord(c->ev_dl_ready, c->stream));
    CU(cudaStreamWaitEvent(c->dl_stream, c->ev_dl_ready, 0));
    CU(cudaMemcpyAsync(out, c->d_field + first, (size_t)count * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->dl_stream));
    CU(cudaEventRecord(c->ev_dl_done, c->dl_stream));
    c->dl_pending = true;
    return 0;
}

int lumcu_record_fence(LumcuCtx *c, int32_t slot) {
    if (int rc = use(c)) return rc;
    if (slot < 0 || slot >= LUMCU_FENCES) return fail(LUMCU_E_ARG, "fence slot %d outside [0,%d)", slot, LUMCU_FENCES);
    for (int k = 0; k < 2; k++)
        if (!c->fences[2 * slot + k]) CU(cudaEventCreateWithFlags(&c->fences[2 * slot + k], cudaEventDisableTiming));
    CU(cudaEventRecord(c->fences[2 * slot], c->stream));           // the stage's stream ...
    CU(cudaEventRecord(c->fences[2 * slot + 1], c->dl_stream));    // ... and the downloads queued so far
    return 0;
}

int lumcu_wait_fence(LumcuCtx *c, int32_t slot) {
    if (int rc = use(c)) return rc;
    if (slot < 0 || slot >= LUMCU_FENCES) return fail(LUMCU_E_ARG, "fence slot %d outside [0,%d)", slot, LUMCU_FENCES);
    for (int k = 0; k < 2; k++)
        if (c->fences[2 * slot + k]) CU(cudaEventSynchronize(c->fences[2 * slot + k]));
    return 0;
}

int lumcu_direction_table(LumcuCtx *c, float *out) {
    if (!c || !out) return fail(LUMCU_E_ARG, "null argument");
    memcpy(out, c->h_table.data(), c->h_table.size() * sizeof(float));
    return 0;
}

int lumcu_set_debug(LumcuCtx *c, int32_t flags) {
    if (int rc = use(c)) return rc;
    c->debug_flags = flags & (LUMCU_DEBUG_PRE_BLEND | LUMCU_DEBUG_RAYS);
    if (flags & ~(LUMCU_DEBUG_PRE_BLEND | LUMCU_DEBUG_RAYS)) return fail(LUMCU_E_ARG, "unknown debug flags 0x%x", flags);
    return 0;
}

int lumcu_debug_extent(LumcuCtx *c, int64_t *nwg, int64_t *wg_begin) {
    if (!c || !nwg) return fail(LUMCU_E_ARG, "null argument");
    *nwg = c->dbg_nwg;
    if (wg_begin) *wg_begin = c->dbg_wg_begin;
    return 0;
}

int lumcu_download_debug(LumcuCtx *c, float *pre_blend, uint32_t *ray_samples, uint8_t *ray_hits, uint64_t totals[3]) {
    if (int rc = use(c)) return rc;
    if (!c->d_totals) return fail(LUMCU_E_STATE, "no instrumented update has run (lumcu_set_debug first)");
    const long long n = c->dbg_nwg, b = c->dbg_wg_begin;
    if (pre_blend) {
        if (!c->d_pre_blend) return fail(LUMCU_E_STATE, "LUMCU_DEBUG_PRE_BLEND was not enabled");
        CU(cudaMemcpyAsync(pre_blend, c->d_pre_blend + 3 * b, (size_t)n * 3 * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    }
    if (ray_samples || ray_hits) {
        if (!c->d_ray_samples) return fail(LUMCU_E_STATE, "LUMCU_DEBUG_RAYS was not enabled");
        if (ray_samples) CU(cudaMemcpyAsync(ray_samples, c->d_ray_samples + b * c->R, (size_t)n * c->R * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
        if (ray_hits) CU(cudaMemcpyAsync(ray_hits, c->d_ray_hits + b * c->R, (size_t)n * c->R, cudaMemcpyDeviceToHost, c->stream));
    }
    unsigned long long t[2] = {0, 0};
    CU(cudaMemcpyAsync(t, c->d_totals, sizeof t, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    if (totals) { totals[0] = t[0]; totals[1] = t[1]; totals[2] = (uint64_t)c->dbg_traced; }
    return 0;
}

int lumcu_debug_unpack_table(LumcuCtx *c, float *out) {
    if (int rc = use(c)) return rc;
    if (!out) return fail(LUMCU_E_ARG, "null argument");
    float *d = nullptr;
    CU(cudaMalloc(&d, 1024 * sizeof(float)));
    unpack_table_kernel<<<4, 256, 0, c->stream>>>(d);
    c->launches++;
    cudaError_t e = cudaMemcpyAsync(out, d, 1024 * sizeof(float), cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    cudaFree(d);
    if (e != cudaSuccess) return fail(LUMCU_E_CUDA, "unpack table: %s", cudaGetErrorString(e));
    return 0;
}

int lumcu_wait_idle(LumcuCtx *c) {
    if (int rc = use(c)) return rc;
    CU(cudaStreamSynchronize(c->prep_stream));
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaStreamSynchronize(c->dl_stream));
    CU(cudaStreamSynchronize(c->copy_stream));
    return 0;
}

int lumcu_last_update_ms(LumcuCtx *c, float *ms) {
    if (int rc = use(c)) return rc;
    if (!ms) return fail(LUMCU_E_ARG, "null argument");
    if (!c->timed) return fail(LUMCU_E_STATE, "no update has run");
    CU(cudaEventSynchronize(c->ev_end));
    CU(cudaEventElapsedTime(ms, c->ev_begin, c->ev_end));
    return 0;
}

int lumcu_set_kernel_variant(LumcuCtx *c, int32_t variant) {
    if (!c) return fail(LUMCU_E_ARG, "null context");
    if (variant < 0 || variant > 2) return fail(LUMCU_E_ARG, "unknown kernel variant %d", variant);
    c->kernel_variant = variant;
    return 0;
}

int lumcu_set_shard_mode(LumcuCtx *c, int32_t mode) {
    if (!c) return fail(LUMCU_E_ARG, "null context");
    if (mode != LUMCU_SHARD_CONTIGUOUS && mode != LUMCU_SHARD_INTERLEAVED) return fail(LUMCU_E_ARG, "unknown shard mode %d", mode);
    c->shard_chunk = mode == LUMCU_SHARD_INTERLEAVED ? IL_CHUNK : 0;
    return 0;
}

int lumcu_shard_chunk(LumcuCtx *c, int32_t size, int32_t nranks, int32_t *chunk) {
    if (!c || !chunk || size < 1 || nranks < 1) return fail(LUMCU_E_ARG, "bad argument");
    *chunk = shard_of(c, c->shard_basis / size, 0, nranks).chunk;
    return 0;
}

int lumcu_set_profiling(LumcuCtx *c, int32_t on) {
    if (!c) return fail(LUMCU_E_ARG, "null context");
    c->profiling = on != 0;
    return 0;
}

int lumcu_last_kernel_ms(LumcuCtx *c, float *ms) {
    if (int rc = use(c)) return rc;
    if (!ms) return fail(LUMCU_E_ARG, "null argument");
    if (!c->ktimed) return fail(LUMCU_E_STATE, "no profiled update has run (lumcu_set_profiling first)");
    CU(cudaEventSynchronize(c->ev_k1));
    CU(cudaEventElapsedTime(ms, c->ev_k0, c->ev_k1));
    return 0;
}

int lumcu_launch_count(LumcuCtx *c, uint64_t *count) {
    if (!c || !count) return fail(LUMCU_E_ARG, "null argument");
    *count = c->launches;
    return 0;
}

int lumcu_glossy_device(LumcuCtx *c, const LumcuGlossyView *v, int32_t width, int32_t height,
                        uint64_t mat_norm_dev, uint64_t depth_dev, uint64_t frame_dev, uint64_t covered_dev) {
    if (int rc = use(c)) return rc;
    if (!v || width < 1 || height < 1 || (long long)width * height > (1ll << 28))
        return fail(LUMCU_E_ARG, "bad glossy frame (%d x %d)", width, height);
    if (!mat_norm_dev || !depth_dev || !frame_dev) return fail(LUMCU_E_ARG, "null G-buffer / frame");
    if (!c->have_world || !c->have_bricks || !c->have_materials)
        return fail(LUMCU_E_STATE, "lumcu_glossy before world/bricks/materials were uploaded");
    if (int rc = ensure_validated(c)) return rc;
    GlossyArgs a{};
    for (int k = 0; k < 3; k++) {
        a.campos[k] = v->campos[k]; a.camdir[k] = v->camdir[k]; a.hs[k] = v->horizline_scaled[k];
        a.vs[k] = v->vertiline_scaled[k]; a.gld[k] = v->global_light_dir[k];
    }
    a.fw = v->frame_size[0]; a.fh = v->frame_size[1];
    a.width = width; a.height = height;
    a.mat_norm = (const uchar4 *)(uintptr_t)mat_norm_dev;
    a.depth = (const float *)(uintptr_t)depth_dev;
    a.frame = (float4 *)(uintptr_t)frame_dev;
    a.covered = (uint8_t *)(uintptr_t)covered_dev;
    dim3 grid((unsigned)((width + 31) / 32), (unsigned)((height + 3) / 4));
    glossy_kernel<<<grid, 128, 0, c->stream>>>(scene_view(c), a);
    c->launches++;
    CU(cudaGetLastError());
    return 0;
}

int lumcu_glossy(LumcuCtx *c, const LumcuGlossyView *v, int32_t width, int32_t height,
                 const uint8_t *mat_norm, const float *depth, float *frame, uint8_t *covered) {
    if (int rc = use(c)) return rc;
    if (!v || width < 1 || height < 1 || (long long)width * height > (1ll << 28))
        return fail(LUMCU_E_ARG, "bad glossy frame (%d x %d)", width, height);
    if (!mat_norm || !depth || !frame) return fail(LUMCU_E_ARG, "null G-buffer / frame");
    const long long n = (long long)width * height;
    if (c->g_cap < n) {
        CU(cudaStreamSynchronize(c->stream));
        cudaFree(c->d_g_matnorm); cudaFree(c->d_g_depth); cudaFree(c->d_g_frame); cudaFree(c->d_g_cov);
        c->d_g_matnorm = nullptr; c->d_g_depth = nullptr; c->d_g_frame = nullptr; c->d_g_cov = nullptr; c->g_cap = 0;
        CU(cudaMalloc(&c->d_g_matnorm, (size_t)n * sizeof(uchar4)));
        CU(cudaMalloc(&c->d_g_depth, (size_t)n * sizeof(float)));
        CU(cudaMalloc(&c->d_g_frame, (size_t)n * sizeof(float4)));
        CU(cudaMalloc(&c->d_g_cov, (size_t)n));
        c->g_cap = n;
    }
    // caller buffers are read before the call returns: plain synchronous-with-respect-to-host copies
    CU(cudaMemcpyAsync(c->d_g_matnorm, mat_norm, (size_t)n * 4, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(c->d_g_depth, depth, (size_t)n * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(c->d_g_frame, frame, (size_t)n * sizeof(float4), cudaMemcpyHostToDevice, c->stream));
    if (int rc = lumcu_glossy_device(c, v, width, height, (uint64_t)(uintptr_t)c->d_g_matnorm, (uint64_t)(uintptr_t)c->d_g_depth,
                                     (uint64_t)(uintptr_t)c->d_g_frame, (uint64_t)(uintptr_t)c->d_g_cov)) return rc;
    CU(cudaMemcpyAsync(frame, c->d_g_frame, (size_t)n * sizeof(float4), cudaMemcpyDeviceToHost, c->stream));
    if (covered) CU(cudaMemcpyAsync(covered, c->d_g_cov, (size_t)n, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return 0;
}

int lumcu_field_device_ptr(LumcuCtx *c, uint64_t *p) {
    if (!c || !p) return fail(LUMCU_E_ARG, "null argument");
    *p = (uint64_t)(uintptr_t)c->d_field;
    return 0;
}
int lumcu_world_device_ptr(LumcuCtx *c, uint64_t *p) {
    if (!c || !p) return fail(LUMCU_E_ARG, "null argument");
    *p = (uint64_t)(uintptr_t)c->d_world;
    c->world_shared = true;      // the holder may write ids: the march revalidates before every update
    return 0;
}
int lumcu_stream(LumcuCtx *c, uint64_t *s) {
    if (!c || !s) return fail(LUMCU_E_ARG, "null argument");
    *s = (uint64_t)(uintptr_t)c->stream;
    return 0;
}

}  // extern "C"
