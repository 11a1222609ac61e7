/*
 * lumcu.h -- C-ABI of the B200 radiance stage (liblumcu.so).
 *
 * This is the seam a Lum maintainer binds in place of the Vulkan calls of the radiance pass:
 * every entry point cites the reference code it replaces (paths relative to the reference
 * tree, platonvin/lum).  POD types and plain pointers only; no C++/CUDA/torch types cross it.
 *
 * Conventions
 *   - every function returns 0 on success, <0 on error (LUMCU_E_*); lumcu_last_error() gives
 *     the message of the last failure on the calling thread.  (The reference aborts on error,
 *     src/macros + lum-al VK_CHECK; the C++ facade turns a non-zero code into abort().)
 *   - one context = one CUDA device = one host thread at a time (reference: single render
 *     thread, al.cpp:315-317).  All work is enqueued on the context's stream; calls that
 *     return data to the host (download_*, last_update_ms) synchronise that stream.
 *   - host buffers are consumed before the call returns (reference: renderer.cpp:135-138,
 *     load_stuff.cpp:80,105); the caller never sees device memory except through the
 *     explicit *_device_ptr getters (interop / tests / NCCL plumbing).
 *   - there is NO CPU fallback: without a CUDA device lumcu_create fails with LUMCU_E_CUDA.
 *
 * Data layouts (kept from Lum)
 *   world      int16  [z][y][x]      block ids, x fastest        (internal_render.hpp:136-139)
 *   bricks     uint8  [id][x][y][z]  material ids, z fastest     (internal_types.hpp:46)
 *   materials  float  [256][6]       r,g,b,a,emmit,rough         (internal_types.hpp:26-30)
 *   field      uint32 [z][y][x]      A2B10G10R10 unorm           (internal_render.hpp:456)
 *   requests   int16  [n][4]         x,y,z,0  (reference: i8vec4, internal_render.hpp:410;
 *                                     widened so worlds larger than 127 blocks are addressable)
 */
#ifndef LUMCU_H
#define LUMCU_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LUMCU_OK            0
#define LUMCU_E_ARG        -1   /* bad argument                                  */
#define LUMCU_E_CUDA       -2   /* CUDA runtime error (incl. "no device")        */
#define LUMCU_E_STATE      -3   /* call order violated (e.g. update before upload) */

#define LUMCU_DEBUG_PRE_BLEND   1   /* per-workgroup mean radiance before the temporal mix */
#define LUMCU_DEBUG_RAYS        2   /* per-ray marching-sample count and hit flag          */

typedef struct LumcuCtx LumcuCtx;

/* Resource description; replaces the sizes baked into createImages()
 * (src/renderer/src/internal_render.cpp:107-209) and the shader constants
 * (shaders/common/consts.glsl:1, shaders/radiance.comp:27-28,37). */
typedef struct LumcuDesc {
    int32_t world_size[3];             /* blocks; reference default 48,48,16                      */
    int32_t palette_capacity;          /* bricks addressable, 1..4096 (64x64 atlas, internal_types.hpp:16-18) */
    int32_t rays_per_probe;            /* 16,32,64,128,256; reference 64 (radiance.comp:37)       */
    float   reactiveness;              /* REACTIVNESS, reference 0.01 (radiance.comp:28)          */
    int32_t device;                    /* CUDA device ordinal                                     */
    int64_t max_requests;              /* 0 -> world cells + 4096                                 */
} LumcuDesc;

const char *lumcu_last_error(void);
int lumcu_version(void);

/* createImages + setupDescriptors + createPipilines for the radiance pass
 * (internal_render.cpp:128-135,198-207,236-243,588).  The field starts as all-zero words. */
int lumcu_create(const LumcuDesc *desc, LumcuCtx **out);
/* cleanup() of those resources (internal_render.cpp deleteImages / destroy pipes). */
int lumcu_destroy(LumcuCtx *ctx);

/* end_blockify + the world part of exec_copies: stagingWorld -> world image
 * (internal_render.cpp:1092-1103,1375-1476).  world_xyz: world_size x*y*z int16.  The caller's buffer is consumed
 * before the call returns (the reference copies synchronously too, renderer.cpp:135-138). */
int lumcu_upload_world(LumcuCtx *ctx, const int16_t *world_xyz);
/* The same without the wait, for a caller that keeps its world in page-locked memory (cudaHostAlloc /
 * cudaHostRegister) -- the role of the reference's persistently mapped stagingWorld ring
 * (internal_render.cpp:198-207, :1092-1097): the copy is queued behind the frame in flight and the call returns at
 * once; the buffer must stay untouched until lumcu_wait_idle / a download / the next upload from it has returned.
 * Pageable memory takes the synchronous path of lumcu_upload_world. */
int lumcu_upload_world_async(LumcuCtx *ctx, const int16_t *world_xyz_pinned);

/* updateBlockPalette (load_stuff.cpp:36-86) and the per-block brick copies of exec_copies
 * (internal_render.cpp:1402-1446): bricks [first_id, first_id+count) in BlockVoxels layout. */
int lumcu_upload_bricks(LumcuCtx *ctx, int32_t first_id, int32_t count, const uint8_t *voxels);

/* updateMaterialPalette (load_stuff.cpp:89-111): Material[256] as 6 floats each. */
int lumcu_upload_materials(LumcuCtx *ctx, const float *materials);

/* The staging copy + vkCmdCopyBuffer of the request list (internal_render.cpp:1170-1184).
 * If a probe appears more than once, which of its workgroups' results is kept is unspecified
 * (the reference races the same way, internal_render.cpp:1164-1168 keeps duplicate specials). */
int lumcu_set_requests(LumcuCtx *ctx, const int16_t *xyzw, int64_t n);

/* Device-side restatement of the request-list builder (internal_render.cpp:1142-1169): probes
 * whose edge-clamped 3x3x3 neighbourhood of the uploaded world sums > 0, in z,y,x order, then the
 * special requests not already selected, in submission order.
 * out_count == NULL: nothing waits for the GPU -- the count stays in device memory, the kernels of the following
 * updates cut their dispatch down to it, and the host learns it when it first asks (lumcu_request_count,
 * lumcu_download_requests, lumcu_slab_range, an instrumented update).  out_count != NULL: one 16-byte readback. */
int lumcu_build_requests(LumcuCtx *ctx, const int16_t *special_xyzw, int32_t nspecial, int64_t *out_count);
int lumcu_request_count(LumcuCtx *ctx, int64_t *out_count);
int lumcu_download_requests(LumcuCtx *ctx, int16_t *out_xyzw, int64_t cap);

/* Push constants + vkCmdDispatch(N/size) of radiance.comp (internal_render.cpp:1202-1213;
 * shaders/radiance.comp:361-414): workgroup w in [0, N/size) updates request w*size+shift with
 * RNG seed time_seed.  Jacobi semantics: every read of the field sees the state before the call. */
int lumcu_update(LumcuCtx *ctx, int32_t time_seed, int32_t size, int32_t shift);
/* Same, restricted to workgroups [wg_begin, wg_end) of that dispatch (probe sharding across
 * GPUs: each rank runs its slab of the z-sorted list; seeds keep using the global w). */
int lumcu_update_range(LumcuCtx *ctx, int32_t time_seed, int32_t size, int32_t shift,
                       int64_t wg_begin, int64_t wg_end);
/* Workgroup range [wg_begin, wg_end) of the dispatch (size, shift) whose probes have z in
 * [z_begin, z_end), searched over the leading z-sorted part of the request list (everything the
 * builder rule produced; for lumcu_set_requests the longest z-monotonic prefix). */
int lumcu_slab_range(LumcuCtx *ctx, int32_t size, int32_t shift, int32_t z_begin, int32_t z_end,
                     int64_t *wg_begin, int64_t *wg_end);
/* lumcu_update restricted to the probes of one z-slab (the unit of multi-GPU sharding, one slab
 * per rank): the slab's range of the sorted part plus those unsorted tail requests (specials)
 * whose z falls in the slab.  The union over disjoint slabs equals lumcu_update. */
int lumcu_update_slab(LumcuCtx *ctx, int32_t time_seed, int32_t size, int32_t shift,
                      int32_t z_begin, int32_t z_end);

/* ---- multi-GPU exchange over peer memory (one process per GPU; no reference counterpart) --------
 * Instead of "update the shard, then all-gather", the store of radiance.comp:353 IS the exchange: the trace kernel
 * stores each new texel into slot w (its workgroup) of the "inbox" of EVERY rank -- the rank's own staging array and
 * the peers' through IPC-mapped pointers (NVLink P2P) -- straight from its last phase, so the transfer rides under
 * the tracing.  A one-warp kernel then publishes "this rank has arrived" in every rank's flag array, and each rank's
 * commit kernel waits for all flags (inside the kernel) and scatters the whole dispatch from its local inbox into its
 * local field.  One barrier per frame, no collective call, nothing remote on the read side.  The inbox has two
 * halves used alternately, so the next frame's stores cannot overtake a slower rank's commit.
 *   1. every rank: lumcu_ipc_export(ctx, blob)            (LUMCU_IPC_BYTES bytes: inbox + flag array)
 *   2. exchange the blobs out of band (e.g. torch.distributed all_gather)
 *   3. every rank: lumcu_peers_open(ctx, rank, nranks, blobs)   (nranks x LUMCU_IPC_BYTES bytes, rank order)
 *   4. per frame, every rank: lumcu_update_shard_exchange(ctx, seed, size, shift), or
 *      lumcu_update_slab_exchange(ctx, seed, size, shift, z0, z1) with its z-slab
 * All ranks must hold the same world, request list, field and max_requests; results are bit-identical to
 * lumcu_update on one GPU.  A rank that stops participating makes the others' wait time out (~20 s) and their next
 * call fail with LUMCU_E_STATE instead of hanging. */
#define LUMCU_IPC_BYTES 128
#define LUMCU_MAX_PEERS 16
int lumcu_ipc_export(LumcuCtx *ctx, uint8_t *blob);
int lumcu_peers_open(LumcuCtx *ctx, int32_t rank, int32_t nranks, const uint8_t *blobs);
int lumcu_peers_close(LumcuCtx *ctx);
int lumcu_update_slab_exchange(LumcuCtx *ctx, int32_t time_seed, int32_t size, int32_t shift,
                               int32_t z_begin, int32_t z_end);
/* Probe sharding by workgroup index (no z-sorted list needed).  lumcu_set_shard_mode chooses how the dispatch's
 * workgroups are dealt to `nranks` shards:
 *   LUMCU_SHARD_INTERLEAVED (default) the dispatch is cut into chunks of 3072 workgroups (fewer for small dispatches,
 *       lumcu_shard_chunk) and shard r takes chunks
 *       r, r + n, ...: every rank samples every part of the world, so the ranks finish together whatever the scene looks
 *       like (C4 at 8 ranks: all trace kernels within 0.7 % of each other);
 *   LUMCU_SHARD_CONTIGUOUS            shard r takes workgroups [T*r/n, T*(r+1)/n) of the dispatch's T: equal probe
 *       counts, and each rank touches only its own part of world and field (the first rank is 2 % faster for it) --
 *       but the cost per probe varies with height even in the uniform C4 world (ground below, sky above: the middle
 *       ranks take 3 % longer than the first), and the frame waits for the slowest rank.
 * Both need the texel-wise exchange (a shard's texels are not a slab of the field).  lumcu_update_shard: trace + local
 * commit of one shard (the union over all shards, each from the same starting field, equals lumcu_update).
 * lumcu_update_shard_exchange: the per-frame call of step 4 above (rank and nranks as given to lumcu_peers_open). */
#define LUMCU_SHARD_CONTIGUOUS 0
#define LUMCU_SHARD_INTERLEAVED 1
int lumcu_set_shard_mode(LumcuCtx *ctx, int32_t mode);
/* The chunk size the interleaved deal of the current request list's dispatch (size) to nranks shards uses: 3072,
 * halved for smaller dispatches until every rank gets at least 64 chunks (not below 48); 0 in contiguous mode.  It is
 * derived from the list's size as lumcu_set_requests gave it, or from a built list's upper bound (cells + specials) --
 * never from the count learnt later -- so all ranks agree without talking, whichever of them has asked for its count. */
int lumcu_shard_chunk(LumcuCtx *ctx, int32_t size, int32_t nranks, int32_t *chunk);
int lumcu_update_shard(LumcuCtx *ctx, int32_t time_seed, int32_t size, int32_t shift, int32_t rank, int32_t nranks);
int lumcu_update_shard_exchange(LumcuCtx *ctx, int32_t time_seed, int32_t size, int32_t shift);

/* shift_radiance (internal_render.cpp:1218-1335): field[p] = field[p + d] over the overlap box. */
int lumcu_shift(LumcuCtx *ctx, int32_t dx, int32_t dy, int32_t dz);

/* ---- feeders of the pass: blockify copies + map.comp -------------------- */

/* create_RayTrace_VoxelImages of load_mesh (load_stuff.cpp:376-385): device copy of a model's voxel
 * grid, voxels[x + sx*(y + sy*z)] (the .vox order, load_stuff.cpp:296).  *out_model is a handle >= 0. */
int lumcu_upload_model(LumcuCtx *ctx, const uint8_t *voxels, int32_t sx, int32_t sy, int32_t sz, int32_t *out_model);
/* free_mesh (load_stuff.cpp:397-430), voxel image part. */
int lumcu_free_model(LumcuCtx *ctx, int32_t model);

/* The brick part of exec_copies (internal_render.cpp:1402-1446) for the blocks blockify_mesh allocated
 * (:1029-1076): for i in [0,n): bricks[dst_ids[i]] = src_ids[i] == 0 ? all-zero : bricks[src_ids[i]].
 * The copies run concurrently: a destination may be neither a source nor named twice (LUMCU_E_ARG otherwise; the
 * reference copies between two distinct images, and blockify's pairs satisfy this by construction: src < static <= dst).
 * Device-to-device: no brick data crosses PCIe per frame.  (The reference clears the whole atlas and
 * re-copies the static bricks; here the static bricks are never overwritten, and every dynamic brick
 * the world references is rewritten by this call, so the visible state is the same.) */
int lumcu_copy_bricks(LumcuCtx *ctx, const int32_t *src_ids, const int32_t *dst_ids, int32_t n);

/* map_mesh + shaders/map.comp (internal_render.cpp:1485-1527, map.comp:32-67): for every world voxel
 * g in [0, 4*ceil(area/4)) (the dispatch's invocations) at world position (g + 0.5 + shift), fetch the
 * model voxel at trunc(inverse_trans * (pos,1)) and, if it is non-zero and the world block under pos
 * is dynamic (id >= static_block_count; map.comp:28 hard-codes 15), store it into that block's brick.
 * inverse_trans: 16 floats, column-major (glm mat4 memory order), = mat4(inverse(shift*rotate));
 * shift = floor(aabb.min), area = ceil(aabb.max) - floor(aabb.min) (internal_render.cpp:1506-1512).
 * Models mapped by successive calls are applied in call order (the reference records the dispatches
 * without barriers between them; overlapping models race there, here the later one wins). */
int lumcu_map_model(LumcuCtx *ctx, int32_t model, const float inverse_trans[16], const int32_t shift[3],
                    const int32_t area[3], int32_t static_block_count);

/* Readbacks for tests / host-side consumers. */
int lumcu_download_bricks(LumcuCtx *ctx, int32_t first_id, int32_t count, uint8_t *voxels);
int lumcu_download_world(LumcuCtx *ctx, int16_t *world_xyz);

/* Field snapshots (the reference never reads the image back; needed by tests and by any
 * host-side consumer).  field_xyz: world cells uint32. */
int lumcu_upload_field(LumcuCtx *ctx, const uint32_t *field_xyz);
int lumcu_download_field(LumcuCtx *ctx, uint32_t *field_xyz);
/* Words [first, first + count) of the field into page-locked host memory, queued on the stage's stream; returns at
 * once.  The data is there once a fence recorded after this call has been waited for (or after lumcu_wait_idle). */
int lumcu_download_field_async(LumcuCtx *ctx, uint32_t *out_pinned, int64_t first, int64_t count);
/* Frames in flight (the reference keeps fif = 2 frames in flight behind fences, al.cpp:315-317): lumcu_record_fence
 * marks the point the stage's stream has reached in one of LUMCU_FENCES slots, lumcu_wait_fence blocks the host until
 * the GPU has passed it.  Waiting on a slot that was never recorded returns at once. */
#define LUMCU_FENCES 4
int lumcu_record_fence(LumcuCtx *ctx, int32_t slot);
int lumcu_wait_fence(LumcuCtx *ctx, int32_t slot);

/* The 64-entry sphericalFibonacci table the kernel uses (radiance.comp:293-314, evaluated on the
 * host in f32 at create time).  out: rays_per_probe*3 floats. */
int lumcu_direction_table(LumcuCtx *ctx, float *out);

/* Instrumentation (parity tests + roofline numerator).  flags = LUMCU_DEBUG_* bits; while
 * non-zero, lumcu_update runs the instrumented kernel variant and fills the debug buffers. */
int lumcu_set_debug(LumcuCtx *ctx, int32_t flags);
/* Extent of what the last instrumented update left in the debug buffers: *nwg workgroups starting at workgroup
 * *wg_begin of the dispatch (for a shard update: the whole range, of which only the shard's entries were written).
 * lumcu_download_debug copies exactly this many entries: size the output buffers from it. */
int lumcu_debug_extent(LumcuCtx *ctx, int64_t *nwg, int64_t *wg_begin);
/* pre_blend: nwg*3 floats; ray_samples: nwg*R uint32; ray_hits: nwg*R uint8 (nwg from lumcu_debug_extent); totals:
 * S,H,N of the last instrumented update.  Any pointer may be NULL. */
int lumcu_download_debug(LumcuCtx *ctx, float *pre_blend, uint32_t *ray_samples, uint8_t *ray_hits,
                         uint64_t totals[3]);

/* Self-test hook: the device's unorm10 -> float conversion for all 1024 codes (the kernel computes
 * c/1023 with a reciprocal + one FMA correction step; tests compare this table bit-for-bit with
 * IEEE division).  out: 1024 floats. */
int lumcu_debug_unpack_table(LumcuCtx *ctx, float *out);

/* vkDeviceWaitIdle for this stage. */
int lumcu_wait_idle(LumcuCtx *ctx);
/* GPU time of the last update, the pair of timestamps around vkCmdDispatch
 * (internal_render.cpp:1205-1207, :2173 timeTakenByRadiance). */
int lumcu_last_update_ms(LumcuCtx *ctx, float *ms);
/* Kernel selection.  0 = one CTA per workgroup, the dispatch shape of radiance.comp itself (best for demo-sized
 * dispatches, which are latency-bound); 1 = pooled kernel, one CTA per 384 rays with lane-level ray refill (best once
 * the dispatch gives it a wave of CTAs); 2 (default) = choose by dispatch size (pooled from 6,144 workgroups on).
 * All give bit-identical results. */
int lumcu_set_kernel_variant(LumcuCtx *ctx, int32_t variant);

/* Per-kernel timing for roofline reports: while on, an extra event pair brackets the ray-march
 * kernel alone (the commit / exchange are outside it); lumcu_last_kernel_ms returns its duration
 * for the last update.  Off by default (the pair costs ~2 us per update). */
int lumcu_set_profiling(LumcuCtx *ctx, int32_t on);
int lumcu_last_kernel_ms(LumcuCtx *ctx, float *ms);
/* Number of kernels this library launched on the context so far. */
int lumcu_launch_count(LumcuCtx *ctx, uint64_t *count);

/* ---- glossy reflections: shaders/glossy.frag as a CUDA pass -------------------
 * The second consumer of the marcher: one mirror ray per G-buffer pixel through the same world /
 * palette / materials, lit by the radiance field (glossy.frag:119-226 CastRay_fast, :328-360
 * trace_glossy_ray, :439-491 main).  Replaces the fullscreen draw of LumInternalRenderer::glossy()
 * (internal_render.cpp:2027-2037) and the stencil rule of fillStencilGlossy.frag (roughness <= 0.5).
 * LumcuGlossyView mirrors the fields of the uniform block the shader reads (shaders/common/ubo.glsl,
 * filled at internal_render.cpp:1580-1592). */
typedef struct LumcuGlossyView {
    float campos[3];              /* ubo.campos.xyz                                                */
    float camdir[3];              /* ubo.camdir.xyz                                                */
    float horizline_scaled[3];    /* camera.horizline * viewSize.x / 2                             */
    float vertiline_scaled[3];    /* camera.vertiline * viewSize.y / 2                             */
    float global_light_dir[3];    /* ubo.globalLightDir.xyz                                        */
    float frame_size[2];          /* ubo.frame_size (swapchain extent)                             */
} LumcuGlossyView;
/* mat_norm: uint8 [height][width][4] = highresMatNorm, R8G8B8A8_UINT: material id, then the normal as
 * (n+1)/2*255 (internal_render.hpp:455).  depth: float [height][width] = the depth attachment's value in
 * [0,1] (the shader scales it by 1000).  frame: float [height][width][4], the fragment colour BEFORE the
 * blend stage (rgb = light/8, a = 1 - roughness); only covered pixels are written.  covered (may be NULL):
 * uint8 [height][width], 1 where the stencil rule lets the pass run.  Host buffers; the pass reads the
 * field as it is on the stream (after the last update/commit). */
int lumcu_glossy(LumcuCtx *ctx, const LumcuGlossyView *view, int32_t width, int32_t height,
                 const uint8_t *mat_norm, const float *depth, float *frame, uint8_t *covered);
/* Same with DEVICE buffers (interop with a renderer that keeps its G-buffer on the GPU); asynchronous on
 * the context's stream. */
int lumcu_glossy_device(LumcuCtx *ctx, const LumcuGlossyView *view, int32_t width, int32_t height,
                        uint64_t mat_norm_dev, uint64_t depth_dev, uint64_t frame_dev, uint64_t covered_dev);

/* Interop getters: raw device addresses and the stream (cudaStream_t) as integers.  The world address is that of
 * the buffer the next launches read; lumcu_upload_world_async switches to the other of its two buffers, so ask again
 * after it.  Once the world address has been handed out the stage re-validates the world before every update (its
 * holder may have written ids). */
int lumcu_field_device_ptr(LumcuCtx *ctx, uint64_t *ptr);
int lumcu_world_device_ptr(LumcuCtx *ctx, uint64_t *ptr);
int lumcu_stream(LumcuCtx *ctx, uint64_t *stream);

#ifdef __cplusplus
}
#endif
#endif /* LUMCU_H */
