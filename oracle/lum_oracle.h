/*
 * lum_oracle.h -- CPU restatement (ORACLE) of Lum's radiance-probe update path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may build, load or call it, and there only as the checker / CPU baseline.
 *
 * HOW PARITY IS PINNED.  The reference (platonvin/lum) has no tests, golden vectors or
 * fixtures for this path, and no Vulkan implementation exists in the build container or on
 * the GPU box (no loader / ICD / lavapipe, no SPIR-V executor).  What does run here is the
 * reference's own SHADER TEXT: oracle/Makefile reads shaders/radiance.comp and shaders/map.comp
 * where they lie under /root/reference, token-rewrites them (oracle/glsl_prep.py: qualifiers,
 * float-literal suffixes, swizzle syntax -- no statement changes) into oracle/_ref/ and compiles
 * them as C++ on a small GLSL environment (oracle/glsl_shim.hpp).  This restatement is checked
 * against live runs of that build (tests/test_ref_glsl.py) and against golden vectors generated
 * from it (tests/golden/ref_glsl_golden.npz, generator committed): packed field words and
 * per-invocation sample counts equal, per-probe radiance within 1e-6 relative (observed 1.7e-7,
 * from the texture-coordinate un-normalisation only; bit-equal on a zero field).
 * What stays a DECREE (DESIGN.md "Oracle semantics"), because a Vulkan driver and not the
 * shader text defines it: the filter arithmetic of textureLod, the order of subgroupAdd,
 * out-of-range fetches, FMA contraction, the precision of normalize/sqrt/division, Jacobi
 * field update.  The shim and this file follow the same decrees.  The host-side C++ of the
 * reference on the path (request-list rule, get_shift, blockify_mesh, map_mesh's arithmetic) is
 * pinned the same way: cut out of internal_render.cpp verbatim and compiled in place
 * (oracle/host_prep.py, oracle/ref_host.cpp -> oracle/_ref/libref_host.so), live and through
 * tests/golden/ref_host_golden.npz (tests/test_ref_host.py): equal, floats bit for bit.
 *
 * All file:line citations are relative to the reference tree (/root/reference).
 */
#ifndef LUM_ORACLE_H
#define LUM_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* Scene inputs in Lum's own layouts. */
typedef struct {
    int wx, wy, wz;          /* world size in blocks (shaders/common/consts.glsl:1)        */
    int palette_capacity;    /* bricks addressable (<= 4096 = 64x64 atlas)                  */
    int rays_per_probe;      /* RAYS_PER_PROBE (radiance.comp:37); reference = 64           */
    float reactiveness;      /* REACTIVNESS (radiance.comp:28) = 0.01                       */
    const int16_t *world;    /* [z][y][x] block ids, x fastest (internal_render.hpp:136)   */
    const uint8_t *bricks;   /* [id][x][y][z] material ids, z fastest (internal_types:46)  */
    const float   *materials;/* [256][6] r,g,b,a,emmit,rough (internal_types.hpp:26-30)    */
} LumOracleScene;

/* Per-ray trace record (debug / roofline numerator). */
typedef struct {
    float light[3];      /* trace_ray result (radiance.comp:316-347)              */
    float origin[3];
    float dir[3];
    float fraction;      /* final fraction (refined on hit, coarse on miss)       */
    float normal[3];
    int32_t hit;         /* 1 = CastRay_fast returned true                        */
    int32_t coarse_samples; /* iterations of the loop at radiance.comp:205-244    */
    int32_t refine_samples; /* GetVoxel calls at radiance.comp:267,280            */
    int32_t voxel;       /* material id used by GetMat on hit                     */
} LumOracleRay;

/* radiance.comp:51-59 (integer part) and :58 (float scale). */
void lum_oracle_hash33(const uint32_t in[3], uint32_t out_u[3], float out_f[3]);

/* radiance.comp:293-314, evaluated in f32 on the host; out = [n][3]. */
void lum_oracle_direction_table(int n, float *out);

/* normalize(vec3(0.5,0.5,-0.9)), radiance.comp:42. */
void lum_oracle_sun_dir(float out[3]);

/* A2B10G10R10 unorm pack/unpack (internal_render.hpp:456; Vulkan unorm rules). */
uint32_t lum_oracle_pack_rgb10a2(const float rgb[3]);
void lum_oracle_unpack_rgb10a2(uint32_t word, float rgb[3]);

/* GetVoxel (radiance.comp:121-143): out[0]=voxel, out[1]=block id at integer voxel position. */
void lum_oracle_get_voxel(const LumOracleScene *s, int x, int y, int z, int out[2]);

/* sample_radiance gather (radiance.comp:161-166 + sampler setup.cpp:75-81) at block-space bp. */
void lum_oracle_gather(const LumOracleScene *s, const uint32_t *field, const float bp[3], float rgb[3]);

/* shaders/glossy.frag as a CPU pass.  View = the uniform-block fields the shader reads
 * (shaders/common/ubo.glsl).  mat_norm uint8 [h][w][4] (R8G8B8A8_UINT: material, normal), depth float [h][w]
 * in [0,1]; frame float [h][w][4] = fragment colour before the blend (written where covered);
 * covered / samples (per-pixel GetVoxel calls of the mirror ray) / coarse (iterations of its coarse loop)
 * may be NULL. */
typedef struct {
    float campos[3], camdir[3], horizline_scaled[3], vertiline_scaled[3], global_light_dir[3];
    float frame_size[2];
} LumOracleGlossyView;
void lum_oracle_glossy(const LumOracleScene *s, const uint32_t *field, const LumOracleGlossyView *v,
                       int width, int height, const uint8_t *mat_norm, const float *depth,
                       float *frame, uint8_t *covered, uint32_t *samples, uint32_t *coarse, int nthreads);
/* Test-input fabrication (not a restatement): G-buffer of an orthographic view of the scene. */
void lum_oracle_gbuffer(const LumOracleScene *s, const LumOracleGlossyView *v, int width, int height,
                        uint8_t *mat_norm, float *depth);

/* update_radiance request list (internal_render.cpp:1142-1169).
 * out: int16 x,y,z,w per request; returns the count (may exceed cap: nothing is written past cap). */
int64_t lum_oracle_build_requests(int wx, int wy, int wz, const int16_t *world,
                                  const int16_t *special_xyzw, int nspecial,
                                  int16_t *out_xyzw, int64_t cap);

/* One ray of one workgroup: ray setup (radiance.comp:370-381) + trace_ray (:316-347). */
void lum_oracle_trace(const LumOracleScene *s, const uint32_t *field,
                      const float *dir_table, const int16_t probe[3],
                      uint32_t global_invocation_x, int ray_index, int time_seed,
                      LumOracleRay *out);

/* Full dispatch: workgroups w in [0, nreq/size) process request w*size+shift
 * (radiance.comp:361-414, internal_render.cpp:1202-1206).
 * field_old is read (gather + old texel), field_new receives a copy of field_old
 * with the processed probes overwritten (Jacobi semantics, decree 1).
 * Optional outputs (may be NULL): pre_blend [nwg][3] mean radiance before the
 * temporal mix; ray_samples [nwg*R] marching samples per ray; ray_hits [nwg*R].
 * totals[0]=S (samples), totals[1]=H (hit rays), totals[2]=N (probes).
 * nthreads <= 0 -> omp_get_max_threads().  Returns the thread count used. */
int lum_oracle_update(const LumOracleScene *s, const uint32_t *field_old, uint32_t *field_new,
                      const int16_t *requests_xyzw, int64_t nreq,
                      int time_seed, int size, int shift,
                      float *pre_blend, uint32_t *ray_samples, uint8_t *ray_hits,
                      uint64_t totals[3], int nthreads);

/* Same, with one more optional output, ray_coarse [nwg*R] = iterations of the coarse loop per ray
 * (used to study warp scheduling policies offline), and an optional sub-range [wg_begin, wg_end)
 * of the dispatch's workgroups (wg_end < 0 = all): only those probes are updated, with their
 * global workgroup indices (what one rank of a probe-sharded update computes).  Output arrays
 * stay indexed by the global workgroup index. */
int lum_oracle_update_ex(const LumOracleScene *s, const uint32_t *field_old, uint32_t *field_new,
                         const int16_t *requests_xyzw, int64_t nreq,
                         int time_seed, int size, int shift,
                         float *pre_blend, uint32_t *ray_samples, uint8_t *ray_hits, uint16_t *ray_coarse,
                         uint64_t totals[3], int nthreads, int64_t wg_begin, int64_t wg_end);

/* shift_radiance (internal_render.cpp:1218-1335): out[p] = in[p + shift] where p + shift is
 * inside the world, else unchanged in[p] (the reference copies only the overlapping box via a
 * temporary image).  |shift| >= world size on any axis is a no-op (the reference tests x only,
 * :1230-1235, and is undefined for y/z; pinned by decree). */
void lum_oracle_shift_field(int wx, int wy, int wz, const uint32_t *in, uint32_t *out,
                            int dx, int dy, int dz);

/* ---- feeders of the pass (lum_oracle_feeders.c) ------------------------------------------------ */

/* What blockify_mesh and map_mesh derive from a MeshTransform + model size
 * (internal_render.cpp:989-1016,1031-1041,1499-1513). */
typedef struct {
    float aabb_min[3], aabb_max[3];     /* get_shift: transformed AABB of box = size-1 (float)          */
    int32_t block_min[3], block_max[3]; /* blockify border in blocks, clamped to the world              */
    int32_t map_shift[3], map_area[3];  /* map_mesh: floor(min), ceil(max) - floor(min)                 */
    float inverse_trans[16];            /* mat4(inverse(shift*rotate)), column-major                    */
} LumOraclePlacement;

/* rot = (w,x,y,z) as the C API passes it (crenderer.cpp:139-145). */
void lum_oracle_place_model(const float rot_wxyz[4], const float shift[3], const int32_t size[3],
                            const int32_t world_size[3], LumOraclePlacement *out);

/* start_blockify + blockify_mesh per model in draw order (internal_render.cpp:977-987,1029-1076). */
int32_t lum_oracle_blockify(int wx, int wy, int wz, const int16_t *origin_world, int16_t *current_world,
                            int static_count, const LumOraclePlacement *models, int nmodels,
                            int32_t *copy_src, int32_t *copy_dst, int32_t copy_cap);

/* Per-block brick copies of exec_copies (internal_render.cpp:1402-1446); src 0 = zero the brick. */
void lum_oracle_copy_bricks(uint8_t *bricks, int capacity, const int32_t *src, const int32_t *dst, int32_t n);

/* shaders/map.comp:32-67 for one map_mesh dispatch (internal_render.cpp:1485-1518). */
void lum_oracle_map_model(int wx, int wy, int wz, const int16_t *world, uint8_t *bricks, int capacity,
                          const uint8_t *model, int sx, int sy, int sz,
                          const float inverse_trans[16], const int32_t shift[3], const int32_t area[3],
                          int static_count);

int lum_oracle_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif
