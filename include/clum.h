/*
 * clum.h -- C99 API of Lum, headless B200 edition (radiance pass only).
 *
 * Same symbols, handle types and POD layouts as the reference's C bindings
 * (src/renderer/api/crenderer.h:20-114); implemented in lum_b200/liblum.so on top of the C++
 * mirror (include/lum.hpp) and the CUDA stage (include/lumcu.h).  Raster-only calls are accepted
 * and ignored.  The lum_b200_* functions at the end are additions: the reference never reads the
 * radiance field back, a headless consumer has to.
 */
#ifndef LUM_C_BINDINGS_H
#define LUM_C_BINDINGS_H

#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* LumRenderer;
typedef void* LumMeshModel;
typedef int16_t LumMeshBlock;
typedef void* LumMeshFoliage;
typedef void* LumMeshLiquid;
typedef void* LumMeshVolumetric;
typedef void* LumMeshUi;

/* crenderer.h:28-41 */
typedef struct {
    int world_size[3];
    int static_block_palette_size;
    int maxParticleCount;

    int timestampCount;
    int fif;
    char vsync;
    char fullscreen;
    char debug;
    char profile;
} LumSettings;

#define LUM_DEFAULT_SETTINGS ((LumSettings){{48, 48, 16}, 15, 8128, 1024, 2, false, false, false, false})

/* crenderer.h:55-58; rot = (w, x, y, z) (crenderer.cpp:139-145) */
typedef struct LumMeshTransform {
    float rot[4];
    float shift[3];
} LumMeshTransform;
#define LUM_DEFAULT_MESH_TRANSFORM ((LumMeshTransform){{1,0,0,0}, {0,0,0}})

/* -- instance life cycle (crenderer.cpp:12-21, 30-45, 168-171) -- */
LumRenderer lum_create_instance(const LumSettings* settings);
void lum_init(LumRenderer instance, const LumSettings* settings);
void lum_wait_idle(LumRenderer instance);
void lum_cleanup(LumRenderer instance);                       /* also destroys the instance */
void lum_destroy_instance(LumRenderer instance);

/* -- world grid of block ids, x fastest (crenderer.cpp:23-27, 47-62) -- */
void lum_demo_load_scene(LumRenderer instance, const char* path);
void lum_load_world_file(LumRenderer instance, const char* file);
void lum_load_world(LumRenderer instance, const LumMeshBlock* world_data);
LumMeshBlock lum_get_world_block(const LumRenderer instance, int x, int y, int z);
void lum_set_world_block(LumRenderer instance, int x, int y, int z, LumMeshBlock block);

/* -- block palette (16^3 voxel bricks, BlockVoxels[x][y][z]) and material palette (crenderer.cpp:64-80) -- */
void lum_load_block_file(LumRenderer instance, LumMeshBlock block, const char* file);
void lum_load_block(LumRenderer instance, LumMeshBlock block, void* block_data);
void lum_upload_block_palette_to_gpu(LumRenderer instance);
void lum_upload_material_palette_to_gpu(LumRenderer instance);

/* -- voxel models; the first model loaded with try_extract_palette defines the materials (crenderer.cpp:82-99) -- */
LumMeshModel lum_load_mesh_data(LumRenderer instance, void* mesh_data, int x_size, int y_size, int z_size);
LumMeshModel lum_load_mesh_file(LumRenderer instance, const char* file, bool try_extract_palette);
void lum_free_mesh(LumRenderer instance, LumMeshModel model);

/* -- raster-side resources: accepted, recorded, never drawn here (crenderer.cpp:101-121) -- */
LumMeshLiquid lum_load_liquid(LumRenderer instance, int main_mat, int foam_mat);
LumMeshVolumetric lum_load_volumetric(LumRenderer instance, float max_density, float density_deviation, const uint8_t color[3]);
LumMeshFoliage lum_load_foliage(LumRenderer instance, const char* vertex_shader_file, int vertices_per_foliage, int density);

/* -- one frame: start, draw*, prepare, submit (crenderer.cpp:123-166) -- */
void lum_start_frame(LumRenderer instance);
void lum_draw_model(LumRenderer instance, const LumMeshModel mesh, const LumMeshTransform* trans);
void lum_draw_world(LumRenderer instance);
void lum_draw_particles(LumRenderer instance);
void lum_draw_liquid_block(LumRenderer instance, const LumMeshLiquid liquid, const int pos[3]);
void lum_draw_volumetric_block(LumRenderer instance, const LumMeshVolumetric volumetric, const int pos[3]);
void lum_draw_foliage_block(LumRenderer instance, const LumMeshFoliage foliage, const int pos[3]);
void lum_prepare_frame(LumRenderer instance);
void lum_submit_frame(LumRenderer instance);                  /* runs the radiance stage on the GPU */

/* -- window state; headless: no window (crenderer.cpp:173-186) -- */
bool lum_set_should_close(const LumRenderer instance, bool should_close);
bool lum_get_should_close(const LumRenderer instance);
void* lum_get_glfw_ptr(const LumRenderer instance);           /* always NULL */

/* -- GPU timestamps: 0 and 1 bracket the radiance dispatch, milliseconds (crenderer.cpp:188-209) -- */
int lum_get_current_timestamp_count(const LumRenderer instance);
double lum_get_timestamp_difference(const LumRenderer instance, int index_start, int index_end);
double lum_get_timestamp(const LumRenderer instance, int index);
void lum_get_timestamp_name(const LumRenderer instance, int index, char* out_name, size_t name_length);
bool lum_is_profiling_enabled(const LumRenderer instance);

/* ---- additions of the headless B200 edition ---------------------------------------------- */

/* examples/demo.cpp:212-213 pushes into opaque_members->render.specialRadianceUpdates; C has no
 * access to that member, so: queue one API-requested probe update for the next submitted frame. */
void lum_b200_request_radiance_update(LumRenderer instance, int x, int y, int z);
/* Camera of the visibility test in drawModel (renderer.cpp:219-256). */
void lum_b200_set_camera(LumRenderer instance, const double pos[3], const double dir[3]);
/* stored_radiance_shift (renderer.hpp:129-131). */
void lum_b200_set_radiance_shift(LumRenderer instance, int dx, int dy, int dz);
/* Probe-update controller (internal_render.cpp:1190-1201): enabled = reference behaviour; disabled
 * = keep magic_size fixed (magic_size 1 = every probe every frame). */
void lum_b200_set_radiance_controller(LumRenderer instance, bool enabled, int magic_size);
int lum_b200_get_magic_size(const LumRenderer instance);
/* ms >= 0: the controller sees this as the measured time of every dispatch (timeTakenByRadiance,
 * internal_render.cpp:2173) instead of the GPU timestamps; ms < 0: measure (default). */
void lum_b200_set_radiance_time_override(LumRenderer instance, float ms);
long long lum_b200_get_request_count(const LumRenderer instance);
int lum_b200_get_frame_index(const LumRenderer instance);
/* Radiance field, A2B10G10R10 words, index x + X*(y + Y*z): world_size cells. */
void lum_b200_download_radiance_field(LumRenderer instance, uint32_t* words);
void lum_b200_upload_radiance_field(LumRenderer instance, const uint32_t* words);
/* The glossy pass (shaders/glossy.frag; LumInternalRenderer::glossy, internal_render.cpp:2027-2037) over a
 * caller-supplied G-buffer, with the renderer's camera (lum_b200_set_camera) and light direction: mat_norm
 * uint8 [h][w][4] (R8G8B8A8_UINT: material, normal), depth float [h][w] in [0,1]; frame float [h][w][4] receives
 * the fragment colour (before the blend) where covered (uint8 [h][w], may be NULL) is 1.  See lumcu_glossy. */
void lum_b200_glossy(LumRenderer instance, int width, int height, const uint8_t* mat_norm, const float* depth,
                     float* frame, uint8_t* covered);
/* Device block palette (bricks [first, first+count), BlockVoxels layout) and world ids as the
 * radiance kernel sees them after the last submitted frame. */
void lum_b200_download_block_palette(LumRenderer instance, int first, int count, uint8_t* voxels);
void lum_b200_download_world(LumRenderer instance, int16_t* ids);
/* Material palette on the host side (Material[256] as 6 floats each). */
void lum_b200_get_material_palette(const LumRenderer instance, float* materials);
void lum_b200_set_material_palette(LumRenderer instance, const float* materials);
/* Size of a loaded model (MeshModelWithGetters::getSize). */
void lum_b200_get_model_size(const LumMeshModel model, int size[3]);
/* The .vox decoding Lum's loaders rely on (load_stuff.cpp:233-334), without a renderer: models[0]
 * size and voxels (x + sx*(y + sy*z), at most voxel_cap bytes written; pass NULL to query the size)
 * and the 256 extracted materials (6 floats each; may be NULL).  Returns 0, or -1 on a malformed file. */
int lum_b200_read_vox(const char* path, int size[3], uint8_t* voxels, size_t voxel_cap, float* materials);

#ifdef __cplusplus
}
#endif

#endif /* LUM_C_BINDINGS_H */
