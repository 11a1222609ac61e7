// lum_c_api.cpp -- the C99 bindings of include/clum.h over Lum::Renderer
// (reference: src/renderer/api/crenderer.cpp:12-209; same symbols, same argument conversions).
#include "clum.h"
#include "lum_opaque.hpp"
#include "vox.hpp"

#include <cstring>
#include <string>

static inline Lum::Renderer* R(LumRenderer p) { return static_cast<Lum::Renderer*>(p); }
static inline LumInternal::LumInternalRenderer& IR(LumRenderer p) { return R(p)->opaque_members->render; }

extern "C" {

LumRenderer lum_create_instance(const LumSettings* settings) {
    // the reference default-constructs (crenderer.cpp:12-17: 1024 palette slots, 1024 meshes)
    (void)settings;
    return new Lum::Renderer();
}
void lum_destroy_instance(LumRenderer instance) { delete R(instance); }

void lum_demo_load_scene(LumRenderer instance, const char* path) { IR(instance).load_scene(path); }

void lum_init(LumRenderer instance, const LumSettings* s) {
    Lum::Settings st = {};
    st.world_size.x = (float)s->world_size[0];
    st.world_size.y = (float)s->world_size[1];
    st.world_size.z = (float)s->world_size[2];
    st.static_block_palette_size = s->static_block_palette_size;
    st.maxParticleCount = s->maxParticleCount;
    st.timestampCount = s->timestampCount;
    st.fif = s->fif;
    st.vsync = s->vsync; st.fullscreen = s->fullscreen; st.debug = s->debug; st.profile = s->profile;
    R(instance)->init(st);
}
void lum_load_world(LumRenderer instance, const LumMeshBlock* world_data) { R(instance)->loadWorld(world_data); }
void lum_load_world_file(LumRenderer instance, const char* file) { R(instance)->loadWorld(std::string(file)); }

void lum_set_world_block(LumRenderer instance, int x, int y, int z, LumMeshBlock block) { R(instance)->setWorldBlock(x, y, z, block); }
LumMeshBlock lum_get_world_block(const LumRenderer instance, int x, int y, int z) { return R(instance)->getWorldBlock(x, y, z); }

void lum_load_block(LumRenderer instance, LumMeshBlock block, void* block_data) {
    R(instance)->loadBlock(block, reinterpret_cast<LumInternal::BlockVoxels*>(block_data));
}
void lum_load_block_file(LumRenderer instance, LumMeshBlock block, const char* file) { R(instance)->loadBlock(block, std::string(file)); }

void lum_upload_block_palette_to_gpu(LumRenderer instance) { R(instance)->uploadBlockPaletteToGPU(); }
void lum_upload_material_palette_to_gpu(LumRenderer instance) { R(instance)->uploadMaterialPaletteToGPU(); }

LumMeshModel lum_load_mesh_file(LumRenderer instance, const char* file, bool try_extract_palette) {
    return R(instance)->loadModel(std::string(file), try_extract_palette).ptr;
}
LumMeshModel lum_load_mesh_data(LumRenderer instance, void* mesh_data, int x_size, int y_size, int z_size) {
    return R(instance)->loadModel(reinterpret_cast<LumInternal::Voxel*>(mesh_data), x_size, y_size, z_size).ptr;
}
void lum_free_mesh(LumRenderer instance, LumMeshModel model) {
    if (model) R(instance)->freeMesh(Lum::MeshModel(model));
}

LumMeshFoliage lum_load_foliage(LumRenderer instance, const char* vertex_shader_file, int vertices_per_foliage, int density) {
    return R(instance)->loadFoliage(std::string(vertex_shader_file), vertices_per_foliage, density);
}
LumMeshVolumetric lum_load_volumetric(LumRenderer instance, float max_density, float density_deviation, const uint8_t color[3]) {
    return R(instance)->loadVolumetric(max_density, density_deviation, Lum::u8vec3(color[0], color[1], color[2]));
}
LumMeshLiquid lum_load_liquid(LumRenderer instance, int main_mat, int foam_mat) {
    return R(instance)->loadLiquid((LumInternal::MatID_t)main_mat, (LumInternal::MatID_t)foam_mat);
}

void lum_start_frame(LumRenderer instance) { R(instance)->startFrame(); }
void lum_draw_world(LumRenderer instance) { R(instance)->drawWorld(); }
void lum_draw_particles(LumRenderer instance) { R(instance)->drawParticles(); }
void lum_draw_model(LumRenderer instance, const LumMeshModel mesh, const LumMeshTransform* trans) {
    Lum::MeshTransform t = {};
    t.rot = Lum::quat(trans->rot[0], trans->rot[1], trans->rot[2], trans->rot[3]);   // (w, x, y, z)
    t.shift = Lum::vec3(trans->shift[0], trans->shift[1], trans->shift[2]);
    R(instance)->drawModel(Lum::MeshModel(mesh), t);
}
void lum_draw_foliage_block(LumRenderer instance, const LumMeshFoliage foliage, const int pos[3]) {
    R(instance)->drawFoliageBlock(foliage, Lum::ivec3(pos[0], pos[1], pos[2]));
}
void lum_draw_liquid_block(LumRenderer instance, const LumMeshLiquid liquid, const int pos[3]) {
    R(instance)->drawLiquidBlock(liquid, Lum::ivec3(pos[0], pos[1], pos[2]));
}
void lum_draw_volumetric_block(LumRenderer instance, const LumMeshVolumetric volumetric, const int pos[3]) {
    R(instance)->drawVolumetricBlock(volumetric, Lum::ivec3(pos[0], pos[1], pos[2]));
}

void lum_prepare_frame(LumRenderer instance) { R(instance)->prepareFrame(); }
void lum_submit_frame(LumRenderer instance) { R(instance)->submitFrame(); }

void lum_cleanup(LumRenderer instance) {          // also deletes the instance (crenderer.cpp:168-171)
    R(instance)->cleanup();
    delete R(instance);
}
void* lum_get_glfw_ptr(const LumRenderer instance) { return R(instance)->getGLFWptr(); }
bool lum_get_should_close(const LumRenderer instance) { return R(instance)->getShouldClose(); }
bool lum_set_should_close(const LumRenderer instance, bool should_close) { return R(instance)->setShouldClose(should_close); }
void lum_wait_idle(LumRenderer instance) { R(instance)->waitIdle(); }

bool lum_is_profiling_enabled(const LumRenderer instance) { return IR(instance).lumal.settings.profile; }
int lum_get_current_timestamp_count(const LumRenderer instance) { return IR(instance).lumal.currentTimestamp; }
void lum_get_timestamp_name(const LumRenderer instance, int index, char* out_name, size_t name_length) {
    if (!out_name || name_length == 0) return;
    const auto& names = IR(instance).lumal.timestampNames;
    const char* name = (index >= 0 && (size_t)index < names.size() && names[index]) ? names[index] : "";
    strncpy(out_name, name, name_length - 1);
    out_name[name_length - 1] = '\0';
}
double lum_get_timestamp(const LumRenderer instance, int index) {
    const auto& t = IR(instance).lumal.average_ftimestamps;
    return (index >= 0 && (size_t)index < t.size()) ? t[index] : 0.0;
}
double lum_get_timestamp_difference(const LumRenderer instance, int index_start, int index_end) {
    return lum_get_timestamp(instance, index_end) - lum_get_timestamp(instance, index_start);
}

// ---- additions ----
void lum_b200_request_radiance_update(LumRenderer instance, int x, int y, int z) {
    IR(instance).specialRadianceUpdatesWide.push_back(Lum::i16vec4((int16_t)x, (int16_t)y, (int16_t)z, (int16_t)0));
}
void lum_b200_set_camera(LumRenderer instance, const double pos[3], const double dir[3]) {
    auto& c = IR(instance).camera;
    c.cameraPos = Lum::dvec3(pos[0], pos[1], pos[2]);
    c.cameraDir = Lum::dvec3(dir[0], dir[1], dir[2]);
    c.updateCamera();
}
void lum_b200_set_radiance_shift(LumRenderer instance, int dx, int dy, int dz) { R(instance)->setStoredRadianceShift(Lum::ivec3(dx, dy, dz)); }
void lum_b200_set_radiance_controller(LumRenderer instance, bool enabled, int magic_size) {
    IR(instance).radianceControllerEnabled = enabled;
    if (magic_size >= 1) IR(instance).magicSize = magic_size;
}
int lum_b200_get_magic_size(const LumRenderer instance) { return IR(instance).magicSize; }
void lum_b200_set_radiance_time_override(LumRenderer instance, float ms) { IR(instance).radianceTimeOverride = ms; }
long long lum_b200_get_request_count(const LumRenderer instance) { return IR(instance).request_count(); }
int lum_b200_get_frame_index(const LumRenderer instance) { return IR(instance).lumal.iFrame; }
void lum_b200_download_radiance_field(LumRenderer instance, uint32_t* words) { IR(instance).download_radiance_field(words); }
void lum_b200_upload_radiance_field(LumRenderer instance, const uint32_t* words) { IR(instance).upload_radiance_field(words); }
void lum_b200_glossy(LumRenderer instance, int width, int height, const uint8_t* mat_norm, const float* depth, float* frame, uint8_t* covered) {
    IR(instance).glossy(width, height, mat_norm, depth, frame, covered);
}
void lum_b200_download_block_palette(LumRenderer instance, int first, int count, uint8_t* voxels) { IR(instance).download_block_palette(first, count, voxels); }
void lum_b200_download_world(LumRenderer instance, int16_t* ids) { IR(instance).download_world(ids); }
void lum_b200_get_material_palette(const LumRenderer instance, float* materials) {
    memcpy(materials, R(instance)->opaque_members->mat_palette.data(), 256 * 6 * sizeof(float));
}
void lum_b200_set_material_palette(LumRenderer instance, const float* materials) {
    memcpy(static_cast<void*>(R(instance)->opaque_members->mat_palette.data()), materials, 256 * 6 * sizeof(float));
    IR(instance).hasPalette = true;
}
void lum_b200_get_model_size(const LumMeshModel model, int size[3]) {
    const Lum::ivec3 s = Lum::MeshModel(model).getSize();
    size[0] = s.x; size[1] = s.y; size[2] = s.z;
}

int lum_b200_read_vox(const char* path, int size[3], uint8_t* voxels, size_t voxel_cap, float* materials) {
    lumvox::Scene sc;
    std::string err;
    if (!path || !size || !lumvox::read_scene_file(path, sc, err)) return -1;
    size[0] = sc.sx; size[1] = sc.sy; size[2] = sc.sz;
    if (voxels) memcpy(voxels, sc.voxels.data(), sc.voxels.size() < voxel_cap ? sc.voxels.size() : voxel_cap);
    if (materials) lumvox::extract_materials(sc, materials);
    return 0;
}

}  // extern "C"
