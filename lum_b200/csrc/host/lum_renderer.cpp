// lum_renderer.cpp -- Lum::Renderer, headless B200 edition (include/lum.hpp).
//
// Same entry points and call order as src/renderer/api/renderer.cpp of the reference; what differs
// is what submitFrame() runs: only the compute steps that feed and update the radiance field
// (renderer.cpp:430-457), on the CUDA stage.
#include "lum_opaque.hpp"
#include "host_math.hpp"

#include <algorithm>
#include <cstdio>
#include <cstring>

using namespace lumhost;

namespace Lum {

OpaqueRendererMembers::OpaqueRendererMembers(size_t block_palette_size, size_t mesh_storage_size, size_t foliage_storage_size,
                                             size_t volumetric_storage_size, size_t liquid_storage_size) noexcept
    : mesh_models_storage(mesh_storage_size), mesh_foliage_storage(foliage_storage_size),
      mesh_volumetric_storage(volumetric_storage_size), mesh_liquid_storage(liquid_storage_size),
      block_palette(block_palette_size) {
    for (auto& b : block_palette) memset(b.voxels, 0, sizeof b.voxels);
    memset(static_cast<void*>(mat_palette.data()), 0, sizeof(LumInternal::Material) * mat_palette.size());
}
OpaqueRendererMembers::~OpaqueRendererMembers() noexcept {}

LumInternal::InternalMeshModel* OpaqueRendererMembers::allocate_model() {
    for (auto& m : mesh_models_storage)
        if (!m.allocated) { m = LumInternal::InternalMeshModel(); m.allocated = true; return &m; }
    fprintf(stderr, "lum-b200: mesh storage exhausted (%zu models)\n", mesh_models_storage.size());
    abort();
}

ivec3 MeshModelWithGetters::getSize() const { return static_cast<LumInternal::InternalMeshModel*>(ptr)->size; }
MeshModelWithGetters::operator void*() { return ptr; }

// renderer.cpp:36-50
Renderer::Renderer() noexcept : Renderer(1024, 1024, 64, 64, 64) {}
Renderer::Renderer(size_t block_palette_size, size_t mesh_storage_size, size_t foliage_storage_size,
                   size_t volumetric_storage_size, size_t liquid_storage_size) noexcept {
    opaque_members = new OpaqueRendererMembers(block_palette_size, mesh_storage_size, foliage_storage_size,
                                               volumetric_storage_size, liquid_storage_size);
    curr_time = std::chrono::high_resolution_clock::now();
}
Renderer::~Renderer() noexcept {
    if (opaque_members) {
        if (opaque_members->render.stage) opaque_members->render.cleanup();
        delete opaque_members;
        opaque_members = nullptr;
    }
}

void Renderer::init(Lum::Settings settings) noexcept {                   // renderer.cpp:52-104
    opaque_members->settings = settings;
    // the reference forces fif = 2 and profile = true (:71-73); fif has no meaning without a swapchain
    settings.fif = 2;
    settings.profile = true;
    if ((size_t)settings.static_block_palette_size > opaque_members->block_palette.size()) {
        fprintf(stderr, "lum-b200: static_block_palette_size %d exceeds the block palette storage %zu\n",
                settings.static_block_palette_size, opaque_members->block_palette.size());
        abort();
    }
    opaque_members->render.init(settings);
    should_close = false;
}

void Renderer::loadWorld(const LumInternal::BlockID_t* world_data) noexcept {
    // The reference ignores world_data and loads "assets/scene" (renderer.cpp:106-110, marked TODO);
    // here the pointer is honoured: world_size.x*y*z ids, x fastest.
    if (!world_data) abort();
    auto& r = opaque_members->render;
    const ivec3 ws = r.world_size;
    for (int z = 0; z < ws.z; z++) for (int y = 0; y < ws.y; y++) for (int x = 0; x < ws.x; x++)
        r.origin_world(x, y, z) = world_data[(size_t)x + (size_t)ws.x * ((size_t)y + (size_t)ws.y * (size_t)z)];
    r.worldDirty = true;
}
void Renderer::loadWorld(const std::string& file) noexcept { opaque_members->render.load_scene(file.c_str()); }

void Renderer::setWorldBlock(const ivec3& pos, MeshBlock block) noexcept {
    opaque_members->render.origin_world(pos) = block;
    opaque_members->render.worldDirty = true;
}
void Renderer::setWorldBlock(int x, int y, int z, MeshBlock block) noexcept { setWorldBlock(ivec3(x, y, z), block); }
MeshBlock Renderer::getWorldBlock(const ivec3& pos) const noexcept { return opaque_members->render.origin_world(pos); }
MeshBlock Renderer::getWorldBlock(int x, int y, int z) const noexcept { return opaque_members->render.origin_world(x, y, z); }

void Renderer::loadBlock(MeshBlock block, LumInternal::BlockVoxels* block_data) noexcept {
    if (block < 0 || (size_t)block >= opaque_members->block_palette.size()) abort();
    memcpy(opaque_members->block_palette[block].voxels, block_data, sizeof(LumInternal::BlockVoxels));
}
void Renderer::loadBlock(MeshBlock block, const std::string& file) noexcept {
    if (block < 0 || (size_t)block >= opaque_members->block_palette.size()) abort();
    opaque_members->render.load_block(&opaque_members->block_palette[block], file.c_str());
}

void Renderer::uploadBlockPaletteToGPU() noexcept { opaque_members->render.updateBlockPalette(opaque_members->block_palette.data()); }
void Renderer::uploadMaterialPaletteToGPU() noexcept { opaque_members->render.updateMaterialPalette(opaque_members->mat_palette.data()); }
void Renderer::loadPalette(const std::string& file) noexcept { opaque_members->render.extract_palette(file.c_str(), opaque_members->mat_palette.data()); }

MeshModel Renderer::loadModel(const std::string& file, bool try_extract_palette) noexcept {
    LumInternal::InternalMeshModel* m = opaque_members->allocate_model();
    opaque_members->render.load_mesh(m, file.c_str(), true, try_extract_palette, opaque_members->mat_palette.data());
    return m;
}
MeshModel Renderer::loadModel(LumInternal::Voxel* mesh_data, int x_size, int y_size, int z_size) noexcept {
    LumInternal::InternalMeshModel* m = opaque_members->allocate_model();
    opaque_members->render.load_mesh(m, mesh_data, x_size, y_size, z_size, true);
    return m;
}
void Renderer::freeMesh(MeshModel model) noexcept {
    auto* m = static_cast<LumInternal::InternalMeshModel*>(model.ptr);
    opaque_members->render.free_mesh(m);
    m->allocated = false;
}

// raster-side resources: recorded, never drawn
MeshFoliage Renderer::loadFoliage(const std::string& vertex_shader_file, int vertices_per_foliage, int density) noexcept {
    if (opaque_members->foliage_used >= opaque_members->mesh_foliage_storage.size()) abort();
    auto* f = &opaque_members->mesh_foliage_storage[opaque_members->foliage_used++];
    f->vertex_shader_file = vertex_shader_file; f->vertices = vertices_per_foliage; f->dencity = density;
    return f;
}
MeshVolumetric Renderer::loadVolumetric(float max_density, float density_deviation, const u8vec3& color) noexcept {
    if (opaque_members->volumetric_used >= opaque_members->mesh_volumetric_storage.size()) abort();
    auto* v = &opaque_members->mesh_volumetric_storage[opaque_members->volumetric_used++];
    v->max_dencity = max_density; v->variation = density_deviation; v->color = color;
    return v;
}
MeshLiquid Renderer::loadLiquid(LumInternal::MatID_t main_mat, LumInternal::MatID_t foam_mat) noexcept {
    if (opaque_members->liquid_used >= opaque_members->mesh_liquid_storage.size()) abort();
    auto* l = &opaque_members->mesh_liquid_storage[opaque_members->liquid_used++];
    l->main = main_mat; l->foam = foam_mat;
    return l;
}

void Renderer::startFrame() noexcept {                                   // renderer.cpp:367-373
    block_que.clear(); mesh_que.clear(); foliage_que.clear(); liquid_que.clear(); volumetric_que.clear();
}

// renderer.cpp:219-256: any corner of the transformed bounding box inside the NDC cube
static bool is_model_visible(const MeshModel& model, const MeshTransform& trans, const dmat4& view_proj_d) {
    float m[4][4];
    for (int c = 0; c < 4; c++) for (int r = 0; r < 4; r++) m[c][r] = (float)view_proj_d[c][r];
    const ivec3 sz = model.getSize();
    const vec3 mx((float)sz.x, (float)sz.y, (float)sz.z);
    for (int x = 0; x <= 1; x++) for (int y = 0; y <= 1; y++) for (int z = 0; z <= 1; z++) {
        const vec3 corner((float)x * mx.x, (float)y * mx.y, (float)z * mx.z);
        const vec3 r = rotate(trans.rot, corner);
        const vec3 p(r.x + trans.shift.x, r.y + trans.shift.y, r.z + trans.shift.z);
        float clip[4];
        for (int k = 0; k < 4; k++) clip[k] = (m[0][k] * p.x + m[1][k] * p.y) + (m[2][k] * p.z + m[3][k] * 1.0f);
        if (clip[3] != 0.0f) { clip[0] /= clip[3]; clip[1] /= clip[3]; clip[2] /= clip[3]; }
        if (clip[0] >= -1.0f && clip[1] >= -1.0f && clip[2] >= -1.0f && clip[0] <= 1.0f && clip[1] <= 1.0f && clip[2] <= 1.0f) return true;
    }
    return false;
}

void Renderer::drawWorld() noexcept {}                                   // raster: block meshes (renderer.cpp:258-292)
void Renderer::drawParticles() noexcept {}

void Renderer::drawModel(const MeshModel& mesh, const MeshTransform& trans) noexcept {   // renderer.cpp:329-342
    if (mesh.ptr == nullptr) abort();
    if (is_model_visible(mesh, trans, opaque_members->render.camera.cameraTransform)) {
        ModelRenderRequest mrr = {};
        mrr.mesh = mesh;
        mrr.trans = trans;
        mesh_que.push_back(mrr);
    }
}
void Renderer::drawFoliageBlock(const MeshFoliage& foliage, const ivec3& pos) noexcept {
    FoliageRenderRequest r = {}; r.foliage = foliage; r.pos = pos; foliage_que.push_back(r);
}
void Renderer::drawLiquidBlock(const MeshLiquid& liquid, const ivec3& pos) noexcept {
    LiquidRenderRequest r = {}; r.liquid = liquid; r.pos = pos; liquid_que.push_back(r);
}
void Renderer::drawVolumetricBlock(const MeshVolumetric& volumetric, const ivec3& pos) noexcept {
    VolumetricRenderRequest r = {}; r.volumetric = volumetric; r.pos = pos; volumetric_que.push_back(r);
}

void Renderer::prepareFrame() noexcept {                                 // renderer.cpp:395-420
    opaque_members->render.deltaTime = delta_time;
    const dmat4& cam = opaque_members->render.camera.cameraTransform;
    for (auto& mrr : mesh_que) {
        const dvec4 clip = mul(cam, dvec4((double)mrr.trans.shift.x, (double)mrr.trans.shift.y, (double)mrr.trans.shift.z, 1.0));
        mrr.cam_dist = (float)(-clip.z);
    }
    // blockify order = draw order after this sort (far to near); stable so equal depths keep call order
    std::stable_sort(mesh_que.begin(), mesh_que.end(),
                     [](const ModelRenderRequest& a, const ModelRenderRequest& b) { return a.cam_dist > b.cam_dist; });
}

void Renderer::submitFrame() noexcept {                                  // renderer.cpp:422-549
    auto& r = opaque_members->render;
    r.start_frame();
    r.start_blockify();
    for (auto& m : mesh_que) r.blockify_mesh(static_cast<LumInternal::InternalMeshModel*>(m.mesh.ptr), &m.trans);
    r.end_blockify();
    // Snapshot semantics (decree 9): the brick copies and the model mapping of
    // THIS frame run before its radiance update, so the update sees the world the request list was
    // built from.  The reference records update_radiance before exec_copies/map (renderer.cpp:438 vs
    // :447-455) and, through its never-rotating descriptor set, reads a world up to two frames old.
    r.exec_copies();
    r.start_map();
    for (auto& m : mesh_que) r.map_mesh(static_cast<LumInternal::InternalMeshModel*>(m.mesh.ptr), &m.trans);
    r.end_map();
    r.shift_radiance(stored_radiance_shift);
    r.update_radiance();
    updateTime();
    r.end_frame();
}

void Renderer::addParticle(Particle particle) { opaque_members->render.particles.push_back(particle); }

void Renderer::cleanup() noexcept {                                      // renderer.cpp:551-580
    waitIdle();
    for (auto& m : opaque_members->mesh_models_storage)
        if (m.allocated) { opaque_members->render.free_mesh(&m); m.allocated = false; }
    opaque_members->render.cleanup();
}

void Renderer::waitIdle() noexcept { opaque_members->render.wait_idle(); }

void* Renderer::getGLFWptr() const noexcept { return nullptr; }

void Renderer::updateTime() noexcept {                                   // renderer.cpp:589-595
    auto now = std::chrono::high_resolution_clock::now();
    delta_time = std::chrono::duration<double>(now - curr_time).count();
    curr_time = now;
}

Camera& Renderer::getCamera() { return opaque_members->render.camera; }
ivec3 Renderer::getWorldSize() { return opaque_members->render.world_size; }
ivec3 Renderer::getStoredRadianceShift() const noexcept { return stored_radiance_shift; }
void Renderer::setStoredRadianceShift(const ivec3& shift) noexcept { stored_radiance_shift = shift; }

}  // namespace Lum
