// lum_opaque.hpp -- the "transparent" half of Lum::Renderer, headless B200 edition.
//
// The reference keeps its heavy members behind Lum::OpaqueRendererMembers
// (src/renderer/api/opaque_renderer_members.hpp:19-51) and applications reach through it
// (examples/demo.cpp:212-213 pushes into opaque_members->render.specialRadianceUpdates;
// :230-237 reads opaque_members->render.lumal.average_ftimestamps).  This header keeps those
// names.  LumInternalRenderer here holds only what the radiance path needs
// (src/renderer/src/internal_render.hpp:210-458): the CPU world tables, the probe-update queue and
// its controller, and a LumcuCtx (include/lumcu.h) instead of Vulkan objects.
#pragma once
#ifndef __LUM_TRANSPARENT_HPP__
#define __LUM_TRANSPARENT_HPP__

#include <array>
#include <cstdlib>
#include <vector>

#include "lum.hpp"

struct LumcuCtx;

namespace LumInternal {

// internal_render.hpp:121-198 (row-major: x fastest)
template <typename Type>
class table3d {
    std::vector<Type> memory;
    ivec3 _size = {0, 0, 0};
public:
    void allocate(int x, int y, int z) { _size = {x, y, z}; memory.assign((size_t)x * y * z, Type()); }
    void allocate(ivec3 s) { allocate(s.x, s.y, s.z); }
    void deallocate() { memory.clear(); memory.shrink_to_fit(); _size = {0, 0, 0}; }
    void set(Type v) { for (auto& e : memory) e = v; }
    size_t linear_index(int x, int y, int z) const { return (size_t)x + (size_t)_size.x * ((size_t)y + (size_t)_size.y * (size_t)z); }
    bool inside(int x, int y, int z) const { return x >= 0 && x < _size.x && y >= 0 && y < _size.y && z >= 0 && z < _size.z; }
    Type& operator()(int x, int y, int z) { check(x, y, z); return memory[linear_index(x, y, z)]; }
    const Type& operator()(int x, int y, int z) const { check(x, y, z); return memory[linear_index(x, y, z)]; }
    Type& operator()(ivec3 v) { return (*this)(v.x, v.y, v.z); }
    const Type& operator()(ivec3 v) const { return (*this)(v.x, v.y, v.z); }
    Type* data() { return memory.data(); }
    const Type* data() const { return memory.data(); }
    ivec3 size() const { return _size; }
private:
    void check(int x, int y, int z) const { if (!inside(x, y, z)) abort(); }   // reference: ASSERT -> abort
};

// load_stuff.cpp:376-385: a voxel model; `device_model` is the lumcu model handle (the reference
// keeps a ring of R8 3D images here).
struct InternalMeshModel {
    ivec3 size = {0, 0, 0};
    int32_t device_model = -1;
    std::vector<Voxel> voxels;          // x + sx*(y + sy*z)
    bool allocated = false;
};
struct InternalMeshFoliage { std::string vertex_shader_file; int vertices = 0; int dencity = 0; };
struct InternalMeshVolumetric { float max_dencity = 0; float variation = 0; u8vec3 color = {}; };
struct InternalMeshLiquid { MatID_t main = 0; MatID_t foam = 0; };

// internal_render.hpp BlockWithMesh: only the voxels matter here
struct BlockWithMesh { BlockVoxels voxels; };

// the slice of Lumal::Renderer the radiance path and the demo touch (al.hpp:551-602)
struct LumalHeadless {
    struct { bool profile = true; int timestampCount = 48; int fif = 2; } settings;
    int iFrame = 0;                                  // al.cpp:453
    int currentTimestamp = 0;
    std::vector<const char*> timestampNames;
    std::vector<double> ftimestamps;                 // ms since timestamp 0 of the frame
    std::vector<double> average_ftimestamps;         // al.cpp:440-449
};

class LumInternalRenderer {
public:
    // ---- state the reference exposes (internal_render.hpp) ----
    ivec3 world_size = ivec3(48, 48, 16);
    int static_block_palette_size = 15;
    table3d<BlockID_t> origin_world;                 // :388  edited by setWorldBlock
    table3d<BlockID_t> current_world;                // :389  origin + per-frame dynamic ids
    std::vector<i8vec4> specialRadianceUpdates;      // :411  (reference type; coordinates <= 127)
    std::vector<i16vec4> specialRadianceUpdatesWide; // same queue for worlds larger than 127 blocks
    Camera camera;
    LumalHeadless lumal;
    bool hasPalette = false;
    double deltaTime = 0;
    std::vector<Particle> particles;

    // ---- probe-update controller (internal_render.hpp:445-446, internal_render.cpp:1190-1201) ----
    float timeTakenByRadiance = 0.0f;                // ms, measured on the previous frame's dispatch
    int magicSize = 2;
    bool radianceControllerEnabled = true;           // false: magicSize stays as set (tests, benchmarks)
    float radianceTimeOverride = -1.0f;              // >= 0: what end_frame reports as the dispatch time instead of the
                                                     // measured one (makes the controller's trajectory reproducible)

    // ---- B200 stage ----
    LumcuCtx* stage = nullptr;
    int device = 0;                                  // CUDA ordinal (env LUM_B200_DEVICE)
    int rays_per_probe = 64;                         // radiance.comp:37
    float reactiveness = 0.01f;                      // radiance.comp:28
    int64_t request_count();                         // size of the request list of the last frame (asks the stage: one readback)
    bool worldDirty = true;                          // current_world differs from the device copy

    void init(const Lum::Settings& settings);
    void cleanup();

    // load_stuff.cpp
    void load_scene(const char* file);                                                   // :187-219
    void save_scene(const char* file);                                                   // :221-231
    void load_block(BlockWithMesh* block, const char* vox_file);                          // :308-334
    void load_mesh(InternalMeshModel* mesh, const char* vox_file, bool make_vertices, bool extrude_palette, Material* mat_palette);   // :288-305
    void load_mesh(InternalMeshModel* mesh, Voxel* voxels, int sx, int sy, int sz, bool make_vertices);                                // :376-385
    void free_mesh(InternalMeshModel* mesh);                                              // :397-430
    void extract_palette(const char* scene_file, Material* mat_palette);                 // :273-286
    void updateBlockPalette(BlockWithMesh* block_palette);                                // :36-86
    void updateMaterialPalette(Material* material_palette);                               // :89-111

    // per-frame compute sequence (renderer.cpp:422-457)
    void start_frame();                                                                   // internal_render.cpp:957-975
    void start_blockify();                                                                // :977-987
    void blockify_mesh(InternalMeshModel* mesh, MeshTransform* trans);                    // :1029-1076
    void end_blockify();                                                                  // :1092-1097
    void blockify_custom(void* ptr);                                                      // :1099-1103
    void shift_radiance(ivec3 radiance_shift);                                            // :1218-1335
    void update_radiance();                                                               // :1129-1214
    void exec_copies();                                                                   // :1375-1476
    void start_map();                                                                     // :1478-1483
    void map_mesh(InternalMeshModel* mesh, MeshTransform* trans);                         // :1485-1518
    void end_map();                                                                       // :1520-1527
    void end_frame();                                                                     // :2081-2180

    // host-side consumers / tests (the reference never reads these back)
    void download_radiance_field(uint32_t* words);       // world cells, x + X*(y + Y*z)
    void upload_radiance_field(const uint32_t* words);
    // glossy() (internal_render.cpp:2027-2037) as a compute pass over a caller-supplied G-buffer; the uniform block is
    // filled from `camera` and `lightDir` as start_raygen does (:1580-1592).  frame: RGBA f32, written where covered.
    dvec3 lightDir = dvec3(0.4368520327, 0.4368520327, -0.7863336589);   // normalize(0.5, 0.5, -0.9), internal_render.hpp:259
    void glossy(int width, int height, const uint8_t* mat_norm, const float* depth, float* frame, uint8_t* covered);
    void download_block_palette(int first, int count, uint8_t* voxels);
    void download_world(int16_t* ids);
    void wait_idle();                                    // vkDeviceWaitIdle

    struct AABB { vec3 min; vec3 max; };
    AABB get_shift(dmat4 trans, ivec3 size);                                              // :989-1016

private:
    int paletteCounter = 0;                          // internal_render.hpp:443
    std::vector<int32_t> copySrc, copyDst;           // blockCopyQueue (:1046-1065) as id pairs
    void place_timestamp(const char* name, double ms);
    void check(int rc, const char* what);
};

}  // namespace LumInternal

namespace Lum {

// opaque_renderer_members.hpp:19-51
class OpaqueRendererMembers {
public:
    OpaqueRendererMembers() noexcept : OpaqueRendererMembers(1024, 1024, 64, 64, 64) {}
    OpaqueRendererMembers(size_t block_palette_size, size_t mesh_storage_size, size_t foliage_storage_size,
                          size_t volumetric_storage_size, size_t liquid_storage_size) noexcept;
    ~OpaqueRendererMembers() noexcept;

    LumInternal::LumInternalRenderer render;
    Lum::Settings settings;
    // fixed-capacity pools (the reference uses Arena<T>; pointers stay valid for the renderer's life)
    std::vector<LumInternal::InternalMeshModel> mesh_models_storage;
    std::vector<LumInternal::InternalMeshFoliage> mesh_foliage_storage;
    std::vector<LumInternal::InternalMeshVolumetric> mesh_volumetric_storage;
    std::vector<LumInternal::InternalMeshLiquid> mesh_liquid_storage;
    size_t foliage_used = 0, volumetric_used = 0, liquid_used = 0;

    std::vector<LumInternal::BlockWithMesh> block_palette;
    std::array<LumInternal::Material, LumInternal::MATERIAL_PALETTE_SIZE> mat_palette;

    LumInternal::InternalMeshModel* allocate_model();
};

}  // namespace Lum

#endif  // __LUM_TRANSPARENT_HPP__
