// lum.hpp -- C++ API of Lum, headless B200 edition (radiance pass only).
//
// Mirrors the reference's public API for the path this repository accelerates: same names,
// argument meaning and call order as `Lum::Renderer` (src/renderer/api/renderer.hpp:63-157),
// `Lum::Settings` (:44-54) and the types of src/renderer/src/internal_types.hpp:13-64, so that an
// application written against the reference (examples/demo.cpp) drives the CUDA radiance stage
// unchanged.  What runs behind it is NOT a renderer: submitFrame() executes the compute part of the
// reference's frame that feeds and updates the radiance field -- blockify, brick copies, map,
// shift_radiance, update_radiance (renderer.cpp:422-457) -- on a B200 through the C-ABI of
// include/lumcu.h.  Raster-side calls (drawWorld, drawFoliageBlock, particles, ...) are accepted and
// recorded so application code compiles and runs, and do nothing else.
//
// Error behaviour is the reference's: everything is noexcept; a failing CUDA call prints the
// message and abort()s (reference: VK_CHECK -> exit, macros.hpp:38-50).
#pragma once
#ifndef __LUM_RENDERER_HPP__
#define __LUM_RENDERER_HPP__

#include <chrono>
#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

#include "lum_math.hpp"

namespace LumInternal {
using namespace lum_math;

// internal_types.hpp:13-20
const int BLOCK_SIZE = 16;
const int MATERIAL_PALETTE_SIZE = 256;
const int RCACHE_RAYS_PER_PROBE = 32;
const int BLOCK_PALETTE_SIZE_X = 64;
const int BLOCK_PALETTE_SIZE_Y = 64;
const int BLOCK_PALETTE_SIZE = (BLOCK_PALETTE_SIZE_X * BLOCK_PALETTE_SIZE_Y);
const int MAX_FRAMES_IN_FLIGHT = 2;

// internal_types.hpp:22-30
typedef uint8_t MatID_t;
typedef int16_t BlockID_t;
typedef uint8_t Voxel;
const BlockID_t SPECIAL_BLOCK = 32767;
typedef struct Material {
    vec4 color;      // color AND transparancy
    float emmit;     // emmits same color as reflect
    float rough;
} Material;

typedef struct Particle {
    vec3 pos;
    vec3 vel;
    float lifeTime;
    MatID_t matID;
} Particle;

// internal_types.hpp:41-44
typedef struct MeshTransform {
    quat rot = quat(1.0f, 0.0f, 0.0f, 0.0f);
    vec3 shift = vec3(0.0f);
} MeshTransform;

// internal_types.hpp:46 -- [x][y][z], z fastest
typedef Voxel BlockVoxels[BLOCK_SIZE][BLOCK_SIZE][BLOCK_SIZE];

// internal_types.hpp:54-66 (defaults as the reference's; updateCamera = internal_render.cpp:923-934)
typedef struct Camera {
    dvec3 cameraPos = dvec3(60, 0, 194);
    dvec3 cameraDir = dvec3(0.4303180371, 0.7054394051, -0.5643515241);   // normalize(0.61, 1.0, -0.8)
    dmat4 cameraTransform = {};
    double pixelsInVoxel = 5.0;
    dvec2 originViewSize = dvec2(1920, 1080);
    dvec2 viewSize = dvec2(1920 / 5.0, 1080 / 5.0);
    dvec3 cameraRayDirPlane = {};
    dvec3 horizline = {};
    dvec3 vertiline = {};
    void updateCamera();
} Camera;

struct InternalMeshModel;   // lum_opaque.hpp
}  // namespace LumInternal

namespace Lum {
using namespace lum_math;

class OpaqueRendererMembers;   // lum_opaque.hpp

// renderer.hpp:25-34
class MeshModelWithGetters {
public:
    void* ptr = nullptr;
    MeshModelWithGetters(void* ptr) : ptr(ptr) {}
    MeshModelWithGetters() : ptr(nullptr) {}
    ivec3 getSize() const;
    operator void*();
};

typedef MeshModelWithGetters MeshModel;
typedef int16_t MeshBlock;
typedef void* MeshFoliage;
typedef void* MeshLiquid;
typedef void* MeshVolumetric;
typedef void* MeshUi;
typedef LumInternal::MeshTransform MeshTransform;
typedef LumInternal::Particle Particle;
typedef LumInternal::Camera Camera;

// renderer.hpp:44-54
struct Settings {
    vec3 world_size;
    int static_block_palette_size;
    int maxParticleCount;
    int timestampCount = 48;
    int fif = 2;
    bool vsync = false;
    bool fullscreen = false;
    bool debug = false;
    bool profile = true;
};

// renderer.hpp:63-157
struct Renderer {
    Renderer() noexcept;
    Renderer(size_t block_palette_size, size_t mesh_storage_size, size_t foliage_storage_size,
             size_t volumetric_storage_size, size_t liquid_storage_size) noexcept;
    ~Renderer() noexcept;

    void init(Lum::Settings settings) noexcept;
    void loadWorld(const LumInternal::BlockID_t* world_data) noexcept;
    void loadWorld(const std::string& file) noexcept;

    void setWorldBlock(const ivec3& pos, MeshBlock block) noexcept;
    void setWorldBlock(int x, int y, int z, MeshBlock block) noexcept;

    [[nodiscard]] MeshBlock getWorldBlock(const ivec3& pos) const noexcept;
    [[nodiscard]] MeshBlock getWorldBlock(int x, int y, int z) const noexcept;

    void loadBlock(MeshBlock block, LumInternal::BlockVoxels* block_data) noexcept;
    void loadBlock(MeshBlock block, const std::string& file) noexcept;

    void uploadBlockPaletteToGPU() noexcept;
    void uploadMaterialPaletteToGPU() noexcept;

    void loadPalette(const std::string& file) noexcept;
    [[nodiscard]] MeshModel loadModel(const std::string& file, bool try_extract_palette = true) noexcept;
    [[nodiscard]] MeshModel loadModel(LumInternal::Voxel* mesh_data, int x_size, int y_size, int z_size) noexcept;
    void freeMesh(MeshModel model) noexcept;

    [[nodiscard]] MeshFoliage loadFoliage(const std::string& vertex_shader_file, int vertices_per_foliage, int density) noexcept;
    [[nodiscard]] MeshVolumetric loadVolumetric(float max_density, float density_deviation, const u8vec3& color) noexcept;
    [[nodiscard]] MeshLiquid loadLiquid(LumInternal::MatID_t main_mat, LumInternal::MatID_t foam_mat) noexcept;

    void startFrame() noexcept;
        void drawWorld() noexcept;
        void drawParticles() noexcept;
        void drawModel(const MeshModel& mesh, const MeshTransform& trans) noexcept;
        void drawFoliageBlock(const MeshFoliage& foliage, const ivec3& pos) noexcept;
        void drawLiquidBlock(const MeshLiquid& liquid, const ivec3& pos) noexcept;
        void drawVolumetricBlock(const MeshVolumetric& volumetric, const ivec3& pos) noexcept;
    void prepareFrame() noexcept;
    void submitFrame() noexcept;
    void addParticle(Particle particle);

    void cleanup() noexcept;
    void waitIdle() noexcept;

    [[nodiscard]] void* getGLFWptr() const noexcept;   // headless: always nullptr
    [[nodiscard]] bool getShouldClose() const noexcept { return should_close; }
                  bool setShouldClose(bool sc) noexcept { should_close = sc; return should_close; }
    Camera& getCamera();
    ivec3 getWorldSize();

    ivec3 stored_radiance_shift = {};
    ivec3 getStoredRadianceShift() const noexcept;
    void setStoredRadianceShift(const ivec3& shift) noexcept;

    public:
        bool should_close{false};
        OpaqueRendererMembers* opaque_members = nullptr;

    struct RenderRequest { float cam_dist; };
    struct ModelRenderRequest : RenderRequest {
        MeshModel mesh;
        MeshTransform trans;
    };
    struct BlockyRenderRequest     :       RenderRequest { ivec3 pos; };
    struct BlockRenderRequest      : BlockyRenderRequest { MeshBlock block; };
    struct FoliageRenderRequest    : BlockyRenderRequest { MeshFoliage foliage; };
    struct LiquidRenderRequest     : BlockyRenderRequest { MeshLiquid liquid; };
    struct VolumetricRenderRequest : BlockyRenderRequest { MeshVolumetric volumetric; };

    std::vector<BlockRenderRequest> block_que = {};
    std::vector<ModelRenderRequest> mesh_que = {};
    std::vector<FoliageRenderRequest> foliage_que = {};
    std::vector<LiquidRenderRequest> liquid_que = {};
    std::vector<VolumetricRenderRequest> volumetric_que = {};

    std::chrono::high_resolution_clock::time_point curr_time;
    double delta_time{0};
    void updateTime() noexcept;
};

}  // namespace Lum

#endif  // __LUM_RENDERER_HPP__
