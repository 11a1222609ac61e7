// lum_internal.cpp -- LumInternal::LumInternalRenderer, headless B200 edition (include/lum_opaque.hpp).
//
// Host side of the radiance path as the reference structures it
// (src/renderer/src/internal_render.cpp, load_stuff.cpp), with the Vulkan half of every function
// replaced by a call into the C-ABI of the CUDA stage (include/lumcu.h).  C++ only: no CUDA or
// torch types here; this file links against liblumcu.so.
#include "lum_opaque.hpp"
#include "lumcu.h"
#include "host_math.hpp"
#include "vox.hpp"

#include <cmath>
#include <cstdio>
#include <cstring>
#include <limits>

namespace LumInternal {

using namespace lumhost;

void LumInternalRenderer::check(int rc, const char* what) {
    if (rc != 0) {   // reference: VK_CHECK prints and exits (lum-al defines/macros.hpp:38-50)
        fprintf(stderr, "lum-b200: %s failed (%d): %s\n", what, rc, lumcu_last_error());
        abort();
    }
}

// ---- init / cleanup (internal_render.cpp:60-105, createImages :107-209) -------------------
void LumInternalRenderer::init(const Lum::Settings& s) {
    world_size = ivec3((int)s.world_size.x, (int)s.world_size.y, (int)s.world_size.z);
    static_block_palette_size = s.static_block_palette_size;
    lumal.settings.profile = s.profile;
    lumal.settings.timestampCount = s.timestampCount;
    lumal.settings.fif = s.fif;
    lumal.timestampNames.assign((size_t)s.timestampCount, "");
    lumal.ftimestamps.assign((size_t)s.timestampCount, 0.0);
    lumal.average_ftimestamps.assign((size_t)s.timestampCount, 0.0);
    origin_world.allocate(world_size);                                   // :69-70
    current_world.allocate(world_size);
    if (const char* e = getenv("LUM_B200_DEVICE")) device = atoi(e);
    if (const char* e = getenv("LUM_B200_RAYS_PER_PROBE")) rays_per_probe = atoi(e);
    LumcuDesc d{};
    d.world_size[0] = world_size.x; d.world_size[1] = world_size.y; d.world_size[2] = world_size.z;
    d.palette_capacity = BLOCK_PALETTE_SIZE;                             // 64x64 atlas (:137-145)
    d.rays_per_probe = rays_per_probe;
    d.reactiveness = reactiveness;
    d.device = device;
    d.max_requests = 0;
    check(lumcu_create(&d, &stage), "lumcu_create");
    camera.updateCamera();
    worldDirty = true;
}

void LumInternalRenderer::cleanup() {
    if (stage) { lumcu_destroy(stage); stage = nullptr; }
    origin_world.deallocate();
    current_world.deallocate();
}

// ---- loaders (load_stuff.cpp) ---------------------------------------------------------------
static std::vector<char> read_file(const char* path) {
    std::vector<char> buf;
    FILE* f = fopen(path, "rb");
    if (!f) return buf;
    char tmp[65536];
    size_t n;
    while ((n = fread(tmp, 1, sizeof tmp, f)) > 0) buf.insert(buf.end(), tmp, tmp + n);
    fclose(f);
    return buf;
}

void LumInternalRenderer::load_scene(const char* file) {                 // :187-219
    std::vector<char> buffer = read_file(file);
    if (buffer.size() < 3 * sizeof(int32_t)) { origin_world.set(0); worldDirty = true; return; }
    int32_t stored[3];
    memcpy(stored, buffer.data(), sizeof stored);
    const size_t count = (buffer.size() - sizeof stored) / sizeof(BlockID_t);
    const char* base = buffer.data() + sizeof stored;
    ivec3 rd = ivec3(clampi(stored[0], 0, world_size.x), clampi(stored[1], 0, world_size.y), clampi(stored[2], 0, world_size.z));
    for (int zz = 0; zz < rd.z; zz++)
    for (int yy = 0; yy < rd.y; yy++)
    for (int xx = 0; xx < rd.x; xx++) {
        size_t index = (size_t)xx + (size_t)stored[0] * yy + ((size_t)stored[0] * stored[1]) * zz;
        if (index >= count) abort();                                     // reference: assert
        BlockID_t b;
        memcpy(&b, base + index * sizeof(BlockID_t), sizeof b);
        origin_world(xx, yy, zz) = (BlockID_t)clampi(b, 0, static_block_palette_size);
    }
    worldDirty = true;
}

void LumInternalRenderer::save_scene(const char* file) {                 // :221-231
    FILE* f = fopen(file, "wb");
    if (!f) return;
    int32_t hdr[3] = {world_size.x, world_size.y, world_size.z};
    fwrite(hdr, sizeof hdr, 1, f);
    fwrite(origin_world.data(), sizeof(BlockID_t), (size_t)world_size.x * world_size.y * world_size.z, f);
    fclose(f);
}

static void read_vox_or_die(const char* path, lumvox::Scene& sc) {
    std::string err;
    if (!lumvox::read_scene_file(path, sc, err)) {                       // reference: assert(scene != NULL)
        fprintf(stderr, "lum-b200: %s: %s\n", path, err.c_str());
        abort();
    }
    if (sc.num_models != 1) { fprintf(stderr, "lum-b200: %s: expected exactly one model\n", path); abort(); }
}

void LumInternalRenderer::load_block(BlockWithMesh* block, const char* vox_file) {   // :308-334
    lumvox::Scene sc;
    read_vox_or_die(vox_file, sc);
    // the reference writes voxels[x][y][z] for the file's own size and would overrun the 16^3
    // array for anything larger; clamp instead
    const int nx = sc.sx < BLOCK_SIZE ? sc.sx : BLOCK_SIZE, ny = sc.sy < BLOCK_SIZE ? sc.sy : BLOCK_SIZE,
              nz = sc.sz < BLOCK_SIZE ? sc.sz : BLOCK_SIZE;
    memset(block->voxels, 0, sizeof(BlockVoxels));
    for (int x = 0; x < nx; x++)
        for (int y = 0; y < ny; y++)
            for (int z = 0; z < nz; z++)
                block->voxels[x][y][z] = sc.voxels[(size_t)x + (size_t)y * sc.sx + (size_t)z * sc.sx * sc.sy];
}

void LumInternalRenderer::load_mesh(InternalMeshModel* mesh, const char* vox_file, bool make_vertices,
                                    bool extrude_palette, Material* mat_palette) {   // :288-305
    lumvox::Scene sc;
    read_vox_or_die(vox_file, sc);
    load_mesh(mesh, sc.voxels.data(), sc.sx, sc.sy, sc.sz, make_vertices);
    if (extrude_palette && !hasPalette) {
        hasPalette = true;
        lumvox::extract_materials(sc, reinterpret_cast<float*>(mat_palette));
    }
}

void LumInternalRenderer::load_mesh(InternalMeshModel* mesh, Voxel* voxels, int sx, int sy, int sz, bool) {   // :376-385
    if (!mesh || !voxels || sx <= 0 || sy <= 0 || sz <= 0) abort();
    mesh->size = ivec3(sx, sy, sz);
    mesh->voxels.assign(voxels, voxels + (size_t)sx * sy * sz);
    check(lumcu_upload_model(stage, mesh->voxels.data(), sx, sy, sz, &mesh->device_model), "lumcu_upload_model");
}

void LumInternalRenderer::free_mesh(InternalMeshModel* mesh) {           // :397-430
    if (!mesh) abort();
    if (mesh->device_model >= 0) check(lumcu_free_model(stage, mesh->device_model), "lumcu_free_model");
    mesh->device_model = -1;
    mesh->voxels.clear();
}

void LumInternalRenderer::extract_palette(const char* scene_file, Material* mat_palette) {   // :273-286
    lumvox::Scene sc;
    std::string err;
    if (!lumvox::read_scene_file(scene_file, sc, err)) { fprintf(stderr, "lum-b200: %s: %s\n", scene_file, err.c_str()); abort(); }
    hasPalette = true;
    lumvox::extract_materials(sc, reinterpret_cast<float*>(mat_palette));
}

void LumInternalRenderer::updateBlockPalette(BlockWithMesh* block_palette) {   // :36-86
    // The reference linearises blocks [0, static) into the 1024x1024x16 atlas; the device mirror
    // keeps BlockVoxels order, 4 KiB per brick, so the static bricks go up as they are.
    const int n = static_block_palette_size;
    std::vector<uint8_t> bricks((size_t)n * 4096, 0);
    for (int i = 1; i < n; i++) memcpy(bricks.data() + (size_t)i * 4096, block_palette[i].voxels, 4096);   // block 0 = air
    check(lumcu_upload_bricks(stage, 0, n, bricks.data()), "lumcu_upload_bricks");
}

void LumInternalRenderer::updateMaterialPalette(Material* material_palette) {   // :89-111
    static_assert(sizeof(Material) == 6 * sizeof(float), "Material must be 6 packed floats");
    check(lumcu_upload_materials(stage, reinterpret_cast<const float*>(material_palette)), "lumcu_upload_materials");
}

// ---- camera (internal_render.cpp:923-934) ---------------------------------------------------
void Camera::updateCamera() {
    const dvec3 up = dvec3(0.0, 0.0, 1.0);
    viewSize = dvec2(originViewSize.x / pixelsInVoxel, originViewSize.y / pixelsInVoxel);
    const dmat4 view = look_at(cameraPos, cameraPos + cameraDir, up);
    const dmat4 projection = ortho(-viewSize.x / 2.0, viewSize.x / 2.0, viewSize.y / 2.0, -viewSize.y / 2.0, -0.0, +2000.0);
    cameraTransform = mul(projection, view);
    cameraRayDirPlane = normalize(dvec3(cameraDir.x, cameraDir.y, 0));
    horizline = normalize(cross(cameraRayDirPlane, dvec3(0, 0, 1)));
    vertiline = normalize(cross(cameraDir, horizline));
}

// ---- frame -----------------------------------------------------------------------------------
void LumInternalRenderer::place_timestamp(const char* name, double ms) {   // PLACE_TIMESTAMP_OUTSIDE :49-54
    if (!lumal.settings.profile) return;
    if (lumal.currentTimestamp >= (int)lumal.timestampNames.size()) return;
    lumal.timestampNames[lumal.currentTimestamp] = name;
    lumal.ftimestamps[lumal.currentTimestamp] = ms;
    lumal.currentTimestamp++;
}

void LumInternalRenderer::start_frame() {                                // :957-975
    camera.updateCamera();
    lumal.currentTimestamp = 0;
}

void LumInternalRenderer::start_blockify() {                             // :977-987
    copySrc.clear(); copyDst.clear();
    paletteCounter = static_block_palette_size;
    memcpy(current_world.data(), origin_world.data(), (size_t)world_size.x * world_size.y * world_size.z * sizeof(BlockID_t));
}

LumInternalRenderer::AABB LumInternalRenderer::get_shift(dmat4 trans, ivec3 size) {   // :989-1016
    const vec3 box = vec3((float)(size.x - 1), (float)(size.y - 1), (float)(size.z - 1));
    const vec3 corners[8] = {vec3(0, 0, 0), vec3(0, box.y, 0), vec3(0, box.y, box.z), vec3(0, 0, box.z),
                             vec3(box.x, 0, 0), vec3(box.x, box.y, 0), box, vec3(box.x, 0, box.z)};
    vec3 tmin = vec3(std::numeric_limits<float>::max());
    vec3 tmax = vec3(std::numeric_limits<float>::lowest());
    for (const vec3& c : corners) {
        const dvec4 p = mul(trans, dvec4((double)c.x, (double)c.y, (double)c.z, 1.0));
        const vec3 point = vec3((float)p.x, (float)p.y, (float)p.z);
        for (int k = 0; k < 3; k++) {
            tmin[k] = (point[k] < tmin[k]) ? point[k] : tmin[k];
            tmax[k] = (tmax[k] < point[k]) ? point[k] : tmax[k];
        }
    }
    return AABB{tmin, tmax};
}

void LumInternalRenderer::blockify_mesh(InternalMeshModel* mesh, MeshTransform* trans) {   // :1029-1076
    const dmat4 transform = model_transform(trans->rot, trans->shift);   // shift * rotate
    const AABB b = get_shift(transform, mesh->size);
    ivec3 bmin, bmax;
    for (int k = 0; k < 3; k++) {
        bmin[k] = clampi((f2i(b.min[k]) - 1) / 16, 0, world_size[k] - 1);
        bmax[k] = clampi((f2i(b.max[k]) + 1) / 16, 0, world_size[k] - 1);
    }
    for (int zz = bmin.z; zz <= bmax.z; zz++)
    for (int yy = bmin.y; yy <= bmax.y; yy++)
    for (int xx = bmin.x; xx <= bmax.x; xx++) {
        const int current_block = current_world(xx, yy, zz);
        if (current_block < static_block_palette_size) {                 // static: allocate a temp brick
            if (paletteCounter >= BLOCK_PALETTE_SIZE) { fprintf(stderr, "lum-b200: block palette exhausted\n"); abort(); }
            copySrc.push_back(current_block);                            // src 0 = "leave zeroed"
            copyDst.push_back(paletteCounter);
            current_world(xx, yy, zz) = (BlockID_t)paletteCounter;
            paletteCounter++;
        }
    }
    worldDirty = true;
}

void LumInternalRenderer::end_blockify() {                               // :1092-1097 (+ world part of exec_copies)
    // the world only travels when it changed: static scenes keep the device copy
    if (worldDirty) check(lumcu_upload_world(stage, current_world.data()), "lumcu_upload_world");
    worldDirty = !copyDst.empty();     // a frame with dynamic blocks must be followed by a re-upload
}

void LumInternalRenderer::blockify_custom(void* ptr) {                   // :1099-1103
    check(lumcu_upload_world(stage, static_cast<const int16_t*>(ptr)), "lumcu_upload_world");
    worldDirty = true;
}

void LumInternalRenderer::exec_copies() {                                // :1375-1476 (brick part)
    if (!copyDst.empty())
        check(lumcu_copy_bricks(stage, copySrc.data(), copyDst.data(), (int32_t)copyDst.size()), "lumcu_copy_bricks");
}

void LumInternalRenderer::start_map() {}                                 // :1478-1483

void LumInternalRenderer::map_mesh(InternalMeshModel* mesh, MeshTransform* trans) {   // :1485-1518
    const dmat4 transform = model_transform(trans->rot, trans->shift);
    const AABB b = get_shift(transform, mesh->size);
    int32_t shift[3], area[3];
    for (int k = 0; k < 3; k++) {
        shift[k] = f2i(floorf(b.min[k]));
        area[k] = f2i(ceilf(b.max[k])) - shift[k];
    }
    float inv[16];
    inverse_to_float(transform, inv);
    check(lumcu_map_model(stage, mesh->device_model, inv, shift, area, static_block_palette_size), "lumcu_map_model");
}

void LumInternalRenderer::end_map() {}                                   // :1520-1527

void LumInternalRenderer::shift_radiance(ivec3 s) {                      // :1218-1335
    if (s.x == 0 && s.y == 0 && s.z == 0) return;
    check(lumcu_shift(stage, s.x, s.y, s.z), "lumcu_shift");
}

void LumInternalRenderer::update_radiance() {                            // :1129-1214
    // request list: the 3x3x3 rule over the uploaded world + not-yet-selected specials, built on
    // the device (:1142-1169)
    std::vector<int16_t> sp;
    sp.reserve(4 * (specialRadianceUpdates.size() + specialRadianceUpdatesWide.size()));
    for (const auto& u : specialRadianceUpdates) { sp.push_back(u.x); sp.push_back(u.y); sp.push_back(u.z); sp.push_back(u.w); }
    for (const auto& u : specialRadianceUpdatesWide) { sp.push_back(u.x); sp.push_back(u.y); sp.push_back(u.z); sp.push_back(u.w); }
    specialRadianceUpdates.clear();
    specialRadianceUpdatesWide.clear();
    // (no count comes back: the frame is queued without waiting for the GPU, as the reference records its command
    // buffer; request_count() asks the stage when somebody wants the number)
    check(lumcu_build_requests(stage, sp.empty() ? nullptr : sp.data(), (int32_t)(sp.size() / 4), nullptr), "lumcu_build_requests");
    // controller (:1190-1201)
    if (radianceControllerEnabled) {
        if (timeTakenByRadiance < 1.8f) magicSize--;
        else if (timeTakenByRadiance > 2.2f) magicSize++;
    }
    magicSize = magicSize < 1 ? 1 : magicSize;
    magicSize = magicSize > 10 ? 10 : magicSize;
    // push constants {iFrame, 0, magicSize, iFrame % magicSize}; dispatch N / magicSize workgroups (:1202-1206)
    check(lumcu_update(stage, lumal.iFrame, magicSize, lumal.iFrame % magicSize), "lumcu_update");
}

int64_t LumInternalRenderer::request_count() {
    int64_t n = 0;
    check(lumcu_request_count(stage, &n), "lumcu_request_count");
    return n;
}

void LumInternalRenderer::end_frame() {                                  // :2081-2180, al.cpp:385-468
    float ms = timeTakenByRadiance;
    // the pair of timestamps around the dispatch (:1205-1207).  Reading it waits for the dispatch, so
    // it is skipped when nothing consumes it (controller off and profiling off).
    if (radianceTimeOverride >= 0.0f) ms = radianceTimeOverride;
    else if (radianceControllerEnabled || lumal.settings.profile) check(lumcu_last_update_ms(stage, &ms), "lumcu_last_update_ms");
    timeTakenByRadiance = ms;                                            // :2173
    // timestamps 0 and 1 bracket the radiance dispatch (:1205-1207)
    lumal.currentTimestamp = 0;
    place_timestamp("update_radiance", 0.0);
    place_timestamp("update_radiance", (double)ms);
    if (lumal.settings.profile) {
        for (size_t i = 0; i < lumal.average_ftimestamps.size(); i++) {
            if (lumal.iFrame > 5) lumal.average_ftimestamps[i] = lumal.average_ftimestamps[i] * (1.0 - 0.07) + lumal.ftimestamps[i] * 0.07;
            else lumal.average_ftimestamps[i] = lumal.ftimestamps[i];
        }
    }
    lumal.iFrame++;                                                      // al.cpp:453
}

// ---- readbacks -------------------------------------------------------------------------------
void LumInternalRenderer::download_radiance_field(uint32_t* words) { check(lumcu_download_field(stage, words), "lumcu_download_field"); }
void LumInternalRenderer::upload_radiance_field(const uint32_t* words) { check(lumcu_upload_field(stage, words), "lumcu_upload_field"); }
void LumInternalRenderer::glossy(int width, int height, const uint8_t* mat_norm, const float* depth, float* frame, uint8_t* covered) {
    LumcuGlossyView v{};
    const dvec3 hs = dvec3(camera.horizline.x * camera.viewSize.x / 2.0, camera.horizline.y * camera.viewSize.x / 2.0, camera.horizline.z * camera.viewSize.x / 2.0);
    const dvec3 vs = dvec3(camera.vertiline.x * camera.viewSize.y / 2.0, camera.vertiline.y * camera.viewSize.y / 2.0, camera.vertiline.z * camera.viewSize.y / 2.0);
    const double* src[5] = {&camera.cameraPos.x, &camera.cameraDir.x, &hs.x, &vs.x, &lightDir.x};
    float* dst[5] = {v.campos, v.camdir, v.horizline_scaled, v.vertiline_scaled, v.global_light_dir};
    for (int k = 0; k < 5; k++) for (int i = 0; i < 3; i++) dst[k][i] = (float)src[k][i];    // vec4(dvec3, 0): double -> float
    v.frame_size[0] = (float)width; v.frame_size[1] = (float)height;
    check(lumcu_glossy(stage, &v, width, height, mat_norm, depth, frame, covered), "lumcu_glossy");
}
void LumInternalRenderer::download_block_palette(int first, int count, uint8_t* v) { check(lumcu_download_bricks(stage, first, count, v), "lumcu_download_bricks"); }
void LumInternalRenderer::wait_idle() { if (stage) check(lumcu_wait_idle(stage), "lumcu_wait_idle"); }
void LumInternalRenderer::download_world(int16_t* ids) { check(lumcu_download_world(stage, ids), "lumcu_download_world"); }

}  // namespace LumInternal
