// ref_host.cpp -- TEST INFRASTRUCTURE ONLY: the reference's own HOST-side C++ of the radiance path, compiled in place.
//
// oracle/Makefile cuts these functions out of /root/reference/src/renderer/src/internal_render.cpp, verbatim, into
// oracle/_ref/host_*.inc (host_prep.py; build products, never committed):
//   start_blockify            :977-987     get_shift        :989-1013     blockify_mesh   :1029-1076
//   update_radiance (list)    :1129-1169   get_block_xy     :2182-2187
//   map_mesh (the arithmetic) :1501-1510
// plus table3d<T> from internal_render.hpp:121-198, and compiles them against the reference's vendored glm
// (external/lum-al/external/glm) and its own assert macros (src/macros/assert.hpp).  What this file supplies is only
// what the cut-out text needs around it: a LumInternalRenderer with the members those functions touch, the two
// reference structs they take (MeshTransform, InternalMeshModel reduced to `size`), and the handful of Vulkan names
// blockify_mesh fills in (VkImageCopy) -- no Vulkan is called.  The exported C functions run the reference code and
// hand back what it computed, so that tests/test_ref_host.py can hold oracle.place_model / blockify /
// build_requests (oracle/lum_oracle_feeders.c, lum_oracle.c) against the reference itself.
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <vector>

#define GLM_ENABLE_EXPERIMENTAL                 // internal_render.cpp:15
#include <glm/glm.hpp>
#include <glm/gtc/type_precision.hpp>
#include <glm/gtx/quaternion.hpp>
#include "macros/assert.hpp"                     // the reference's own ASSERT / SLOC_DECLARE_DEFARG

using std::vector;
using glm::u8, glm::u16, glm::u32;               // internal_render.cpp:26-41
using glm::i8, glm::i16, glm::i32;
using glm::quat;
using glm::ivec2, glm::ivec3, glm::ivec4;
using glm::i8vec4;
using glm::uvec3;
using glm::vec3, glm::vec4;
using glm::mat4, glm::dmat4;
using glm::identity;

// ---- the little of Vulkan that blockify_mesh writes into (vulkan_core.h: VkImageCopy and its parts) ----
typedef uint32_t VkFlags;
enum { VK_IMAGE_ASPECT_COLOR_BIT = 1 };
struct VkImageSubresourceLayers { VkFlags aspectMask; uint32_t mipLevel, baseArrayLayer, layerCount; };
struct VkOffset3D { int32_t x, y, z; };
struct VkExtent3D { uint32_t width, height, depth; };
struct VkImageCopy { VkImageSubresourceLayers srcSubresource; VkOffset3D srcOffset; VkImageSubresourceLayers dstSubresource; VkOffset3D dstOffset; VkExtent3D extent; };

#define TRACE()                                  // lum-al's trace macro (defines/macros.hpp): a log line, nothing else

namespace LumInternal {
const int BLOCK_PALETTE_SIZE_X = 64;             // internal_types.hpp:16-17
const int BLOCK_PALETTE_SIZE_Y = 64;
typedef glm::i16 BlockID_t;                      // internal_types.hpp:23
typedef struct MeshTransform {                   // internal_types.hpp:41-44
    glm::quat rot = glm::quat_identity<float, glm::defaultp>();
    glm::vec3 shift = glm::vec3(0);
} MeshTransform;
struct InternalMeshModel { ivec3 size; };        // internal_types.hpp: only `size` is read here

#include "_ref/host_table3d.inc"                 // internal_render.hpp:121-198

struct CommandBuffer {};
template <class T> struct ring1 { T v; T& current() { return v; } };

class LumInternalRenderer {                      // the members of internal_render.hpp the cut-out functions use
public:
    uvec3 world_size = uvec3(48, 48, 16);        // :247
    int static_block_palette_size = 15;
    vector<VkImageCopy> blockCopyQueue = {};     // :252
    table3d<BlockID_t> origin_world = {}, current_world = {};   // :388-389
    vector<i8vec4> radianceUpdates;              // :410
    vector<i8vec4> specialRadianceUpdates;       // :411
    int paletteCounter = 0;                      // :443
    ring1<CommandBuffer> computeCommandBuffers;
    struct AABB { vec3 min; vec3 max; };         // :322-328
    AABB get_shift(dmat4 trans, ivec3 size);
    void start_blockify();
    void blockify_mesh(InternalMeshModel* mesh, MeshTransform* trans);
    void update_radiance();
    // map_mesh's arithmetic (:1501-1510), wrapped: the lines in between are descriptor and push-constant calls
    void map_mesh_arithmetic(InternalMeshModel* mesh, MeshTransform* mesh_trans, float* inv16, int* shift3, int* area3) {
        using glm::floor; using glm::ceil; using glm::inverse;
#include "_ref/host_map_mesh.inc"
        std::memcpy(inv16, &itrans_shift.trans[0][0], 16 * sizeof(float));
        for (int k = 0; k < 3; k++) { shift3[k] = itrans_shift.shift[k]; area3[k] = map_area[k]; }
    }
};
}  // namespace LumInternal

using namespace LumInternal;
ivec2 get_block_xy(int N);                       // internal_render.cpp:56
using glm::toMat4; using glm::translate;
#include "_ref/host_functions.inc"               // start_blockify, get_shift, blockify_mesh, update_radiance (list part), get_block_xy

// ---- C entry points for the tests -------------------------------------------------------------------------------
static void fill_world(LumInternalRenderer& r, int wx, int wy, int wz, const int16_t* world) {
    r.world_size = uvec3(wx, wy, wz);
    r.origin_world.allocate(r.world_size);
    r.current_world.allocate(r.world_size);
    std::memcpy(r.origin_world.data(), world, sizeof(int16_t) * (size_t)wx * wy * wz);      // load_scene fills data() x-fastest (load_stuff.cpp:207-216)
    std::memcpy(r.current_world.data(), world, sizeof(int16_t) * (size_t)wx * wy * wz);
}

extern "C" {

// get_shift + the border arithmetic of blockify_mesh (:1030-1037, run on an all-static world so that the block range
// shows as the set of allocated cells) + map_mesh's arithmetic.  rot = (w,x,y,z).
int ref_host_place_model(const float rot_wxyz[4], const float shift[3], const int32_t size[3], const int32_t world_size[3],
                         float aabb_min[3], float aabb_max[3], int32_t block_min[3], int32_t block_max[3],
                         int32_t map_shift[3], int32_t map_area[3], float inverse_trans[16]) {
    LumInternalRenderer r;
    std::vector<int16_t> zero((size_t)world_size[0] * world_size[1] * world_size[2], 0);
    fill_world(r, world_size[0], world_size[1], world_size[2], zero.data());
    r.static_block_palette_size = 1;
    MeshTransform t; t.rot = quat(rot_wxyz[0], rot_wxyz[1], rot_wxyz[2], rot_wxyz[3]); t.shift = vec3(shift[0], shift[1], shift[2]);
    InternalMeshModel m{ivec3(size[0], size[1], size[2])};
    mat4 rotate = toMat4(t.rot);
    mat4 sh = translate(identity<mat4>(), t.shift);
    LumInternalRenderer::AABB b = r.get_shift(sh * rotate, m.size);
    for (int k = 0; k < 3; k++) { aabb_min[k] = b.min[k]; aabb_max[k] = b.max[k]; block_min[k] = INT32_MAX; block_max[k] = INT32_MIN; }
    r.start_blockify();
    r.blockify_mesh(&m, &t);
    for (int z = 0; z < world_size[2]; z++) for (int y = 0; y < world_size[1]; y++) for (int x = 0; x < world_size[0]; x++)
        if (r.current_world(x, y, z) != 0) {
            const int p[3] = {x, y, z};
            for (int k = 0; k < 3; k++) { if (p[k] < block_min[k]) block_min[k] = p[k]; if (p[k] > block_max[k]) block_max[k] = p[k]; }
        }
    r.map_mesh_arithmetic(&m, &t, inverse_trans, map_shift, map_area);
    r.origin_world.deallocate(); r.current_world.deallocate();
    return 0;
}

// start_blockify + blockify_mesh per model in draw order.  Returns the number of blocks allocated; copy pairs as
// palette ids recovered from the VkImageCopy offsets (get_block_xy: id = x + 64*y); zero-source blocks (which the
// reference does not queue, :1051) are reported with src 0 in allocation order, like oracle.blockify does.
int ref_host_blockify(int wx, int wy, int wz, const int16_t* origin_world, int16_t* current_world, int static_count,
                      int nmodels, const float* rot_wxyz, const float* shift, const int32_t* size,
                      int32_t* copy_src, int32_t* copy_dst, int32_t cap, int32_t* queued_copies) {
    LumInternalRenderer r;
    fill_world(r, wx, wy, wz, origin_world);
    r.static_block_palette_size = static_count;
    r.start_blockify();
    for (int i = 0; i < nmodels; i++) {
        MeshTransform t; t.rot = quat(rot_wxyz[4 * i], rot_wxyz[4 * i + 1], rot_wxyz[4 * i + 2], rot_wxyz[4 * i + 3]);
        t.shift = vec3(shift[3 * i], shift[3 * i + 1], shift[3 * i + 2]);
        InternalMeshModel m{ivec3(size[3 * i], size[3 * i + 1], size[3 * i + 2])};
        r.blockify_mesh(&m, &t);
    }
    const int n = r.paletteCounter - static_count;
    std::vector<int32_t> src((size_t)(n > 0 ? n : 1), 0);
    for (const VkImageCopy& c : r.blockCopyQueue) {
        const int s = c.srcOffset.x / 16 + BLOCK_PALETTE_SIZE_X * (c.srcOffset.y / 16);
        const int d = c.dstOffset.x / 16 + BLOCK_PALETTE_SIZE_X * (c.dstOffset.y / 16);
        if (d - static_count >= 0 && d - static_count < n) src[d - static_count] = s;
    }
    for (int i = 0; i < n && i < cap; i++) { copy_src[i] = src[i]; copy_dst[i] = static_count + i; }
    *queued_copies = (int32_t)r.blockCopyQueue.size();
    std::memcpy(current_world, r.current_world.data(), sizeof(int16_t) * (size_t)wx * wy * wz);
    r.origin_world.deallocate(); r.current_world.deallocate();
    return n;
}

// update_radiance's request list (:1132-1169).  special: i8vec4[nspecial]; out: i8vec4[cap].  Returns the list length.
long long ref_host_build_requests(int wx, int wy, int wz, const int16_t* world, const int8_t* special, int nspecial,
                                  int8_t* out, long long cap) {
    LumInternalRenderer r;
    fill_world(r, wx, wy, wz, world);
    for (int i = 0; i < nspecial; i++) r.specialRadianceUpdates.push_back(i8vec4(special[4 * i], special[4 * i + 1], special[4 * i + 2], special[4 * i + 3]));
    r.update_radiance();
    const long long n = (long long)r.radianceUpdates.size();
    for (long long i = 0; i < n && i < cap; i++) for (int k = 0; k < 4; k++) out[4 * i + k] = r.radianceUpdates[(size_t)i][k];
    r.origin_world.deallocate(); r.current_world.deallocate();
    return r.specialRadianceUpdates.empty() ? n : -1;       // the reference clears the queue (:1168)
}

}  // extern "C"
