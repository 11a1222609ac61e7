// ref_map_glsl.cpp -- TEST INFRASTRUCTURE (oracle/_ref tool; built only where /root/reference exists).
//
// Executes the REFERENCE's own shaders/map.comp on the CPU (same recipe as ref_radiance_glsl.cpp:
// the shader text is read where it lies, token-rewritten by glsl_prep.py into oracle/_ref/map_comp.inc
// and compiled on glsl_shim.hpp).  One call = one map_mesh dispatch (internal_render.cpp:1485-1518):
// ceil(area/4) workgroups of 4x4x4 invocations per axis, push constants {inverse_trans, shift}.
// Invocations run in x-fastest order one after another; invocations whose world voxel has a
// negative component are skipped by this driver (D11 -- in the reference they alias onto a low-edge
// block and race with the legitimate invocation).  The shader hard-codes STATIC_BLOCK_COUNT = 15.
#include "glsl_shim.hpp"
#include <vector>

namespace glsl {
struct MapShader : ShaderBase {
#include "_ref/map_comp.inc"
};
}  // namespace glsl

extern "C" {

// world int16 [z][y][x]; bricks uint8 [nbricks][x][y][z], modified in place; model uint8 [sz][sy][sx]
// (x fastest); inverse_trans column-major; shift/area as map_mesh computes them.
int ref_glsl_map_model(int wx, int wy, int wz, const int16_t* world, uint8_t* bricks, int nbricks,
                       const uint8_t* model, int sx, int sy, int sz,
                       const float* inverse_trans, const int32_t* shift, const int32_t* area) {
    using namespace glsl;
    if (nbricks < 1 || nbricks > 4096) return -1;
    std::vector<uint8_t> atlas((size_t)1024 * 1024 * 16, 0);
    auto at = [](int n, int x, int y, int z) { return (size_t)(16 * (n % 64) + x) + 1024 * ((size_t)(16 * (n / 64) + y) + 1024 * (size_t)z); };
    for (int n = 0; n < nbricks; n++)
        for (int x = 0; x < 16; x++) for (int y = 0; y < 16; y++) for (int z = 0; z < 16; z++)
            atlas[at(n, x, y, z)] = bricks[(((size_t)n * 16 + x) * 16 + y) * 16 + z];
    MapShader sh;
    sh.blocks.data = world; sh.blocks.size = ivec3(wx, wy, wz);
    sh.blockPalette.data = atlas.data(); sh.blockPalette.size = ivec3(1024, 1024, 16);
    sh.modelVoxels.data = const_cast<uint8_t*>(model); sh.modelVoxels.size = ivec3(sx, sy, sz);
    for (int c = 0; c < 4; c++)
        sh.pco.inverse_trans.col[c] = vec4(inverse_trans[4 * c], inverse_trans[4 * c + 1], inverse_trans[4 * c + 2], inverse_trans[4 * c + 3]);
    sh.pco.shift = ivec4(shift[0], shift[1], shift[2], 0);
    int ext[3];
    for (int k = 0; k < 3; k++) ext[k] = area[k] > 0 ? ((area[k] + 3) / 4) * 4 : 0;
    for (int gz = 0; gz < ext[2]; gz++)
        for (int gy = 0; gy < ext[1]; gy++)
            for (int gx = 0; gx < ext[0]; gx++) {
                if (gx + shift[0] < 0 || gy + shift[1] < 0 || gz + shift[2] < 0) continue;   // D11 (gid + 0.5 + shift < 0)
                sh.gl_GlobalInvocationID = uvec3((uint)gx, (uint)gy, (uint)gz);
                sh.main();
            }
    for (int n = 0; n < nbricks; n++)
        for (int x = 0; x < 16; x++) for (int y = 0; y < 16; y++) for (int z = 0; z < 16; z++)
            bricks[(((size_t)n * 16 + x) * 16 + y) * 16 + z] = atlas[at(n, x, y, z)];
    return 0;
}

}  // extern "C"
