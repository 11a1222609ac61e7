// ref_glossy_glsl.cpp -- TEST INFRASTRUCTURE (oracle/_ref tool; built only where /root/reference exists).
//
// Executes the REFERENCE's own shaders/glossy.frag on the CPU, the same way ref_radiance_glsl.cpp runs
// radiance.comp: the shader text (and the uniform block of shaders/common/ubo.glsl) is read where it lies
// under /root/reference, token-rewritten by oracle/glsl_prep.py into oracle/_ref/glossy_frag.inc (build
// product, never committed) and compiled as the body of a C++ struct on oracle/glsl_shim.hpp.  One
// invocation per fragment that passes the stencil rule of fillStencilGlossy.frag (roughness <= 0.5), as the
// fullscreen draw of LumInternalRenderer::glossy() (internal_render.cpp:2027-2037) would run it.
// Validates lum_oracle_glossy (oracle/lum_oracle.c) and generates tests/golden/ref_glossy_golden.npz.
#include "glsl_shim.hpp"
#include <cstring>
#include <vector>

namespace glsl {
struct GlossyShader : ShaderBase {
#include "_ref/glossy_frag.inc"
};
}  // namespace glsl

extern "C" {

// view: campos[3], camdir[3], horizline_scaled[3], vertiline_scaled[3], globalLightDir[3], frame_size[2] (17 floats).
// mat_norm uint8 [h][w][4]; depth float [h][w]; frame float [h][w][4] written where covered; covered uint8 [h][w]
// and fetches uint32 [h][w] (texelFetch(blocks) calls of the invocation) may be NULL.
int ref_glsl_glossy(int wx, int wy, int wz, const int16_t* world, const uint8_t* bricks, int nbricks,
                    const float* materials, const uint32_t* field, const float* view, int width, int height,
                    const uint8_t* mat_norm, const float* depth, float* frame, uint8_t* covered, uint32_t* fetches) {
    using namespace glsl;
    if (wx <= 0 || wy <= 0 || wz <= 0 || nbricks < 1 || nbricks > 4096 || width < 1 || height < 1) return -1;
    std::vector<uint8_t> atlas((size_t)1024 * 1024 * 16, 0);              // load_stuff.cpp:38-61
    for (int n = 0; n < nbricks; n++)
        for (int x = 0; x < 16; x++)
            for (int y = 0; y < 16; y++)
                for (int z = 0; z < 16; z++)
                    atlas[(size_t)(16 * (n % 64) + x) + 1024 * ((size_t)(16 * (n / 64) + y) + 1024 * (size_t)z)] =
                        bricks[(((size_t)n * 16 + x) * 16 + y) * 16 + z];
#pragma omp parallel
    {
        GlossyShader sh;
        sh.world_size = ivec3(wx, wy, wz);
        sh.matNorm.data = mat_norm;          sh.matNorm.size = ivec2(width, height);
        sh.depthBuffer.data = depth;         sh.depthBuffer.size = ivec2(width, height);
        sh.blocks.data = world;              sh.blocks.size = ivec3(wx, wy, wz);
        sh.blockPalette.data = atlas.data(); sh.blockPalette.size = ivec3(1024, 1024, 16);
        sh.voxelPalette.data = materials;    sh.voxelPalette.size = ivec2(6, 256);
        sh.radianceCache.data = field;       sh.radianceCache.size = ivec3(wx, wy, wz);
        sh.ubo.campos = vec4(vec3(view[0], view[1], view[2]), 0);
        sh.ubo.camdir = vec4(vec3(view[3], view[4], view[5]), 0);
        sh.ubo.horizline_scaled = vec4(vec3(view[6], view[7], view[8]), 0);
        sh.ubo.vertiline_scaled = vec4(vec3(view[9], view[10], view[11]), 0);
        sh.ubo.globalLightDir = vec4(vec3(view[12], view[13], view[14]), 0);
        sh.ubo.frame_size = vec2(view[15], view[16]);
        sh.ubo.timeseed = 0;
#pragma omp for schedule(dynamic, 4)
        for (int y = 0; y < height; y++)
            for (int x = 0; x < width; x++) {
                const size_t pi = (size_t)y * width + x;
                const float rough = materials[6 * (size_t)mat_norm[4 * pi] + 5];
                const bool cov = !(rough > 0.5f);                          // fillStencilGlossy.frag:34-36
                if (covered) covered[pi] = cov ? 1 : 0;
                if (fetches) fetches[pi] = 0;
                if (!cov) continue;
                sh.gl_FragCoord = vec4((float)x + 0.5f, (float)y + 0.5f, 0.0f, 1.0f);
                sh.blocks.fetches = 0;
                sh.main();
                for (int c = 0; c < 4; c++) frame[4 * pi + c] = sh.frame_color[c];
                if (fetches) fetches[pi] = (uint32_t)sh.blocks.fetches;
            }
    }
    return 0;
}

}  // extern "C"
