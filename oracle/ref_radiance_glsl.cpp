// ref_radiance_glsl.cpp -- TEST INFRASTRUCTURE (oracle/_ref tool; built only where /root/reference exists).
//
// Executes the REFERENCE's own shaders/radiance.comp on the CPU: the shader text is read where it
// lies under /root/reference, token-rewritten by oracle/glsl_prep.py into oracle/_ref/radiance_comp.inc
// (a build product, never committed) and compiled here as the body of a C++ struct on top of
// oracle/glsl_shim.hpp.  Every statement of the algorithm -- RNG, ray setup, marcher, refinement,
// shading, reduction, temporal mix -- is the reference's; the shim supplies what a Vulkan driver
// would (vector types, built-ins, texture access, subgroup ops) under the decrees of DESIGN.md s.3.
// The dispatch follows internal_render.cpp:1202-1206: workgroup w of nreq/size handles request
// w*size + shift, 64 invocations each, gl_GlobalInvocationID.x = 64*w + lane.
//
// The result validates oracle/lum_oracle.c (tests/test_ref_glsl.py) and generates the golden
// vectors under tests/golden/ (tests/golden/make_ref_glsl_golden.py).
#include "glsl_shim.hpp"
#include <cstdio>
#include <cstring>
#include <vector>

namespace glsl {
struct RadianceShader : ShaderBase {
#include "_ref/radiance_comp.inc"
};
}  // namespace glsl

extern "C" {

// world: int16 [z][y][x]; bricks: uint8 [nbricks][x][y][z] (BlockVoxels order, internal_types.hpp:46);
// materials: float [256][6]; field_old/new: packed words; requests: int8 x,y,z,w (the reference's i8vec4).
// Outputs (each may be NULL): pre_blend [nwg][3]; ray_light [nwg*64][3]; ray_fetches [nwg*64] = number
// of texelFetch(blocks) calls of the invocation (= 2 + 2*coarse iterations + refinement samples).
// Returns 0, or -1 on bad arguments.
int ref_glsl_radiance_update(int wx, int wy, int wz, const int16_t* world,
                             const uint8_t* bricks, int nbricks, const float* materials,
                             const uint32_t* field_old, uint32_t* field_new,
                             const int8_t* requests_xyzw, long nreq, int time_seed, int size, int shift,
                             float* pre_blend, float* ray_light, uint32_t* ray_fetches) {
    using namespace glsl;
    if (wx <= 0 || wy <= 0 || wz <= 0 || nbricks < 1 || nbricks > 4096 || size < 1) return -1;
    const long cells = (long)wx * wy * wz;
    // block palette atlas exactly as the reference lays it out (load_stuff.cpp:38-61):
    // brick N texel (16*(N%64)+x, 16*(N/64)+y, z) of a 1024 x 1024 x 16 image
    std::vector<uint8_t> atlas((size_t)1024 * 1024 * 16, 0);
    for (int n = 0; n < nbricks; n++)
        for (int x = 0; x < 16; x++)
            for (int y = 0; y < 16; y++)
                for (int z = 0; z < 16; z++)
                    atlas[(size_t)(16 * (n % 64) + x) + 1024 * ((size_t)(16 * (n / 64) + y) + 1024 * (size_t)z)] =
                        bricks[(((size_t)n * 16 + x) * 16 + y) * 16 + z];
    if (field_new != field_old) memcpy(field_new, field_old, (size_t)cells * 4);   // Jacobi (D1)

    const long nwg = nreq / size;
    const int R = 64;
#pragma omp parallel
    {
        RadianceShader sh;
        sh.world_size = ivec3(wx, wy, wz);
        sh.blocks.data = world;              sh.blocks.size = ivec3(wx, wy, wz);
        sh.blockPalette.data = atlas.data(); sh.blockPalette.size = ivec3(1024, 1024, 16);
        sh.voxelPalette.data = materials;    sh.voxelPalette.size = ivec2(6, 256);
        sh.radianceCache2read.data = field_old;  sh.radianceCache2read.size = ivec3(wx, wy, wz);
        sh.radianceCache2write.data = field_new; sh.radianceCache2write.size = ivec3(wx, wy, wz);
        sh.reqBuffer.updateReqests = reinterpret_cast<const i8vec4*>(requests_xyzw);
        sh.pco.timeSeed = time_seed; sh.pco.iters = 0; sh.pco.size = size; sh.pco.sihft = shift;
#pragma omp for schedule(dynamic, 16)
        for (long w = 0; w < nwg; w++) {
            vec3 operand[R];
            for (int lane = 0; lane < R; lane++) {
                sh.pass = 0;
                sh.gl_GlobalInvocationID = uvec3((uint)(w * R + lane), 0u, 0u);
                sh.gl_SubgroupID = lane / 32; sh.gl_SubgroupInvocationID = lane % 32;
                sh.blocks.fetches = 0;
                try { sh.main(); } catch (const SubgroupYield&) {}
                operand[lane] = sh.subgroup_operand;
                if (ray_light) for (int c = 0; c < 3; c++) ray_light[((size_t)w * R + lane) * 3 + c] = sh.subgroup_operand[c];
                if (ray_fetches) ray_fetches[(size_t)w * R + lane] = (uint32_t)sh.blocks.fetches;
            }
            // subgroupAdd: xor butterfly 16,8,4,2,1 within each 32-lane subgroup (D3)
            vec3 sums[R / 32];
            for (int sg = 0; sg < R / 32; sg++) {
                vec3 v[32], t[32];
                for (int i = 0; i < 32; i++) v[i] = operand[sg * 32 + i];
                for (int m = 16; m >= 1; m >>= 1) {
                    for (int i = 0; i < 32; i++) t[i] = v[i] + v[i ^ m];
                    for (int i = 0; i < 32; i++) v[i] = t[i];
                }
                sums[sg] = v[0];
            }
            // elected lanes, highest subgroup first so _light[] is complete when subgroup 0 reads it
            for (int sg = R / 32 - 1; sg >= 0; sg--) {
                sh.pass = 1;
                sh.gl_GlobalInvocationID = uvec3((uint)(w * R + sg * 32), 0u, 0u);
                sh.gl_SubgroupID = sg; sh.gl_SubgroupInvocationID = 0;
                sh.subgroup_sum = sums[sg];
                sh.main();
            }
            if (pre_blend) for (int c = 0; c < 3; c++) pre_blend[(size_t)w * 3 + c] = sh.mix_y[c];
        }
    }
    return 0;
}

}  // extern "C"
