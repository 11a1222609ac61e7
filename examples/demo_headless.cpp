// demo_headless.cpp -- the setup and frame loop of Lum's examples/demo.cpp (:114-228), written
// against this repository's API mirror: same calls in the same order, minus GLFW/input/ECS.
// Runs N frames of the demo scene on cuda:0 and prints the radiance-stage timings.
//
//   g++ -std=c++17 -Iinclude examples/demo_headless.cpp -Llum_b200 -llum -Wl,-rpath,$PWD/lum_b200 -o demo_headless
//   ./demo_headless /path/to/lum/assets 64
#include <lum.hpp>
#include <lum_opaque.hpp>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

using namespace lum_math;

int main(int argc, char** argv) {
    const std::string assets = argc > 1 ? argv[1] : "assets";
    const int frames = argc > 2 ? atoi(argv[2]) : 16;

    Lum::Settings settings = {};
        settings.fullscreen = false;
        settings.vsync = false;
        settings.world_size = ivec3(48, 48, 16);
        settings.static_block_palette_size = 15;
        settings.maxParticleCount = 8128;
    Lum::Renderer lum = Lum::Renderer(15, 4096, 64, 64, 64);

    Lum::MeshFoliage grass = lum.loadFoliage("shaders/compiled/grass.vert.spv", 6, 10);
    lum.init(settings);
    lum.loadWorld(assets + "/scene");

    Lum::MeshTransform tank_body_trans, tank_head_trans;
    Lum::MeshModel tank_body = lum.loadModel(assets + "/tank_body.vox", /*extract palette if no found = */ true);
    tank_body_trans.shift += vec3(13.1f, 14.1f, 3.1f) * 16.0f;
    Lum::MeshModel tank_head = lum.loadModel(assets + "/tank_head.vox", true);
    Lum::MeshLiquid water = lum.loadLiquid(69, 42);
    Lum::MeshVolumetric smoke = lum.loadVolumetric(1, .5, {});

    const char* blocks[] = {"dirt", "grass", "grassNdirt", "stone_dirt", "bush", "leaves", "iron", "lamp", "stone_brick",
                            "stone_brick_cracked", "stone_pack", "bark", "wood", "planks"};
    for (int i = 0; i < 14; i++) lum.loadBlock(Lum::MeshBlock(i + 1), assets + "/" + blocks[i] + ".vox");

    lum.uploadBlockPaletteToGPU();
    lum.uploadMaterialPaletteToGPU();
    lum.waitIdle();

    for (int f = 0; f < frames && !lum.should_close; f++) {
        // stand-in for process_physics / process_animations: drive forward and turn slowly
        tank_body_trans.shift += vec3(0.75f, 0.5f, 0.0f);
        tank_body_trans.rot = quat(std::cos(0.025f * f), 0.0f, 0.0f, std::sin(0.025f * f));
        tank_head_trans = tank_body_trans;
        tank_head_trans.shift += vec3(2.0f, 3.0f, (float)tank_body.getSize().z);

        lum.startFrame();
        lum.drawWorld();
        lum.drawParticles();
        lum.drawModel(tank_body, tank_body_trans);
        lum.drawModel(tank_head, tank_head_trans);

        for (int xx = 4; xx < 20; xx++) for (int yy = 4; yy < 20; yy++) {
            if ((xx >= 5) and (xx < 12) and (yy >= 6) and (yy < 16)) continue;
            lum.drawFoliageBlock(grass, ivec3(xx * 16, yy * 16, 16));
        }
        for (int xx = 5; xx < 12; xx++) for (int yy = 6; yy < 16; yy++) {
            lum.drawLiquidBlock(water, ivec3(xx * 16, yy * 16, 14));
            lum.opaque_members->render.specialRadianceUpdates.push_back(i8vec4(ivec3(xx, yy, 2), 0));
            lum.opaque_members->render.specialRadianceUpdates.push_back(i8vec4(ivec3(xx, yy, 1), 0));
        }
        for (int xx = 8; xx < 10; xx++) for (int yy = 10; yy < 13; yy++) lum.drawVolumetricBlock(smoke, ivec3(xx * 16, yy * 16, 20));

        lum.prepareFrame();
        lum.submitFrame();
    }

    // prints GPU profile data (examples/demo.cpp:230-237)
    auto& lumal = lum.opaque_members->render.lumal;
    if (lumal.settings.profile) {
        printf("%-22s: %7.3f: %7.3f\n", lumal.timestampNames[0], 0.0, 0.0);
        for (int i = 1; i < lumal.currentTimestamp; i++) {
            double time_point = lumal.average_ftimestamps[i] - lumal.average_ftimestamps[0];
            double time_diff = lumal.average_ftimestamps[i] - lumal.average_ftimestamps[i - 1];
            printf("%3d %-22s: %7.3f: %7.3f\n", i, lumal.timestampNames[i], time_point, time_diff);
        }
    }
    std::vector<uint32_t> field(48 * 48 * 16);
    lum.opaque_members->render.download_radiance_field(field.data());
    unsigned long long sum = 0;
    for (uint32_t w : field) sum += (w & 1023u) + ((w >> 10) & 1023u) + ((w >> 20) & 1023u);
    printf("frames %d, requests last frame %lld, magicSize %d, field checksum %llu\n", lum.opaque_members->render.lumal.iFrame,
           (long long)lum.opaque_members->render.request_count(), lum.opaque_members->render.magicSize, sum);
    lum.cleanup();
    return 0;
}
