/* cdemo_headless.c -- examples/cdemo.c of the reference against include/clum.h (C99). */
#include <clum.h>
#include <assert.h>
#include <stdio.h>
#include <stdlib.h>

int main(int argc, char** argv) {
    const char* scene = argc > 1 ? argv[1] : "assets/scene";
    LumSettings settings = LUM_DEFAULT_SETTINGS;
    settings.world_size[0] = 48; settings.world_size[1] = 48; settings.world_size[2] = 16;
    settings.static_block_palette_size = 15;
    LumRenderer lum = lum_create_instance(&settings);
    assert(lum);
    lum_init(lum, &settings);
    lum_demo_load_scene(lum, scene);
    lum_upload_block_palette_to_gpu(lum);
    lum_upload_material_palette_to_gpu(lum);
    LumMeshTransform t = LUM_DEFAULT_MESH_TRANSFORM;
    (void)t;
    for (int f = 0; f < 4; f++) {
        lum_start_frame(lum);
        lum_draw_world(lum);
        lum_prepare_frame(lum);
        lum_submit_frame(lum);
    }
    printf("radiance dispatch: %.3f ms\n", lum_get_timestamp_difference(lum, 0, 1));
    lum_cleanup(lum);
    return 0;
}
