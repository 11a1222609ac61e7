// ref_vox_dump.cpp -- TEST INFRASTRUCTURE (oracle/_ref tool, build container only).
//
// Decodes a MagicaVoxel .vox file with the REFERENCE's own vendored parser
// (/root/reference/common/ogt_vox.hpp, compiled in place via -I, never copied) and writes
// what Lum's loaders derive from it, so the product's own .vox reader and the golden
// fixtures under tests/golden/ can be validated against the reference's decoder:
//   int32 sx, sy, sz
//   u8    voxel_data[sx*sy*sz]            (x + y*sx + z*sx*sy; load_stuff.cpp:325-331 input)
//   f32   materials[256][6]               (extract_palette_mem, load_stuff.cpp:233-260)
// Usage: ref_vox_dump in.vox out.bin
#define VOX_IMPLEMENTATION
#include <ogt_vox.hpp>
#include <cstdio>
#include <cstdlib>
#include <vector>

int main(int argc, char** argv) {
    if (argc != 3) { fprintf(stderr, "usage: %s in.vox out.bin\n", argv[0]); return 2; }
    FILE* f = fopen(argv[1], "rb");
    if (!f) { perror(argv[1]); return 1; }
    fseek(f, 0, SEEK_END); long n = ftell(f); fseek(f, 0, SEEK_SET);
    std::vector<uint8_t> buf(n);
    if (fread(buf.data(), 1, n, f) != (size_t)n) return 1;
    fclose(f);
    const ogt::vox_scene* scene = ogt::vox_read_scene(buf.data(), (uint32_t)n);
    if (!scene || scene->num_models < 1) { fprintf(stderr, "bad scene\n"); return 1; }
    const ogt::vox_model* m = scene->models[0];
    FILE* o = fopen(argv[2], "wb");
    int32_t hdr[3] = {(int32_t)m->size_x, (int32_t)m->size_y, (int32_t)m->size_z};
    fwrite(hdr, 4, 3, o);
    fwrite(m->voxel_data, 1, (size_t)m->size_x * m->size_y * m->size_z, o);
    for (int i = 0; i < 256; i++) {
        // restates load_stuff.cpp:235-259 (double intermediate, float store, as in the reference)
        float mat[6];
        mat[0] = scene->palette.color[i].r / 256.0;
        mat[1] = scene->palette.color[i].g / 256.0;
        mat[2] = scene->palette.color[i].b / 256.0;
        mat[3] = scene->palette.color[i].a / 256.0;
        mat[4] = 0;
        float rough;
        switch (scene->materials.matl[i].type) {
            case ogt::matl_type_diffuse: rough = 1.0; break;
            case ogt::matl_type_emit:
                mat[4] = scene->materials.matl[i].emit * (2.0 + scene->materials.matl[i].flux * 4.0);
                rough = 0.5; break;
            case ogt::matl_type_metal:
                rough = (scene->materials.matl[i].rough + (1.0 - scene->materials.matl[i].metal)) / 2.0; break;
            default: rough = 0; break;
        }
        mat[5] = rough;
        fwrite(mat, 4, 6, o);
    }
    fclose(o);
    ogt::vox_destroy_scene(scene);
    return 0;
}
