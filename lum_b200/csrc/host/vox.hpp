// vox.hpp -- MagicaVoxel .vox reader for the parts Lum's loaders consume.
//
// The reference decodes .vox with the vendored ogt_vox (common/ogt_vox.hpp) and uses, per file,
// models[0] (size + dense voxel grid, load_stuff.cpp:296,319-331) and the palette + MATL
// materials (extract_palette_mem, load_stuff.cpp:233-260).  This is an independent reader of the
// published format (chunks SIZE, XYZI, RGBA, IMAP, MATL) that reproduces the decoder conventions
// those call sites depend on: palette rotated by one so colour index i is palette[i], IMAP
// display-order remapping of palette, materials and voxel indices, material ids taken & 0xFF.
// tests/test_vox_loader.py checks it against the reference's own decoder (oracle/_ref) on the
// demo assets.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace lumvox {

enum MatlType { MATL_DIFFUSE = 0, MATL_METAL = 1, MATL_GLASS = 2, MATL_EMIT = 3, MATL_BLEND = 4, MATL_MEDIA = 5 };

struct Matl {
    int type = MATL_DIFFUSE;
    float metal = 0, rough = 0, emit = 0, flux = 0;
};

struct Scene {
    int sx = 0, sy = 0, sz = 0;          // models[0]
    std::vector<uint8_t> voxels;         // x + sx*(y + sy*z), colour indices (0 = empty)
    int num_models = 0;
    uint8_t palette[256][4];             // r,g,b,a per colour index
    Matl matl[256];
};

// Returns false (and sets err) on malformed input.
bool read_scene(const uint8_t* data, size_t size, Scene& out, std::string& err);
bool read_scene_file(const char* path, Scene& out, std::string& err);

// extract_palette_mem (load_stuff.cpp:233-260): Material[256] as 6 floats each
// (r,g,b,a = palette/256; emmit = emit*(2+4*flux) for emissive; rough by type).
void extract_materials(const Scene& s, float* out_256x6);

}  // namespace lumvox
