// vox.cpp -- see vox.hpp.
#include "vox.hpp"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>

namespace lumvox {
namespace {

struct Reader {
    const uint8_t* p;
    size_t n, i = 0;
    size_t left() const { return n - i; }
    bool u32(uint32_t& v) { if (left() < 4) return false; memcpy(&v, p + i, 4); i += 4; return true; }
    bool i32(int32_t& v) { uint32_t u; if (!u32(u)) return false; v = (int32_t)u; return true; }
    bool skip(size_t k) { if (left() < k) { i = n; return false; } i += k; return true; }
    bool str(std::string& s) {
        int32_t len;
        if (!i32(len) || len < 0 || left() < (size_t)len) return false;
        s.assign((const char*)p + i, (size_t)len);
        i += (size_t)len;
        return true;
    }
};

constexpr uint32_t id4(char a, char b, char c, char d) {
    return (uint32_t)(uint8_t)a | ((uint32_t)(uint8_t)b << 8) | ((uint32_t)(uint8_t)c << 16) | ((uint32_t)(uint8_t)d << 24);
}

// MagicaVoxel's default palette in file order (entry k = colour index k+1): the 6x6x6 cube with
// levels ff,cc,99,66,33,00 (blue fastest... stored r,g,b with b varying fastest) minus black,
// then 10-step ramps of red, green, blue and grey, then black.
void default_palette(uint8_t pal[256][4]) {
    static const uint8_t lv[6] = {0xff, 0xcc, 0x99, 0x66, 0x33, 0x00};
    static const uint8_t ramp[10] = {0xee, 0xdd, 0xbb, 0xaa, 0x88, 0x77, 0x55, 0x44, 0x22, 0x11};
    int k = 0;
    for (int r = 0; r < 6; r++)
        for (int g = 0; g < 6; g++)
            for (int b = 0; b < 6; b++) {
                if (r == 5 && g == 5 && b == 5) continue;
                pal[k][0] = lv[r]; pal[k][1] = lv[g]; pal[k][2] = lv[b]; pal[k][3] = 0xff; k++;
            }
    for (int c = 0; c < 3; c++)
        for (int j = 0; j < 10; j++) { pal[k][0] = pal[k][1] = pal[k][2] = 0; pal[k][c] = ramp[j]; pal[k][3] = 0xff; k++; }
    for (int j = 0; j < 10; j++) { pal[k][0] = pal[k][1] = pal[k][2] = ramp[j]; pal[k][3] = 0xff; k++; }
    pal[k][0] = pal[k][1] = pal[k][2] = 0; pal[k][3] = 0xff; k++;      // k == 256
}

}  // namespace

bool read_scene(const uint8_t* data, size_t size, Scene& out, std::string& err) {
    Reader rd{data, size};
    uint32_t magic = 0, version = 0;
    if (!rd.u32(magic) || !rd.u32(version) || magic != id4('V', 'O', 'X', ' ') || (version != 150 && version != 200)) {
        err = "not a .vox file (header/version)";
        return false;
    }
    out = Scene();
    default_palette(out.palette);
    uint8_t imap[256];
    bool have_imap = false;
    uint32_t sx = 0, sy = 0, sz = 0;
    struct Model { uint32_t sx, sy, sz; std::vector<uint8_t> v; };
    std::vector<Model> models;

    while (rd.left() >= 12) {
        uint32_t cid = 0, csize = 0, cchild = 0;
        rd.u32(cid); rd.u32(csize); rd.u32(cchild);
        const size_t body = rd.i;
        if (cid == id4('M', 'A', 'I', 'N')) continue;                       // children follow inline
        if (cid == id4('S', 'I', 'Z', 'E')) {
            if (!rd.u32(sx) || !rd.u32(sy) || !rd.u32(sz) || !sx || !sy || !sz) { err = "bad SIZE chunk"; return false; }
            // XYZI coordinates are bytes: the format cannot address more than 256 voxels per axis.  Anything larger is
            // a corrupt or hostile file (the product of three unchecked 32-bit sizes can wrap the allocation below).
            if (sx > 256 || sy > 256 || sz > 256) { err = "SIZE chunk exceeds the .vox limit of 256 per axis"; return false; }
        } else if (cid == id4('X', 'Y', 'Z', 'I')) {
            uint32_t count = 0;
            if (!sx || !rd.u32(count)) { err = "XYZI before SIZE"; return false; }
            if (count != 0) {                                               // empty models are dropped by the decoder
                Model m{sx, sy, sz, std::vector<uint8_t>((size_t)sx * sy * sz, 0)};
                const size_t avail = rd.left() / 4;
                const size_t nread = count < avail ? count : avail;
                const uint8_t* q = data + rd.i;
                for (size_t k = 0; k < nread; k++) {
                    const uint32_t x = q[4 * k], y = q[4 * k + 1], z = q[4 * k + 2];
                    if (x >= sx || y >= sy || z >= sz) { err = "XYZI voxel outside SIZE"; return false; }
                    m.v[x + (size_t)sx * (y + (size_t)sy * z)] = q[4 * k + 3];
                }
                models.push_back(std::move(m));
            }
        } else if (cid == id4('R', 'G', 'B', 'A')) {
            if (csize != 1024 || rd.left() < 1024) { err = "bad RGBA chunk"; return false; }
            memcpy(out.palette, data + rd.i, 1024);
        } else if (cid == id4('I', 'M', 'A', 'P')) {
            if (csize != 256 || rd.left() < 256) { err = "bad IMAP chunk"; return false; }
            memcpy(imap, data + rd.i, 256);
            have_imap = true;
        } else if (cid == id4('M', 'A', 'T', 'L')) {
            int32_t mid = 0, pairs = 0;
            if (!rd.i32(mid) || !rd.i32(pairs) || pairs < 0) { err = "bad MATL chunk"; return false; }
            Matl& m = out.matl[mid & 0xFF];                                 // material 256 is material 0
            for (int k = 0; k < pairs; k++) {
                std::string key, val;
                if (!rd.str(key) || !rd.str(val)) { err = "bad MATL dictionary"; return false; }
                if (key == "_type") {
                    if (val == "_diffuse") m.type = MATL_DIFFUSE;
                    else if (val == "_metal") m.type = MATL_METAL;
                    else if (val == "_glass") m.type = MATL_GLASS;
                    else if (val == "_emit") m.type = MATL_EMIT;
                    else if (val == "_blend") m.type = MATL_BLEND;
                    else if (val == "_media") m.type = MATL_MEDIA;
                } else if (key == "_metal") m.metal = (float)atof(val.c_str());
                else if (key == "_rough") m.rough = (float)atof(val.c_str());
                else if (key == "_emit") m.emit = (float)atof(val.c_str());
                else if (key == "_flux") m.flux = (float)atof(val.c_str());
            }
        }
        // every chunk (known or not) ends at body + csize; children of non-MAIN chunks are skipped too
        rd.i = body;
        if (!rd.skip((size_t)csize + (size_t)cchild)) break;
    }
    if (models.empty()) { err = "no model in file"; return false; }

    if (have_imap) {
        // imap[display index] = stored index.  Palette, materials and voxel indices go to display order.
        // (an IMAP that is not a permutation leaves entries of the inverse unnamed -- indeterminate in the reference
        // decoder, ogt_vox.hpp:2016-2019; zero here)
        uint8_t inv[256] = {0};
        for (int i = 0; i < 256; i++) inv[imap[i]] = (uint8_t)i;
        uint8_t oldp[256][4];
        memcpy(oldp, out.palette, sizeof oldp);
        for (int i = 0; i < 256; i++) memcpy(out.palette[i], oldp[(imap[i] + 255) & 0xFF], 4);
        Matl oldm[256];
        for (int i = 0; i < 256; i++) oldm[i] = out.matl[i];
        for (int i = 0; i < 256; i++) out.matl[i] = oldm[imap[(i + 255) & 0xFF]];
        for (auto& m : models)
            for (auto& v : m.v) v = (uint8_t)(1 + inv[v]);
    }
    {   // file entry k is colour index k+1: rotate so palette[i] is colour index i; index 0 = empty
        uint8_t last[4];
        memcpy(last, out.palette[255], 4);
        memmove(out.palette[1], out.palette[0], 255 * 4);
        memcpy(out.palette[0], last, 4);
        out.palette[0][3] = 0;
    }
    out.sx = (int)models[0].sx; out.sy = (int)models[0].sy; out.sz = (int)models[0].sz;
    out.voxels = std::move(models[0].v);
    out.num_models = (int)models.size();
    return true;
}

bool read_scene_file(const char* path, Scene& out, std::string& err) {
    FILE* f = fopen(path, "rb");
    if (!f) { err = std::string("cannot open ") + path; return false; }
    std::vector<uint8_t> buf;
    uint8_t tmp[65536];
    size_t n;
    while ((n = fread(tmp, 1, sizeof tmp, f)) > 0) buf.insert(buf.end(), tmp, tmp + n);
    fclose(f);
    return read_scene(buf.data(), buf.size(), out, err);
}

void extract_materials(const Scene& s, float* out) {
    for (int i = 0; i < 256; i++) {
        float* m = out + 6 * i;
        // the reference divides in double and stores float (load_stuff.cpp:236-241)
        m[0] = (float)(s.palette[i][0] / 256.0);
        m[1] = (float)(s.palette[i][1] / 256.0);
        m[2] = (float)(s.palette[i][2] / 256.0);
        m[3] = (float)(s.palette[i][3] / 256.0);
        m[4] = 0.0f;
        float rough;
        switch (s.matl[i].type) {
            case MATL_DIFFUSE: rough = 1.0f; break;
            case MATL_EMIT:
                m[4] = (float)(s.matl[i].emit * (2.0 + s.matl[i].flux * 4.0));
                rough = 0.5f;
                break;
            case MATL_METAL: rough = (float)((s.matl[i].rough + (1.0 - s.matl[i].metal)) / 2.0); break;
            default: rough = 0.0f; break;
        }
        m[5] = rough;
    }
}

}  // namespace lumvox
