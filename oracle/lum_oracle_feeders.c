/*
 * lum_oracle_feeders.c -- CPU restatement (ORACLE) of the feeders of Lum's radiance pass:
 * model placement, blockify, the brick copies of exec_copies, and shaders/map.comp.
 *
 * TEST INFRASTRUCTURE ONLY (see lum_oracle.h).  map.comp is PINNED against the reference's own
 * shader text executed on the CPU (oracle/_ref/libref_map_glsl.so, tests/test_ref_glsl.py);
 * placement / blockify are PINNED against the reference's own host C++ -- get_shift, blockify_mesh and
 * map_mesh's arithmetic cut out of internal_render.cpp and compiled in place (oracle/_ref/libref_host.so,
 * tests/test_ref_host.py, tests/golden/ref_host_golden.npz); the brick copies follow from blockify's pairs.
 * Each function follows the statement order of the reference text it cites (paths relative to
 * /root/reference).  Arithmetic rules: as lum_oracle.c (binary32/binary64 exactly as the reference
 * types them, one rounding per written operation, no FMA contraction).
 */
#include "lum_oracle.h"
#include <math.h>
#include <string.h>

static inline int32_t f2i(float f) {                    /* ivec3(vec3): trunc toward zero, saturating */
    if (!(f == f)) return 0;
    if (f >= 2147483648.0f) return INT32_MAX;
    if (f <= -2147483648.0f) return INT32_MIN;
    return (int32_t)f;
}
static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* glm::mat3_cast / toMat4 (quaternion -> rotation matrix), float, column-major m[col][row].
 * rot = (w, x, y, z) as the C API passes it (crenderer.cpp:139-145). */
static void quat_to_mat(const float q[4], float m[3][3]) {
    const float w = q[0], x = q[1], y = q[2], z = q[3];
    const float xx = x * x, yy = y * y, zz = z * z;
    const float xz = x * z, xy = x * y, yz = y * z;
    const float wx = w * x, wy = w * y, wz = w * z;
    m[0][0] = 1.0f - 2.0f * (yy + zz); m[0][1] = 2.0f * (xy + wz);        m[0][2] = 2.0f * (xz - wy);
    m[1][0] = 2.0f * (xy - wz);        m[1][1] = 1.0f - 2.0f * (xx + zz); m[1][2] = 2.0f * (yz + wx);
    m[2][0] = 2.0f * (xz + wy);        m[2][1] = 2.0f * (yz - wx);        m[2][2] = 1.0f - 2.0f * (xx + yy);
}

/* mat4(inverse(transform)) (internal_render.cpp:1510): glm's double 4x4 inverse, each entry rounded to float.  a[col][row]. */
static void inverse4(const double a[4][4], float out[16]) {
    /* glm::inverse(dmat4) as the vendored glm computes it (external/lum-al/external/glm/glm/detail/func_matrix.inl:388-447,
     * compute_inverse<4,4>): the eighteen 2x2 sub-determinants, cofactor columns signed by a multiplication with +-1
     * (which is where its -0.0 entries come from), the determinant as the dot of the first row, every entry times
     * 1/det.  m[col][row]. */
    const double (*m)[4] = a;
    const double Coef00 = m[2][2] * m[3][3] - m[3][2] * m[2][3];
    const double Coef02 = m[1][2] * m[3][3] - m[3][2] * m[1][3];
    const double Coef03 = m[1][2] * m[2][3] - m[2][2] * m[1][3];
    const double Coef04 = m[2][1] * m[3][3] - m[3][1] * m[2][3];
    const double Coef06 = m[1][1] * m[3][3] - m[3][1] * m[1][3];
    const double Coef07 = m[1][1] * m[2][3] - m[2][1] * m[1][3];
    const double Coef08 = m[2][1] * m[3][2] - m[3][1] * m[2][2];
    const double Coef10 = m[1][1] * m[3][2] - m[3][1] * m[1][2];
    const double Coef11 = m[1][1] * m[2][2] - m[2][1] * m[1][2];
    const double Coef12 = m[2][0] * m[3][3] - m[3][0] * m[2][3];
    const double Coef14 = m[1][0] * m[3][3] - m[3][0] * m[1][3];
    const double Coef15 = m[1][0] * m[2][3] - m[2][0] * m[1][3];
    const double Coef16 = m[2][0] * m[3][2] - m[3][0] * m[2][2];
    const double Coef18 = m[1][0] * m[3][2] - m[3][0] * m[1][2];
    const double Coef19 = m[1][0] * m[2][2] - m[2][0] * m[1][2];
    const double Coef20 = m[2][0] * m[3][1] - m[3][0] * m[2][1];
    const double Coef22 = m[1][0] * m[3][1] - m[3][0] * m[1][1];
    const double Coef23 = m[1][0] * m[2][1] - m[2][0] * m[1][1];
    const double Fac0[4] = {Coef00, Coef00, Coef02, Coef03}, Fac1[4] = {Coef04, Coef04, Coef06, Coef07};
    const double Fac2[4] = {Coef08, Coef08, Coef10, Coef11}, Fac3[4] = {Coef12, Coef12, Coef14, Coef15};
    const double Fac4[4] = {Coef16, Coef16, Coef18, Coef19}, Fac5[4] = {Coef20, Coef20, Coef22, Coef23};
    const double Vec0[4] = {m[1][0], m[0][0], m[0][0], m[0][0]}, Vec1[4] = {m[1][1], m[0][1], m[0][1], m[0][1]};
    const double Vec2[4] = {m[1][2], m[0][2], m[0][2], m[0][2]}, Vec3[4] = {m[1][3], m[0][3], m[0][3], m[0][3]};
    const double SignA[4] = {+1.0, -1.0, +1.0, -1.0}, SignB[4] = {-1.0, +1.0, -1.0, +1.0};
    double inv[4][4];
    for (int k = 0; k < 4; k++) {
        inv[0][k] = ((Vec1[k] * Fac0[k] - Vec2[k] * Fac1[k]) + Vec3[k] * Fac2[k]) * SignA[k];
        inv[1][k] = ((Vec0[k] * Fac0[k] - Vec2[k] * Fac3[k]) + Vec3[k] * Fac4[k]) * SignB[k];
        inv[2][k] = ((Vec0[k] * Fac1[k] - Vec1[k] * Fac3[k]) + Vec3[k] * Fac5[k]) * SignA[k];
        inv[3][k] = ((Vec0[k] * Fac2[k] - Vec1[k] * Fac4[k]) + Vec2[k] * Fac5[k]) * SignB[k];
    }
    const double Dot1 = (m[0][0] * inv[0][0] + m[0][1] * inv[1][0]) + (m[0][2] * inv[2][0] + m[0][3] * inv[3][0]);
    const double OneOverDeterminant = 1.0 / Dot1;
    for (int c = 0; c < 4; c++) for (int r = 0; r < 4; r++) out[4 * c + r] = (float)(inv[c][r] * OneOverDeterminant);
}

/* get_shift (internal_render.cpp:989-1016) + the border arithmetic of blockify_mesh (:1031-1041)
 * and map_mesh (:1499-1513). */
void lum_oracle_place_model(const float rot_wxyz[4], const float shift[3], const int32_t size[3],
                            const int32_t world_size[3], LumOraclePlacement *out) {
    float r[3][3];
    quat_to_mat(rot_wxyz, r);                                            /* toMat4(trans.rot) :1030,1499 */
    /* shift * rotate: translate(identity, shift) * rotation = [R | shift] exactly (:1031,1501) */
    double a[4][4];
    for (int c = 0; c < 3; c++) { for (int k = 0; k < 3; k++) a[c][k] = (double)r[c][k]; a[c][3] = 0.0; }
    a[3][0] = (double)shift[0]; a[3][1] = (double)shift[1]; a[3][2] = (double)shift[2]; a[3][3] = 1.0;
    /* get_shift: 8 corners of box = size-1, transformed in double, stored as float (:990-1013) */
    const float bx = (float)(size[0] - 1), by = (float)(size[1] - 1), bz = (float)(size[2] - 1);
    const float corners[8][3] = { {0, 0, 0}, {0, by, 0}, {0, by, bz}, {0, 0, bz},
                                  {bx, 0, 0}, {bx, by, 0}, {bx, by, bz}, {bx, 0, bz} };
    float tmin[3] = { 3.402823466e+38f, 3.402823466e+38f, 3.402823466e+38f };
    float tmax[3] = { -3.402823466e+38f, -3.402823466e+38f, -3.402823466e+38f };
    for (int i = 0; i < 8; i++) {
        for (int k = 0; k < 3; k++) {
            /* (m[0]*v.x + m[1]*v.y) + (m[2]*v.z + m[3]*1) */
            const double p = (a[0][k] * (double)corners[i][0] + a[1][k] * (double)corners[i][1]) +
                             (a[2][k] * (double)corners[i][2] + a[3][k] * 1.0);
            const float pf = (float)p;
            tmin[k] = (pf < tmin[k]) ? pf : tmin[k];                     /* glm::min(tmin, point) */
            tmax[k] = (tmax[k] < pf) ? pf : tmax[k];                     /* glm::max(tmax, point) */
        }
    }
    for (int k = 0; k < 3; k++) {
        out->aabb_min[k] = tmin[k]; out->aabb_max[k] = tmax[k];
        /* blockify: (ivec3(min) - 1) / 16, (ivec3(max) + 1) / 16, clamped to the world (:1034-1037) */
        out->block_min[k] = clampi((f2i(tmin[k]) - 1) / 16, 0, world_size[k] - 1);
        out->block_max[k] = clampi((f2i(tmax[k]) + 1) / 16, 0, world_size[k] - 1);
        /* map: ivec3(floor(min)), ivec3(ceil(max)) (:1506-1507) */
        out->map_shift[k] = f2i(floorf(tmin[k]));
        out->map_area[k] = f2i(ceilf(tmax[k])) - out->map_shift[k];
    }
    inverse4(a, out->inverse_trans);                                     /* :1513 */
}

/* start_blockify + blockify_mesh for every model in draw order (internal_render.cpp:977-987,
 * 1029-1076).  copy_src/copy_dst receive one pair per allocated dynamic block (src may be 0 =
 * "zero the brick": the reference clears the atlas instead, :1402).  Returns the number of
 * allocated blocks; nothing is written past copy_cap. */
int32_t lum_oracle_blockify(int wx, int wy, int wz, const int16_t *origin_world, int16_t *current_world,
                            int static_count, const LumOraclePlacement *models, int nmodels,
                            int32_t *copy_src, int32_t *copy_dst, int32_t copy_cap) {
    memcpy(current_world, origin_world, (size_t)wx * wy * wz * sizeof(int16_t));   /* :985 */
    int palette_counter = static_count;                                            /* :982 */
    int32_t n = 0;
    for (int m = 0; m < nmodels; m++) {
        const LumOraclePlacement *p = &models[m];
        for (int zz = p->block_min[2]; zz <= p->block_max[2]; zz++)
        for (int yy = p->block_min[1]; yy <= p->block_max[1]; yy++)
        for (int xx = p->block_min[0]; xx <= p->block_max[0]; xx++) {
            int16_t *cell = &current_world[(size_t)xx + (size_t)wx * ((size_t)yy + (size_t)wy * (size_t)zz)];
            const int cur = *cell;
            if (cur < static_count) {                                              /* :1044 */
                if (n < copy_cap) { copy_src[n] = cur; copy_dst[n] = palette_counter; }
                n++;
                *cell = (int16_t)palette_counter;                                  /* :1068 */
                palette_counter++;                                                 /* :1069 */
            }
        }
    }
    return n;
}

/* The per-block copies of exec_copies (internal_render.cpp:1402-1446). */
void lum_oracle_copy_bricks(uint8_t *bricks, int capacity, const int32_t *src, const int32_t *dst, int32_t n) {
    for (int32_t i = 0; i < n; i++) {
        if (dst[i] < 0 || dst[i] >= capacity) continue;
        if (src[i] > 0 && src[i] < capacity) memmove(bricks + (size_t)dst[i] * 4096u, bricks + (size_t)src[i] * 4096u, 4096u);
        else memset(bricks + (size_t)dst[i] * 4096u, 0, 4096u);
    }
}

/* shaders/map.comp:32-67 over the invocations of vkCmdDispatch((area+3)/4 ...) x local size 4^3
 * (internal_render.cpp:1517). */
void lum_oracle_map_model(int wx, int wy, int wz, const int16_t *world, uint8_t *bricks, int capacity,
                          const uint8_t *model, int sx, int sy, int sz,
                          const float m[16], const int32_t shift[3], const int32_t area[3], int static_count) {
    int ext[3];
    for (int k = 0; k < 3; k++) { int g = (area[k] + 3) / 4; ext[k] = g > 0 ? g * 4 : 0; }
    for (int gz = 0; gz < ext[2]; gz++)
    for (int gy = 0; gy < ext[1]; gy++)
    for (int gx = 0; gx < ext[0]; gx++) {
        const float px = ((float)gx / 1.0f + 0.5f) + (float)shift[0];    /* :41-43 */
        const float py = ((float)gy / 1.0f + 0.5f) + (float)shift[1];
        const float pz = ((float)gz / 1.0f + 0.5f) + (float)shift[2];
        /* decree 11: an invocation whose world voxel lies below the world (a negative component)
         * is discarded.  In the reference its truncating ivec3()/"/16" alias it onto a voxel of a
         * low-edge block, where it races with the legitimate invocation for that voxel. */
        if (px < 0.0f || py < 0.0f || pz < 0.0f) continue;
        const int ix = f2i(px), iy = f2i(py), iz = f2i(pz);             /* :44 */
        /* :50  inverse_trans * vec4(world_voxel, 1): columns accumulated in order 0,1,2,3 */
        const float mx = ((m[0] * px + m[4] * py) + m[8] * pz) + m[12] * 1.0f;
        const float my = ((m[1] * px + m[5] * py) + m[9] * pz) + m[13] * 1.0f;
        const float mz = ((m[2] * px + m[6] * py) + m[10] * pz) + m[14] * 1.0f;
        const int jx = f2i(mx), jy = f2i(my), jz = f2i(mz);             /* :51 */
        const int bx = ix / 16, by = iy / 16, bz = iz / 16;             /* :53 (OpSDiv truncates) */
        int id = 0;                                                     /* :55 out of range -> 0 */
        if (bx >= 0 && bx < wx && by >= 0 && by < wy && bz >= 0 && bz < wz)
            id = world[(size_t)bx + (size_t)wx * ((size_t)by + (size_t)wy * (size_t)bz)];
        if (id >= static_count) {                                       /* :62 */
            int v = 0;                                                  /* :63 out of range -> 0 */
            if (jx >= 0 && jx < sx && jy >= 0 && jy < sy && jz >= 0 && jz < sz)
                v = model[(size_t)jx + (size_t)sx * ((size_t)jy + (size_t)sy * (size_t)jz)];
            if (v != 0 && id < capacity)                                /* :64-65; % 16 is floor-mod (OpSMod) */
                bricks[(size_t)id * 4096u + (size_t)(ix & 15) * 256u + (size_t)(iy & 15) * 16u + (size_t)(iz & 15)] = (uint8_t)v;
        }
    }
}
