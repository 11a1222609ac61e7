// host_math.hpp -- the few linear-algebra routines the host side of the radiance path needs
// (the reference takes them from glm: toMat4, translate, inverse, lookAt, ortho, quat*vec3).
// Written from the textbook definitions with glm's conventions (column-major, right-handed,
// clip depth -1..1) so results agree with the reference's host code to rounding.
#pragma once
#include <cmath>
#include <cstdint>

#include "lum_math.hpp"

namespace lumhost {
using namespace lum_math;

inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
inline int32_t f2i(float f) {                       // C++ float -> int conversion, made total
    if (!(f == f)) return 0;
    if (f >= 2147483648.0f) return INT32_MAX;
    if (f <= -2147483648.0f) return INT32_MIN;
    return (int32_t)f;
}

inline dvec3 cross(const dvec3& a, const dvec3& b) { return dvec3(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y); }
inline double dot(const dvec3& a, const dvec3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline dvec3 normalize(const dvec3& a) { const double l = std::sqrt(dot(a, a)); return dvec3(a.x / l, a.y / l, a.z / l); }

inline dvec4 mul(const dmat4& m, const dvec4& v) {
    dvec4 r;
    for (int k = 0; k < 4; k++) r[k] = (m[0][k] * v.x + m[1][k] * v.y) + (m[2][k] * v.z + m[3][k] * v.w);
    return r;
}
inline dmat4 mul(const dmat4& a, const dmat4& b) {
    dmat4 r;
    for (int c = 0; c < 4; c++) {
        const dvec4 col = mul(a, b[c]);
        r[c] = col;
    }
    return r;
}

// rotation matrix of a unit quaternion (float arithmetic, as toMat4 on a float quat)
inline void quat_to_mat3(const quat& q, float m[3][3]) {
    const float xx = q.x * q.x, yy = q.y * q.y, zz = q.z * q.z;
    const float xz = q.x * q.z, xy = q.x * q.y, yz = q.y * q.z;
    const float wx = q.w * q.x, wy = q.w * q.y, wz = q.w * q.z;
    m[0][0] = 1.0f - 2.0f * (yy + zz); m[0][1] = 2.0f * (xy + wz);        m[0][2] = 2.0f * (xz - wy);
    m[1][0] = 2.0f * (xy - wz);        m[1][1] = 1.0f - 2.0f * (xx + zz); m[1][2] = 2.0f * (yz + wx);
    m[2][0] = 2.0f * (xz + wy);        m[2][1] = 2.0f * (yz - wx);        m[2][2] = 1.0f - 2.0f * (xx + yy);
}

// translate(identity, shift) * toMat4(rot) = [R | shift] (exact: the product only adds zeros)
inline dmat4 model_transform(const quat& rot, const vec3& shift) {
    float r[3][3];
    quat_to_mat3(rot, r);
    dmat4 m;
    for (int c = 0; c < 3; c++) m[c] = dvec4((double)r[c][0], (double)r[c][1], (double)r[c][2], 0.0);
    m[3] = dvec4((double)shift.x, (double)shift.y, (double)shift.z, 1.0);
    return m;
}

// mat4(inverse(m)) as internal_render.cpp:1510 gets it from glm, each entry rounded to float; out is column-major.
inline void inverse_to_float(const dmat4& a, float out[16]) {
    // glm::inverse(dmat4) as the vendored glm computes it (external/lum-al/external/glm/glm/detail/func_matrix.inl:388-447,
    // compute_inverse<4,4>): the eighteen 2x2 sub-determinants, cofactor columns signed by a multiplication with +-1
    // (which is where its -0.0 entries come from), the determinant as the dot of the first row, every entry times
    // 1/det.  m[col][row].
    const dmat4& m = a;
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

// right-handed look-at view matrix
inline dmat4 look_at(const dvec3& eye, const dvec3& center, const dvec3& up) {
    const dvec3 f = normalize(dvec3(center.x - eye.x, center.y - eye.y, center.z - eye.z));
    const dvec3 s = normalize(cross(f, up));
    const dvec3 u = cross(s, f);
    dmat4 m;
    m[0] = dvec4(s.x, u.x, -f.x, 0.0);
    m[1] = dvec4(s.y, u.y, -f.y, 0.0);
    m[2] = dvec4(s.z, u.z, -f.z, 0.0);
    m[3] = dvec4(-dot(s, eye), -dot(u, eye), dot(f, eye), 1.0);
    return m;
}

// right-handed orthographic projection, clip depth -1..1
inline dmat4 ortho(double l, double r, double b, double t, double n, double f) {
    dmat4 m;
    m[0] = dvec4(2.0 / (r - l), 0, 0, 0);
    m[1] = dvec4(0, 2.0 / (t - b), 0, 0);
    m[2] = dvec4(0, 0, -2.0 / (f - n), 0);
    m[3] = dvec4(-(r + l) / (r - l), -(t + b) / (t - b), -(f + n) / (f - n), 1.0);
    return m;
}

// q * v for a unit quaternion: v + 2*(w*(qv x v) + qv x (qv x v))
inline vec3 rotate(const quat& q, const vec3& v) {
    const vec3 qv(q.x, q.y, q.z);
    const vec3 uv(qv.y * v.z - v.y * qv.z, qv.z * v.x - v.z * qv.x, qv.x * v.y - v.x * qv.y);
    const vec3 uuv(qv.y * uv.z - uv.y * qv.z, qv.z * uv.x - uv.z * qv.x, qv.x * uv.y - uv.x * qv.y);
    return vec3(v.x + ((uv.x * q.w) + uuv.x) * 2.0f, v.y + ((uv.y * q.w) + uuv.y) * 2.0f, v.z + ((uv.z * q.w) + uuv.z) * 2.0f);
}

}  // namespace lumhost
