// glsl_shim.hpp -- TEST INFRASTRUCTURE (oracle/_ref recipe).
//
// A minimal GLSL-in-C++ environment so that the REFERENCE's own shader text (preprocessed by
// oracle/glsl_prep.py into oracle/_ref/*.inc, never committed) compiles with g++ and executes on
// the CPU.  The shader supplies every statement of the algorithm; this header supplies only what
// a Vulkan driver would: vector types, the built-in functions, texture/image access and the
// subgroup built-ins.  Where GLSL leaves precision or behaviour to the implementation, the shim
// follows the decrees of DESIGN.md section 3 (cited as D<n>), which are the same ones the C
// oracle (lum_oracle.c) and the CUDA kernels follow.
//
// Compile with -ffp-contract=off (D4: one rounding per written operation, no FMA).
#pragma once
#include <cmath>
#include <cstdint>
#include <type_traits>

namespace glsl {

typedef unsigned int uint;

template <class T> struct tvec2;
template <class T> struct tvec3;
template <class T> struct tvec4;

template <class From, class To>
inline constexpr bool implicit_ok = (std::is_same_v<To, float> && (std::is_same_v<From, int> || std::is_same_v<From, uint>)) ||
                                    (std::is_same_v<To, uint> && std::is_same_v<From, int>);

// GLSL % on signed operands is OpSMod (floor-mod) as glslang emits it
inline int  imod(int a, int b)   { int r = a % b; return (r != 0 && ((r < 0) != (b < 0))) ? r + b : r; }
inline uint imod(uint a, uint b) { return a % b; }

#define GLSL_VEC_OPS(V, N)                                                                             \
    friend V operator+(V a, V b) { V r; for (int i = 0; i < N; i++) r[i] = a[i] + b[i]; return r; }    \
    friend V operator-(V a, V b) { V r; for (int i = 0; i < N; i++) r[i] = a[i] - b[i]; return r; }    \
    friend V operator*(V a, V b) { V r; for (int i = 0; i < N; i++) r[i] = a[i] * b[i]; return r; }    \
    friend V operator/(V a, V b) { V r; for (int i = 0; i < N; i++) r[i] = a[i] / b[i]; return r; }    \
    friend V operator+(V a, T b) { V r; for (int i = 0; i < N; i++) r[i] = a[i] + b; return r; }       \
    friend V operator-(V a, T b) { V r; for (int i = 0; i < N; i++) r[i] = a[i] - b; return r; }       \
    friend V operator*(V a, T b) { V r; for (int i = 0; i < N; i++) r[i] = a[i] * b; return r; }       \
    friend V operator/(V a, T b) { V r; for (int i = 0; i < N; i++) r[i] = a[i] / b; return r; }       \
    friend V operator+(T a, V b) { V r; for (int i = 0; i < N; i++) r[i] = a + b[i]; return r; }       \
    friend V operator-(T a, V b) { V r; for (int i = 0; i < N; i++) r[i] = a - b[i]; return r; }       \
    friend V operator*(T a, V b) { V r; for (int i = 0; i < N; i++) r[i] = a * b[i]; return r; }       \
    friend V operator/(T a, V b) { V r; for (int i = 0; i < N; i++) r[i] = a / b[i]; return r; }       \
    friend V operator-(V a) { V r; for (int i = 0; i < N; i++) r[i] = -a[i]; return r; }               \
    V& operator+=(V b) { *this = *this + b; return *this; }                                            \
    V& operator-=(V b) { *this = *this - b; return *this; }                                            \
    V& operator*=(V b) { *this = *this * b; return *this; }                                            \
    V& operator/=(V b) { *this = *this / b; return *this; }                                            \
    V& operator*=(T b) { *this = *this * b; return *this; }                                            \
    V& operator/=(T b) { *this = *this / b; return *this; }

template <class T> struct tvec2 {
    union { T x, r; };
    union { T y, g; };
    tvec2() : x(), y() {}
    tvec2(T a, T b) : x(a), y(b) {}
    explicit tvec2(T s) : x(s), y(s) {}
    template <class U> explicit(!implicit_ok<U, T>) tvec2(const tvec2<U>& o) : x((T)o.x), y((T)o.y) {}
    template <class U> explicit tvec2(const tvec4<U>& o);            // vec2(gl_FragCoord): first two components
    friend tvec2 operator+(tvec2 a) { return a; }
    T& operator[](int i) { return i == 0 ? x : y; }
    T operator[](int i) const { return i == 0 ? x : y; }
    GLSL_VEC_OPS(tvec2, 2)
};

template <class T> struct tvec3 {
    union { T x, r; };
    union { T y, g; };
    union { T z, b; };
    tvec3() : x(), y(), z() {}
    tvec3(T a, T b_, T c) : x(a), y(b_), z(c) {}
    explicit tvec3(T s) : x(s), y(s), z(s) {}
    template <class U, class S> tvec3(const tvec2<U>& xy_, S c) : x((T)xy_.x), y((T)xy_.y), z((T)c) {}
    template <class U> explicit(!implicit_ok<U, T>) tvec3(const tvec3<U>& o) : x((T)o.x), y((T)o.y), z((T)o.z) {}
    T& operator[](int i) { return i == 0 ? x : (i == 1 ? y : z); }
    T operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
    tvec3 xyz() const { return *this; }
    tvec3 rgb() const { return *this; }
    tvec3 yzx() const { return tvec3(y, z, x); }
    tvec2<T> xy() const { return tvec2<T>(x, y); }
    GLSL_VEC_OPS(tvec3, 3)
    // integer-only operators (instantiated on use only)
    friend tvec3 operator%(tvec3 a, T b_) { tvec3 r_; for (int i = 0; i < 3; i++) r_[i] = imod(a[i], b_); return r_; }
    friend tvec3 operator>>(tvec3 a, T b_) { tvec3 r_; for (int i = 0; i < 3; i++) r_[i] = a[i] >> b_; return r_; }
    friend tvec3 operator^(tvec3 a, tvec3 b_) { tvec3 r_; for (int i = 0; i < 3; i++) r_[i] = a[i] ^ b_[i]; return r_; }
};

template <class T> struct tvec4 {
    union { T x, r; };
    union { T y, g; };
    union { T z, b; };
    union { T w, a; };
    tvec4() : x(), y(), z(), w() {}
    tvec4(T a_, T b_, T c, T d) : x(a_), y(b_), z(c), w(d) {}
    template <class S> tvec4(const tvec3<T>& v, S d) : x(v.x), y(v.y), z(v.z), w((T)d) {}
    T& operator[](int i) { return i == 0 ? x : (i == 1 ? y : (i == 2 ? z : w)); }
    T operator[](int i) const { return i == 0 ? x : (i == 1 ? y : (i == 2 ? z : w)); }
    tvec3<T> xyz() const { return tvec3<T>(x, y, z); }
    tvec3<T> rgb() const { return tvec3<T>(x, y, z); }
    tvec3<T> gba() const { return tvec3<T>(y, z, w); }
    tvec2<T> xy() const { return tvec2<T>(x, y); }
    GLSL_VEC_OPS(tvec4, 4)
};
template <class T> template <class U> tvec2<T>::tvec2(const tvec4<U>& o) : x((T)o.x), y((T)o.y) {}

typedef tvec2<float> vec2;   typedef tvec3<float> vec3;   typedef tvec4<float> vec4;
typedef tvec2<int> ivec2;    typedef tvec3<int> ivec3;    typedef tvec4<int> ivec4;
typedef tvec3<uint> uvec3;   typedef tvec4<uint> uvec4;
typedef tvec3<bool> bvec3;
// GLSL converts an unsigned vector implicitly when it meets a float (glossy.frag:369 `uvec3 / 255.0`)
inline tvec3<float> operator/(const tvec3<uint>& a, float b) { return tvec3<float>((float)a.x, (float)a.y, (float)a.z) / b; }

// Column-major 4x4 matrix.  mat4 * vec4 accumulates the columns in order 0,1,2,3, one rounding per
// product and per sum (D11).
struct mat4 {
    vec4 col[4];
    friend vec4 operator*(const mat4& m, vec4 v) { return ((m.col[0] * v.x + m.col[1] * v.y) + m.col[2] * v.z) + m.col[3] * v.w; }
};

// GL_EXT_shader_explicit_arithmetic_types i8vec4, as the request SSBO stores it (radiance.comp:23-25)
struct i8vec4 {
    int8_t x, y, z, w;
    ivec3 xyz() const { return ivec3(x, y, z); }
};

// ---- built-in functions (D4: IEEE binary32 sqrt and division, min/max as selects) ----
inline float sqrt(float v) { return ::sqrtf(v); }
inline float floor(float v) { return ::floorf(v); }
inline float abs(float v) { return ::fabsf(v); }
inline float cos(float v) { return ::cosf(v); }
inline float sin(float v) { return ::sinf(v); }
inline float sign(float v) { return v > 0.0f ? 1.0f : (v < 0.0f ? -1.0f : 0.0f); }
inline float min(float x, float y) { return y < x ? y : x; }
inline float max(float x, float y) { return x < y ? y : x; }
inline int min(int x, int y) { return y < x ? y : x; }
inline int max(int x, int y) { return x < y ? y : x; }
inline float clamp(float v, float lo, float hi) { return min(max(v, lo), hi); }
inline float clamp(float v, int lo, int hi) { return clamp(v, (float)lo, (float)hi); }     // clamp(x, 0, 1)
inline int clamp(int v, int lo, uint hi) { return min(max(v, lo), (int)hi); }
inline vec3 clamp(vec3 v, float lo, float hi) { return vec3(clamp(v.x, lo, hi), clamp(v.y, lo, hi), clamp(v.z, lo, hi)); }
inline vec3 floor(vec3 v) { return vec3(::floorf(v.x), ::floorf(v.y), ::floorf(v.z)); }
inline vec3 abs(vec3 v) { return vec3(::fabsf(v.x), ::fabsf(v.y), ::fabsf(v.z)); }
inline float dot(vec3 a, vec3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }           // x -> y -> z
inline vec3 normalize(vec3 v) { return v / sqrt(dot(v, v)); }
inline vec3 cross(vec3 a, vec3 b) { return vec3(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y); }
inline bvec3 lessThan(vec3 a, vec3 b) { return bvec3(a.x < b.x, a.y < b.y, a.z < b.z); }
inline bvec3 greaterThan(vec3 a, vec3 b) { return bvec3(a.x > b.x, a.y > b.y, a.z > b.z); }
inline bvec3 greaterThanEqual(vec3 a, vec3 b) { return bvec3(a.x >= b.x, a.y >= b.y, a.z >= b.z); }
inline bvec3 lessThan(ivec3 a, ivec3 b) { return bvec3(a.x < b.x, a.y < b.y, a.z < b.z); }
inline bvec3 greaterThanEqual(ivec3 a, ivec3 b) { return bvec3(a.x >= b.x, a.y >= b.y, a.z >= b.z); }
inline bool any(bvec3 v) { return v.x || v.y || v.z; }
inline float dot(vec2 a, vec2 b) { return a.x * b.x + a.y * b.y; }
// reflect(I, N) = I - 2.0 * dot(N, I) * N (GLSL 4.50 section 8.5), evaluated left to right (D14)
inline vec3 reflect(vec3 I, vec3 N) { return I - 2.0f * dot(N, I) * N; }

// ---- A2B10G10R10 unorm (internal_render.hpp:456; D7) ----
inline vec3 unpack_rgb10(uint32_t w) {
    return vec3((float)(w & 1023u) / 1023.0f, (float)((w >> 10) & 1023u) / 1023.0f, (float)((w >> 20) & 1023u) / 1023.0f);
}
inline uint32_t pack_unorm10(float v) {
    if (!(v == v)) return 0u;
    v = v < 0.0f ? 0.0f : (v > 1.0f ? 1.0f : v);
    return (uint32_t)::rintf(v * 1023.0f);
}
inline uint32_t pack_rgb10a2(vec4 c) {
    uint32_t a = (uint32_t)::rintf((c.w < 0.0f ? 0.0f : (c.w > 1.0f ? 1.0f : c.w)) * 3.0f);
    return pack_unorm10(c.x) | (pack_unorm10(c.y) << 10) | (pack_unorm10(c.z) << 20) | (a << 30);
}

// ---- resources.  Out-of-range texelFetch returns 0 (robust access, D2). ----
struct isampler3D {          // world grid R16_SINT, x + X*(y + Y*z) (internal_render.hpp:136-139)
    const int16_t* data = nullptr; ivec3 size; mutable uint64_t fetches = 0;
};
struct usampler3D {          // block palette atlas R8_UINT 1024 x 1024 x 16 (internal_render.cpp:137-145)
    const uint8_t* data = nullptr; ivec3 size;
};
struct sampler2D {           // material palette R32F 6 x 256 (internal_render.cpp:146-153)
    const float* data = nullptr; ivec2 size;
};
struct sampler3D {           // radiance field A2B10G10R10, LINEAR + REPEAT, normalised coords (setup.cpp:75-81)
    const uint32_t* data = nullptr; ivec3 size;
};
struct usampler2D {          // G-buffer material + normal, R8G8B8A8_UINT (internal_render.hpp:455), x + W*y
    const uint8_t* data = nullptr; ivec2 size;
};
struct image3D {
    uint32_t* data = nullptr; ivec3 size;
};
struct iimage3D { const int16_t* data = nullptr; ivec3 size; };      // r16i world grid (map.comp:16)
struct uimage3D { uint8_t* data = nullptr; ivec3 size; };            // r8ui palette atlas / model voxels
inline bool inside(ivec3 p, ivec3 s) { return p.x >= 0 && p.y >= 0 && p.z >= 0 && p.x < s.x && p.y < s.y && p.z < s.z; }
inline ivec4 texelFetch(const isampler3D& t, ivec3 p, int) {
    t.fetches++;
    if (!inside(p, t.size)) return ivec4(0, 0, 0, 0);
    return ivec4(t.data[p.x + (int64_t)t.size.x * (p.y + (int64_t)t.size.y * p.z)], 0, 0, 1);
}
inline uvec4 texelFetch(const usampler3D& t, ivec3 p, int) {
    if (!inside(p, t.size)) return uvec4(0, 0, 0, 0);
    return uvec4(t.data[p.x + (int64_t)t.size.x * (p.y + (int64_t)t.size.y * p.z)], 0, 0, 1);
}
inline vec4 texelFetch(const sampler2D& t, ivec2 p, int) {
    if (p.x < 0 || p.y < 0 || p.x >= t.size.x || p.y >= t.size.y) return vec4(0, 0, 0, 0);
    return vec4(t.data[p.x + t.size.x * p.y], 0, 0, 1);
}
inline uvec4 texelFetch(const usampler2D& t, ivec2 p, int) {
    if (p.x < 0 || p.y < 0 || p.x >= t.size.x || p.y >= t.size.y) return uvec4(0, 0, 0, 0);
    const uint8_t* q = t.data + 4 * (p.x + (int64_t)t.size.x * p.y);
    return uvec4(q[0], q[1], q[2], q[3]);
}
inline ivec2 textureSize(const usampler2D& t, int) { return t.size; }
// texture() on the NEAREST sampler the glossy pass binds for the depth attachment (internal_render.cpp:278,
// nearestSampler): texel floor(uv * size), clamped to the edge
inline vec4 texture(const sampler2D& t, vec2 uv) {
    int x = (int)::floorf(uv.x * (float)t.size.x), y = (int)::floorf(uv.y * (float)t.size.y);
    x = x < 0 ? 0 : (x >= t.size.x ? t.size.x - 1 : x);
    y = y < 0 ? 0 : (y >= t.size.y ? t.size.y - 1 : y);
    return vec4(t.data[x + t.size.x * y], 0, 0, 1);
}
inline vec4 texelFetch(const sampler3D& t, ivec3 p, int) {
    if (!inside(p, t.size)) return vec4(0, 0, 0, 0);
    return vec4(unpack_rgb10(t.data[p.x + (int64_t)t.size.x * (p.y + (int64_t)t.size.y * p.z)]), 1);
}
// Trilinear filter (D6): unnormalise u*size - 0.5, taps floor and floor+1 with REPEAT wrap,
// weights = fract, lerp x then y then z as a + w*(b - a).  The unnormalisation is evaluated in
// f32 here, as a driver would; lum_oracle.c uses the algebraically equal block-space coordinate
// directly, so the two can differ in the last bits of the weights (hence the tolerance in
// tests/test_ref_glsl.py).
inline vec4 textureLod(const sampler3D& t, vec3 uvw, float) {
    float c[3]; int i0[3], i1[3]; float w[3];
    for (int a = 0; a < 3; a++) {
        c[a] = uvw[a] * (float)t.size[a] - 0.5f;
        float f = ::floorf(c[a]);
        w[a] = c[a] - f;
        int n = t.size[a];
        int i = (int)f;
        i0[a] = imod(i, n);
        i1[a] = imod(i + 1, n);
    }
    auto tap = [&](int xi, int yi, int zi) { return unpack_rgb10(t.data[xi + (int64_t)t.size.x * (yi + (int64_t)t.size.y * zi)]); };
    auto lerp = [](vec3 a, vec3 b, float wt) { return a + wt * (b - a); };
    vec3 c00 = lerp(tap(i0[0], i0[1], i0[2]), tap(i1[0], i0[1], i0[2]), w[0]);
    vec3 c10 = lerp(tap(i0[0], i1[1], i0[2]), tap(i1[0], i1[1], i0[2]), w[0]);
    vec3 c01 = lerp(tap(i0[0], i0[1], i1[2]), tap(i1[0], i0[1], i1[2]), w[0]);
    vec3 c11 = lerp(tap(i0[0], i1[1], i1[2]), tap(i1[0], i1[1], i1[2]), w[0]);
    vec3 c0 = lerp(c00, c10, w[1]);
    vec3 c1 = lerp(c01, c11, w[1]);
    return vec4(lerp(c0, c1, w[2]), 1);
}
inline void imageStore(image3D& img, ivec3 p, vec4 c) {
    if (!inside(p, img.size)) return;
    img.data[p.x + (int64_t)img.size.x * (p.y + (int64_t)img.size.y * p.z)] = pack_rgb10a2(c);
}

inline ivec3 imageSize(const uimage3D& t) { return t.size; }
inline ivec4 imageLoad(const iimage3D& t, ivec3 p) {                  // out of range -> 0 (D2)
    if (!inside(p, t.size)) return ivec4(0, 0, 0, 0);
    return ivec4(t.data[p.x + (int64_t)t.size.x * (p.y + (int64_t)t.size.y * p.z)], 0, 0, 1);
}
inline uvec4 imageLoad(const uimage3D& t, ivec3 p) {
    if (!inside(p, t.size)) return uvec4(0, 0, 0, 0);
    return uvec4(t.data[p.x + (int64_t)t.size.x * (p.y + (int64_t)t.size.y * p.z)], 0, 0, 1);
}
inline void imageStore(uimage3D& t, ivec3 p, uvec4 c) {
    if (!inside(p, t.size)) return;
    t.data[p.x + (int64_t)t.size.x * (p.y + (int64_t)t.size.y * p.z)] = (uint8_t)c.x;
}

// ---- per-invocation built-ins and the subgroup model ----
// Invocations of a workgroup run one after another.  subgroupAdd() is resolved in two passes:
// pass 0 runs every invocation up to its subgroupAdd() call, records the operand and unwinds;
// the driver forms the subgroup sums with the pinned butterfly (D3); pass 1 re-runs the elected
// invocations with subgroupAdd() returning that sum (the shader is deterministic, so the re-run
// retraces the same ray).  barrier() is a no-op because the driver orders the elected lanes so
// that every _light[] slot is written before subgroup 0 reads them.
struct SubgroupYield {};
struct ShaderBase {
    uvec3 gl_GlobalInvocationID;
    vec4 gl_FragCoord;                // fragment shaders (glossy.frag)
    uint gl_SubgroupSize = 32;
    uint gl_SubgroupID = 0;
    uint gl_SubgroupInvocationID = 0;
    ivec3 world_size;                 // shaders/common/consts.glsl:1 (a constant in the reference)
    int pass = 0;
    vec3 subgroup_operand;            // pass 0 out
    vec3 subgroup_sum;                // pass 1 in
    vec3 mix_y;                       // last mix() second operand = pre-blend mean (debug tap)
    vec3 subgroupAdd(vec3 v) {
        if (pass == 0) { subgroup_operand = v; throw SubgroupYield(); }
        return subgroup_sum;
    }
    bool subgroupElect() const { return gl_SubgroupInvocationID == 0; }
    void barrier() const {}
    vec3 mix(vec3 a, vec3 b, float t) { mix_y = b; return a * (1.0f - t) + b * t; }   // D4
};

}  // namespace glsl
