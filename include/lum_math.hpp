// lum_math.hpp -- the handful of vector types Lum's public API exposes (the reference takes them
// from glm: src/renderer/src/internal_types.hpp:3-8).  Layout-compatible with glm's defaults
// (vecN = N packed scalars; quat stored x,y,z,w and constructed (w,x,y,z); matrices column-major),
// so a Lum application that already uses glm can define LUM_B200_USE_GLM and pass its own types.
#pragma once
#ifndef LUM_B200_MATH_HPP
#define LUM_B200_MATH_HPP
#include <cstdint>

#ifdef LUM_B200_USE_GLM
#define GLM_ENABLE_EXPERIMENTAL
#include <glm/glm.hpp>
#include <glm/gtc/type_precision.hpp>
#include <glm/gtx/quaternion.hpp>
namespace lum_math {
using glm::vec2; using glm::vec3; using glm::vec4; using glm::ivec2; using glm::ivec3; using glm::ivec4;
using glm::uvec3; using glm::dvec2; using glm::dvec3; using glm::dvec4; using glm::dmat4; using glm::mat4;
using glm::quat; using glm::i8vec4; using glm::i16vec4; using glm::u8vec3;
}
#else
namespace lum_math {

template <typename T> struct tvec2 {
    T x{}, y{};
    constexpr tvec2() = default;
    constexpr tvec2(T a, T b) : x(a), y(b) {}
    constexpr explicit tvec2(T s) : x(s), y(s) {}
    T& operator[](int i) { return (&x)[i]; }
    const T& operator[](int i) const { return (&x)[i]; }
};
template <typename T> struct tvec3 {
    T x{}, y{}, z{};
    constexpr tvec3() = default;
    constexpr tvec3(T a, T b, T c) : x(a), y(b), z(c) {}
    constexpr explicit tvec3(T s) : x(s), y(s), z(s) {}
    template <typename U> constexpr tvec3(const tvec3<U>& o) : x((T)o.x), y((T)o.y), z((T)o.z) {}   // glm: implicit
    T& operator[](int i) { return (&x)[i]; }
    const T& operator[](int i) const { return (&x)[i]; }
    tvec3& operator+=(const tvec3& o) { x += o.x; y += o.y; z += o.z; return *this; }
    tvec3& operator-=(const tvec3& o) { x -= o.x; y -= o.y; z -= o.z; return *this; }
};
template <typename T> constexpr tvec3<T> operator+(tvec3<T> a, tvec3<T> b) { return {T(a.x + b.x), T(a.y + b.y), T(a.z + b.z)}; }
template <typename T> constexpr tvec3<T> operator-(tvec3<T> a, tvec3<T> b) { return {T(a.x - b.x), T(a.y - b.y), T(a.z - b.z)}; }
template <typename T> constexpr tvec3<T> operator*(tvec3<T> a, T s) { return {T(a.x * s), T(a.y * s), T(a.z * s)}; }
template <typename T> constexpr tvec3<T> operator*(T s, tvec3<T> a) { return {T(a.x * s), T(a.y * s), T(a.z * s)}; }
template <typename T> constexpr tvec3<T> operator-(tvec3<T> a) { return {T(-a.x), T(-a.y), T(-a.z)}; }
template <typename T> constexpr bool operator==(tvec3<T> a, tvec3<T> b) { return a.x == b.x && a.y == b.y && a.z == b.z; }

template <typename T> struct tvec4 {
    T x{}, y{}, z{}, w{};
    constexpr tvec4() = default;
    constexpr tvec4(T a, T b, T c, T d) : x(a), y(b), z(c), w(d) {}
    constexpr explicit tvec4(T s) : x(s), y(s), z(s), w(s) {}
    template <typename U, typename V> constexpr tvec4(const tvec3<U>& v, V d) : x((T)v.x), y((T)v.y), z((T)v.z), w((T)d) {}
    T& operator[](int i) { return (&x)[i]; }
    const T& operator[](int i) const { return (&x)[i]; }
};

// quaternion: memory order x,y,z,w; constructor order (w,x,y,z) -- glm's default convention
struct quat {
    float x = 0.0f, y = 0.0f, z = 0.0f, w = 1.0f;
    constexpr quat() = default;
    constexpr quat(float w_, float x_, float y_, float z_) : x(x_), y(y_), z(z_), w(w_) {}
};

template <typename T> struct tmat4 {
    tvec4<T> c[4];                                  // columns
    tvec4<T>& operator[](int i) { return c[i]; }
    const tvec4<T>& operator[](int i) const { return c[i]; }
};

using vec2 = tvec2<float>;   using vec3 = tvec3<float>;   using vec4 = tvec4<float>;
using ivec2 = tvec2<int32_t>; using ivec3 = tvec3<int32_t>; using ivec4 = tvec4<int32_t>;
using uvec3 = tvec3<uint32_t>;
using dvec2 = tvec2<double>; using dvec3 = tvec3<double>; using dvec4 = tvec4<double>;
using mat4 = tmat4<float>;   using dmat4 = tmat4<double>;
using i8vec4 = tvec4<int8_t>; using i16vec4 = tvec4<int16_t>; using u8vec3 = tvec3<uint8_t>;

}  // namespace lum_math
#endif  // LUM_B200_USE_GLM
#endif
