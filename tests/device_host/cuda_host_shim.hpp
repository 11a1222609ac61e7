// cuda_host_shim.hpp -- TEST INFRASTRUCTURE ONLY.
// Lets g++ compile lum_b200/csrc/radiance_device.cuh (the per-thread arithmetic of the CUDA kernels) as plain
// host C++, so that the CPU-side test-suite can run the *device source itself* against the oracle where there is
// no GPU (tests/test_device_source_on_host.py).  Nothing under lum_b200/ includes or loads this; the product has
// no CPU path.  Each intrinsic below is the IEEE / PTX definition of the CUDA intrinsic of the same name.
#pragma once
#include <cuda_runtime.h>   // vector types (float4, uint3, short4) and empty __device__/__forceinline__ for g++
#include <cmath>
#include <cstdint>
#include <cstring>

template <class T> static inline T __ldg(const T *p) { return *p; }
static inline unsigned __float_as_uint(float f) { unsigned u; std::memcpy(&u, &f, 4); return u; }
static inline float __uint_as_float(unsigned u) { float f; std::memcpy(&f, &u, 4); return f; }
static inline float __int_as_float(int i) { float f; std::memcpy(&f, &i, 4); return f; }
// cvt.rzi.s32.f32: truncate, saturate, NaN -> 0
static inline int __float2int_rz(float f) {
    if (!(f == f)) return 0;
    if (f >= 2147483648.0f) return 2147483647;
    if (f <= -2147483648.0f) return (int)0x80000000;
    return (int)f;
}
// cvt.rni.s32.f32: round to nearest even, saturate, NaN -> 0 (the default rounding mode is never changed here)
static inline int __float2int_rn(float f) {
    if (!(f == f)) return 0;
    if (f >= 2147483648.0f) return 2147483647;
    if (f <= -2147483648.0f) return (int)0x80000000;
    return (int)std::nearbyintf(f);
}
static inline float __fmaf_rn(float a, float b, float c) { return std::fmaf(a, b, c); }
#ifndef LUMCU_HOST_EMU
static inline void __syncwarp() {}     // one thread at a time
#endif
