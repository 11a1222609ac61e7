// simt_emu.hpp -- TEST INFRASTRUCTURE ONLY.
// Runs the __global__ kernels of lum_b200/csrc/radiance_kernels.cuh on the host, one OS thread per CUDA thread, one CTA
// at a time: threadIdx / blockIdx are thread-local, __shared__ becomes `static` (one CTA is resident at a time), the warp
// collectives (__ballot_sync, __shfl_*_sync) are rendezvous of the warp's threads, __syncthreads and the named barrier are
// counting barriers, atomics are the compiler's.  It is slow (a barrier per collective) and exists so that the
// ORCHESTRATION of the kernels -- ray pool, refill, parking, tail passes, reduction -- can be checked against the oracle
// in the CPU-side test-suite.  Nothing under lum_b200/ includes it.  Restrictions honoured by the kernels it runs: the
// threads of a warp reach every collective together, and a thread that returns early does so with its whole CTA.
#pragma once
#define LUMCU_HOST_EMU 1
#include "cuda_host_shim.hpp"

#include <algorithm>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <vector>

#undef __shared__
#define __shared__ static
#undef __launch_bounds__
#define __launch_bounds__(...)

struct EmuBarrier {
    std::mutex m;
    std::condition_variable cv;
    int count = 0;
    unsigned gen = 0;
    void arrive_and_wait(int n) {
        std::unique_lock<std::mutex> l(m);
        const unsigned g = gen;
        if (++count == n) { count = 0; gen++; cv.notify_all(); }
        else cv.wait(l, [&] { return gen != g; });
    }
};
struct EmuWarp { EmuBarrier bar; unsigned long long slot[32]; int size = 32; };
struct EmuCta {
    EmuBarrier cta_bar, named[16];
    EmuWarp warps[32];
    int nthreads = 0;
    int red = 0;
};

static thread_local uint3 threadIdx, blockIdx;
static thread_local dim3 blockDim, gridDim;
static thread_local EmuCta *emu_cta = nullptr;

// ---- warp collectives ----------------------------------------------------------------------------------------------
static inline void emu_exchange(unsigned long long v, unsigned long long out[32], int &size) {
    EmuWarp &w = emu_cta->warps[threadIdx.x >> 5];
    size = w.size;
    w.slot[threadIdx.x & 31] = v;
    w.bar.arrive_and_wait(w.size);
    for (int i = 0; i < w.size; i++) out[i] = w.slot[i];
    w.bar.arrive_and_wait(w.size);
}
template <class T> static inline unsigned long long emu_bits(T v) { unsigned long long u = 0; std::memcpy(&u, &v, sizeof v); return u; }
template <class T> static inline T emu_from(unsigned long long u) { T v; std::memcpy(&v, &u, sizeof v); return v; }

static inline unsigned __ballot_sync(unsigned, int pred) {
    unsigned long long all[32]; int n;
    emu_exchange(pred ? 1ull : 0ull, all, n);
    unsigned b = 0;
    for (int i = 0; i < n; i++) b |= (unsigned)(all[i] & 1ull) << i;
    return b;
}
template <class T> static inline T __shfl_sync(unsigned, T v, int src) {
    unsigned long long all[32]; int n;
    emu_exchange(emu_bits(v), all, n);
    return (src >= 0 && src < n) ? emu_from<T>(all[src]) : v;
}
template <class T> static inline T __shfl_xor_sync(unsigned, T v, int k) {
    unsigned long long all[32]; int n;
    emu_exchange(emu_bits(v), all, n);
    const int src = (int)(threadIdx.x & 31) ^ k;
    return (src < n) ? emu_from<T>(all[src]) : v;
}
template <class T> static inline T __shfl_up_sync(unsigned, T v, int k) {
    unsigned long long all[32]; int n;
    emu_exchange(emu_bits(v), all, n);
    const int src = (int)(threadIdx.x & 31) - k;
    return (src >= 0) ? emu_from<T>(all[src]) : v;
}
static inline void __syncwarp(unsigned = 0xffffffffu) {
    EmuWarp &w = emu_cta->warps[threadIdx.x >> 5];
    w.bar.arrive_and_wait(w.size);
}
static inline int __popc(unsigned v) { return __builtin_popcount(v); }

// ---- CTA barriers --------------------------------------------------------------------------------------------------
static inline void __syncthreads() { emu_cta->cta_bar.arrive_and_wait(emu_cta->nthreads); }
static inline int __syncthreads_count(int pred) {
    if (pred) __atomic_fetch_add(&emu_cta->red, 1, __ATOMIC_SEQ_CST);
    __syncthreads();
    const int r = __atomic_load_n(&emu_cta->red, __ATOMIC_SEQ_CST);
    __syncthreads();
    if (threadIdx.x == 0) emu_cta->red = 0;
    __syncthreads();
    return r;
}
static inline int __syncthreads_or(int pred) { return __syncthreads_count(pred) != 0; }
static inline void lumcu_host_named_barrier(int id, int count) { emu_cta->named[id].arrive_and_wait(count); }

// ---- atomics, fences, clocks ---------------------------------------------------------------------------------------
static inline int atomicAdd(int *p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline unsigned atomicAdd(unsigned *p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline unsigned atomicOr(unsigned *p, unsigned v) { return __atomic_fetch_or(p, v, __ATOMIC_SEQ_CST); }
static inline long long atomicMin(long long *p, long long v) {
    long long o = __atomic_load_n(p, __ATOMIC_SEQ_CST);
    while (v < o && !__atomic_compare_exchange_n(p, &o, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
    return o;
}
static inline void __threadfence_system() { __sync_synchronize(); }
static inline void __threadfence() { __sync_synchronize(); }
static inline long long clock64() { return 0; }
static inline void __nanosleep(unsigned) {}
static inline int __ffs(int x) { return __builtin_ffs(x); }
template <class T> static inline T __ldcv(const T *p) { return *(const volatile T *)p; }
using std::max;
using std::min;

// ---- launch: grid x block, CTAs one after another --------------------------------------------------------------------
template <class F> static void emu_launch(unsigned grid_x, unsigned grid_y, unsigned block_x, F &&kernel_call) {
    std::vector<std::thread> th(block_x);
    for (unsigned by = 0; by < grid_y; by++)
        for (unsigned bx = 0; bx < grid_x; bx++) {
            EmuCta cta;
            cta.nthreads = (int)block_x;
            for (unsigned w = 0; w * 32 < block_x; w++) cta.warps[w].size = (int)std::min(32u, block_x - 32 * w);
            for (unsigned t = 0; t < block_x; t++)
                th[t] = std::thread([&, t] {
                    threadIdx = {t, 0, 0}; blockIdx = {bx, by, 0};
                    blockDim = dim3(block_x, 1, 1); gridDim = dim3(grid_x, grid_y, 1);
                    emu_cta = &cta;
                    kernel_call();
                });
            for (auto &x : th) x.join();
        }
}
