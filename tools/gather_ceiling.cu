// gather_ceiling.cu -- the L1 and L2 ceilings of the trace kernel's access pattern, measured.
//
// The trace kernel's loads are gathers of 1-2 bytes per lane at data-dependent addresses: ncu counts 11.5 sectors
// per global load request at C4 (24 live lanes reading from almost as many 32-byte sectors).  "Fraction of peak"
// for such a kernel needs the chip's peak for THAT pattern, not the copy bandwidth.  This benchmark issues
// independent (address does not depend on loaded data) 2-byte gathers in which every warp request touches S
// distinct sectors (S = 1, 4, 8, 12, 16, 32), over a working set that lives in L1 (16 KB, the same lines for
// every SM) or in L2 (32 MB, random sectors, L1 hit rate ~ 0), with many loads in flight per thread, and reports
// requests/s and sector bytes/s.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather_ceiling
// tools/gather_ceiling.cu; run on one GPU: ./gather_ceiling > profiles/r02_gather_ceiling.json
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t mix(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}

// Lane l of a request reads 2 bytes from sector group g = l * S / 32 of the request: the request's S sectors are
// chosen at random (per warp, per iteration) inside the working set; lanes of one group read different halfwords
// of the same sector.  UNROLL independent loads per thread per iteration.
template <int S, int UNROLL>
__global__ void __launch_bounds__(256) gather_kernel(const uint16_t *__restrict__ buf, uint32_t sector_mask, int iters,
                                                     uint32_t *__restrict__ sink) {
    const int lane = threadIdx.x & 31;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t g = (uint32_t)(lane * S) >> 5;             // sector group of this lane
    const uint32_t within = (uint32_t)lane & 15u;             // halfword inside the sector
    uint32_t acc = 0;
    uint32_t key = warp * 0x9E3779B9u + g * 0x85EBCA6Bu;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < UNROLL; u++) {
            const uint32_t sector = mix(key + (uint32_t)(it * UNROLL + u) * 0xC2B2AE35u) & sector_mask;
            acc += buf[(size_t)sector * 16u + within];
        }
    }
    if (acc == 0xFFFFFFFFu) sink[0] = acc;                    // keeps the loads alive
}

template <int S>
static void run_one(const char *level, const uint16_t *buf, uint32_t sector_mask, uint32_t *sink, int sms, bool last) {
    constexpr int UNROLL = 8;
    const int grid = sms * 8, threads = 256;
    int iters = 64;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float ms = 0;
    for (int attempt = 0; attempt < 8; attempt++) {           // grow until the launch takes > 2 ms
        gather_kernel<S, UNROLL><<<grid, threads>>>(buf, sector_mask, iters, sink);    // warm-up (fills L2 / L1)
        CK(cudaEventRecord(e0));
        gather_kernel<S, UNROLL><<<grid, threads>>>(buf, sector_mask, iters, sink);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms > 2.0f) break;
        iters *= 4;
    }
    float best = ms;
    for (int rep = 0; rep < 5; rep++) {
        CK(cudaEventRecord(e0));
        gather_kernel<S, UNROLL><<<grid, threads>>>(buf, sector_mask, iters, sink);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    const double requests = (double)grid * (threads / 32) * (double)iters * UNROLL;
    const double rps = requests / (best * 1e-3);
    printf("    {\"level\": \"%s\", \"sectors_per_request\": %d, \"ms\": %.4f, \"requests_per_s\": %.4e, "
           "\"sector_GBps\": %.1f, \"useful_GBps\": %.1f}%s\n",
           level, S, best, rps, rps * S * 32.0 / 1e9, rps * 64.0 / 1e9, last ? "" : ",");
    CK(cudaEventDestroy(e0)); CK(cudaEventDestroy(e1));
}

int main() {
    cudaDeviceProp p;
    CK(cudaGetDeviceProperties(&p, 0));
    const size_t l2_bytes = 32u << 20, l1_bytes = 16u << 10;
    uint16_t *buf;
    uint32_t *sink;
    CK(cudaMalloc(&buf, l2_bytes));
    CK(cudaMemset(buf, 1, l2_bytes));
    CK(cudaMalloc(&sink, 4));
    printf("{\n  \"gpu\": \"%s\", \"sms\": %d, \"sm_clock_khz\": %d,\n", p.name, p.multiProcessorCount, p.clockRate);
    printf("  \"what\": \"independent 2-byte gathers, S distinct 32-byte sectors per warp request, 8 loads in flight per thread, "
           "8 CTAs x 256 threads per SM; L1 = 16 KB working set (every SM reads the same lines), L2 = 32 MB working set "
           "(random sectors); sector_GBps = requests/s x S x 32 B\",\n  \"results\": [\n");
    const uint32_t m1 = (uint32_t)(l1_bytes / 32 - 1), m2 = (uint32_t)(l2_bytes / 32 - 1);
    const int sms = p.multiProcessorCount;
    run_one<1>("L1", buf, m1, sink, sms, false);
    run_one<4>("L1", buf, m1, sink, sms, false);
    run_one<8>("L1", buf, m1, sink, sms, false);
    run_one<12>("L1", buf, m1, sink, sms, false);
    run_one<16>("L1", buf, m1, sink, sms, false);
    run_one<32>("L1", buf, m1, sink, sms, false);
    run_one<1>("L2", buf, m2, sink, sms, false);
    run_one<4>("L2", buf, m2, sink, sms, false);
    run_one<8>("L2", buf, m2, sink, sms, false);
    run_one<12>("L2", buf, m2, sink, sms, false);
    run_one<16>("L2", buf, m2, sink, sms, false);
    run_one<32>("L2", buf, m2, sink, sms, true);
    printf("  ]\n}\n");
    return 0;
}
