// tsan_main.cpp -- TEST INFRASTRUCTURE ONLY.  The emulated kernels under ThreadSanitizer: a host-side racecheck of
// the kernels' shared-memory protocol (every CUDA thread is an OS thread, every barrier / collective a real
// synchronisation, so a missing __syncthreads / __syncwarp in radiance_kernels.cuh is a data race TSAN reports).
//   g++ -fsanitize=thread -O1 -g -std=c++17 -pthread -ffp-contract=off [-DLUMCU_...] tsan_main.cpp && ./a.out
// Exit code 0 = no report (run with TSAN_OPTIONS=exitcode=66).  Scene: a small dense random world, so that lanes
// refill, warps park and the tail warps run.
#include "kernels_on_host.cpp"

#include <cstdio>
#include <cstdlib>
#include <random>

int main(int argc, char **argv) {
    const int wx = 6, wy = 5, wz = 3, P = 12;
    std::mt19937 rng(argc > 1 ? (unsigned)std::atoi(argv[1]) : 7u);      // (other seeds: other worlds, so other refill / park / tail schedules)
    std::vector<int16_t> world(wx * wy * wz);
    for (auto &w : world) w = (rng() % 4 == 0) ? (int16_t)(1 + rng() % (P - 1)) : 0;
    for (int i = 0; i < wx * wy; i++) world[i] = 1;                       // ground layer
    std::vector<uint8_t> bricks((size_t)P * 4096, 0);
    for (size_t i = 4096; i < bricks.size(); i++) bricks[i] = (rng() % 10 < 3) ? (uint8_t)(1 + rng() % 255) : 0;
    std::vector<float> mats(256 * 6);
    for (auto &m : mats) m = (float)(rng() % 1000) / 1000.0f;
    std::vector<uint32_t> f0(world.size()), f1(world.size());
    for (auto &f : f0) f = rng();
    std::vector<int16_t> req;
    for (int z = 0; z < wz; z++) for (int y = 0; y < wy; y++) for (int x = 0; x < wx; x++)
        if (req.size() / 4 < 25) { req.push_back((int16_t)x); req.push_back((int16_t)y); req.push_back((int16_t)z); req.push_back(0); }
    const long long nreq = (long long)req.size() / 4;
    int rc = 0;
    for (int R : {64, 16}) {
        std::vector<float> table((size_t)R * 3);
        for (int i = 0; i < R; i++) { table[3 * i] = 0.6f; table[3 * i + 1] = (i & 2) ? 0.48f : -0.48f; table[3 * i + 2] = (i & 4) ? 0.64f : -0.64f; }
        const float sun[3] = {0.43685203f, 0.43685203f, -0.78633365f};
        std::vector<float> pre((size_t)nreq * 3);
        std::vector<uint32_t> rs((size_t)nreq * R);
        std::vector<uint8_t> rh((size_t)nreq * R);
        unsigned long long tot[2];
        for (int variant : {1, 0}) {
            rc |= kernels_on_host_update(wx, wy, wz, P, 1, R, 0.01f, world.data(), bricks.data(), mats.data(), f0.data(), f1.data(),
                                         req.data(), nreq, 3, 1, 0, table.data(), sun, variant, 0, 1, pre.data(), rs.data(), rh.data(), tot, 0);
            std::printf("R=%d variant=%d: S=%llu H=%llu\n", R, variant, tot[0], tot[1]);
        }
    }
    return rc;
}
