// Host-side fp64 -> int32 table conversion throughput (development aid): g++ -O3 -std=c++17 -pthread tools/host_convert_probe.cpp -o /tmp/conv && /tmp/conv <threads>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cmath>
#include <thread>
#include <vector>
#include <chrono>
#include <atomic>
// returns 1 if every element converted exactly
__attribute__((target("avx2"))) static int conv(const double* __restrict ts, const double* __restrict te, int32_t* __restrict ots, int32_t* __restrict ote, size_t n, double jit) {
    int ok = 1;
    for (size_t i = 0; i < n; ++i) {
        const double a = ts[i], b = te[i] - jit;
        const int32_t ia = (int32_t)a, ib = (int32_t)b;
        ok &= ((double)ia == a) & ((double)ib == b);
        ots[i] = ia; ote[i] = ib;
    }
    return ok;
}
int main(int argc, char** argv) {
    const int T = argc > 1 ? atoi(argv[1]) : 8;
    const size_t n = (size_t)1 << 26;    // 64 Mi lineages: 1 GiB of fp64 in
    double* ts = (double*)aligned_alloc(64, n * 8); double* te = (double*)aligned_alloc(64, n * 8);
    int32_t* ots = (int32_t*)aligned_alloc(64, n * 4); int32_t* ote = (int32_t*)aligned_alloc(64, n * 4);
    for (size_t i = 0; i < n; ++i) { ts[i] = 1800 + (i * 7919) % 200; te[i] = ts[i] + (i % 50) + 0.5; ots[i] = ote[i] = 0; }
    for (int rep = 0; rep < 3; ++rep) {
        std::atomic<int> ok{1};
        auto t0 = std::chrono::steady_clock::now();
        std::vector<std::thread> th;
        for (int t = 0; t < T; ++t) th.emplace_back([&, t] { size_t a = n * t / T, b = n * (t + 1) / T; if (!conv(ts + a, te + a, ots + a, ote + a, b - a, 0.5)) ok = 0; });
        for (auto& x : th) x.join();
        double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        printf("threads %d: %.1f ms, %.1f GB/s of fp64 in, ok=%d\n", T, s * 1e3, n * 16 / s / 1e9, (int)ok);
    }
}
