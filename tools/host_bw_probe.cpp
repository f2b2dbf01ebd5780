// Measures multi-threaded host streaming-store bandwidth (what a host-side triplet expansion could sustain).
#include <immintrin.h>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <vector>
int main(int argc, char** argv)
{
    const int nt = argc > 1 ? atoi(argv[1]) : (int)std::thread::hardware_concurrency();
    const size_t bytes = (size_t)(argc > 2 ? atof(argv[2]) : 4.0) * (1ull << 30);
    double* buf = (double*)aligned_alloc(4096, bytes);
    for (size_t i = 0; i < bytes / 8; i += 512) buf[i] = 0; // touch pages
    for (int mode = 0; mode < 2; ++mode)
        for (int rep = 0; rep < 3; ++rep) {
            auto t0 = std::chrono::steady_clock::now();
            std::vector<std::thread> th;
            for (int t = 0; t < nt; ++t)
                th.emplace_back([=]() {
                    const size_t n = bytes / 8 / nt;
                    double* p = buf + (size_t)t * n;
                    const __m256d v = _mm256_set1_pd(1.5 + t);
                    if (mode == 0) for (size_t i = 0; i + 4 <= n; i += 4) _mm256_stream_pd(p + i, v);
                    else for (size_t i = 0; i + 4 <= n; i += 4) _mm256_store_pd(p + i, v);
                });
            for (auto& x : th) x.join();
            const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            printf("%s stores, %d threads: %.1f GB/s\n", mode == 0 ? "streaming" : "regular", nt, bytes / s / 1e9);
        }
    return 0;
}
