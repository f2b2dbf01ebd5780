// first_touch_probe.cpp -- cost of the FIRST touch of a freshly allocated 1.9 GB std::vector-like block by 16 threads writing
// it with streaming stores (what the merged-triplet delivery does into the caller's new vector), four ways:
//   plain | madvise(MADV_HUGEPAGE) on the 2 MiB-aligned interior | MADV_POPULATE_WRITE by the threads | second touch (warm)
#include <sys/mman.h>
#include <immintrin.h>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cerrno>
#include <thread>
#include <vector>
#ifndef MADV_POPULATE_WRITE
#define MADV_POPULATE_WRITE 23
#endif
static double fill(char* p, size_t n, int nt, int mode)
{
    auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> th;
    for (int t = 0; t < nt; ++t)
        th.emplace_back([=] {
            const size_t a = n * t / nt & ~(size_t)4095, b = (t == nt - 1) ? n : (n * (t + 1) / nt & ~(size_t)4095);
            if (mode == 2 && madvise(p + a, b - a, MADV_POPULATE_WRITE) != 0 && t == 0) fprintf(stderr, "populate_write: %s\n", strerror(errno));
            const __m128i v = _mm_set1_epi32(t + 1);
            for (size_t i = a; i + 16 <= b; i += 16) _mm_stream_si128((__m128i*)(p + i), v);
            _mm_sfence();
        });
    for (auto& x : th) x.join();
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
}
int main()
{
    const size_t n = (size_t)1900 << 20;
    const int nt = (int)std::thread::hardware_concurrency();
    const char* names[3] = {"plain", "MADV_HUGEPAGE", "MADV_POPULATE_WRITE"};
    for (int mode = 0; mode < 3; ++mode)
        for (int rep = 0; rep < 2; ++rep) {
            char* p = (char*)mmap(nullptr, n + (4 << 20), PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
            char* q = (char*)(((uintptr_t)p + 4095) & ~(uintptr_t)4095) + 16; // like a malloc'ed block: 16 bytes into a page
            if (mode == 1) {
                const uintptr_t H = 2u << 20, a = ((uintptr_t)q + H - 1) & ~(H - 1), b = ((uintptr_t)q + n) & ~(H - 1);
                if (madvise((void*)a, b - a, MADV_HUGEPAGE) != 0) fprintf(stderr, "hugepage: %s\n", strerror(errno));
            }
            const double cold = fill(q, n - 4096, nt, mode), warm = fill(q, n - 4096, nt, 0);
            printf("%-20s threads %d: first touch %.1f ms, second touch %.1f ms\n", names[mode], nt, cold, warm);
            munmap(p, n + (4 << 20));
        }
    return 0;
}
