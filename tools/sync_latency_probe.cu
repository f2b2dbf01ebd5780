// sync_latency_probe.cu -- cost of one host round trip for a device-side counter, three ways:
//   (a) cudaMemcpyAsync(4 B, D2H, pageable) + cudaStreamSynchronize      (what the stage bodies did in round 1)
//   (b) the same into pinned memory
//   (c) a 1-warp kernel that publishes the value + a sequence number into MAPPED pinned memory, host spins on the sequence
// each after a short producer kernel, 2000 repetitions, median.
#include <cuda_runtime.h>
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <vector>
#include <immintrin.h>

__global__ void k_produce(unsigned* c) { if (threadIdx.x == 0) atomicAdd(c, 1u); }
__global__ void k_publish(const unsigned* src, volatile unsigned* dst, unsigned seq)
{
    if (threadIdx.x == 0) { dst[0] = src[0]; __threadfence_system(); dst[1] = seq; }
}
static double med(std::vector<double>& v) { std::sort(v.begin(), v.end()); return v[v.size() / 2]; }
int main()
{
    cudaStream_t st; cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
    unsigned* d; cudaMalloc(&d, 64); cudaMemset(d, 0, 64);
    unsigned pageable = 0, *pinned, *mapped, *mappedDev;
    cudaMallocHost(&pinned, 64);
    cudaHostAlloc(&mapped, 64, cudaHostAllocMapped); cudaHostGetDevicePointer(&mappedDev, mapped, 0);
    mapped[0] = mapped[1] = 0;
    auto now = [] { return std::chrono::steady_clock::now(); };
    std::vector<double> a, b, c;
    for (int i = 0; i < 2200; ++i) {
        auto t0 = now();
        k_produce<<<1, 32, 0, st>>>(d);
        cudaMemcpyAsync(&pageable, d, 4, cudaMemcpyDeviceToHost, st); cudaStreamSynchronize(st);
        auto t1 = now();
        k_produce<<<1, 32, 0, st>>>(d);
        cudaMemcpyAsync(pinned, d, 4, cudaMemcpyDeviceToHost, st); cudaStreamSynchronize(st);
        auto t2 = now();
        k_produce<<<1, 32, 0, st>>>(d);
        k_publish<<<1, 32, 0, st>>>(d, mappedDev, (unsigned)i + 1);
        while (((volatile unsigned*)mapped)[1] != (unsigned)i + 1) _mm_pause();
        auto t3 = now();
        if (i >= 200) {
            a.push_back(std::chrono::duration<double, std::micro>(t1 - t0).count());
            b.push_back(std::chrono::duration<double, std::micro>(t2 - t1).count());
            c.push_back(std::chrono::duration<double, std::micro>(t3 - t2).count());
        }
    }
    printf("{\"round_trip_us\": {\"memcpy_pageable_sync\": %.1f, \"memcpy_pinned_sync\": %.1f, \"mapped_publish_spin\": %.1f}}\n", med(a), med(b), med(c));
    return 0;
}
