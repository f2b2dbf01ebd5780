// fp64_fma_probe.cu -- measured FP64 (non-tensor) FMA throughput of this GPU: the denominator of `roofline_fp64` in bench.py.
// Every thread runs NCHAIN independent dependent-FMA chains (enough ILP to cover the DFMA latency at any residency); the grid
// fills the device several times over.  flops = 2 x FMAs.  Usage: fp64_fma_probe [device]  -> one JSON line.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <vector>

template <int NCHAIN>
__global__ void __launch_bounds__(256) k_dfma(double* out, int iters, double a, double b)
{
    double x[NCHAIN];
#pragma unroll
    for (int k = 0; k < NCHAIN; ++k) x[k] = 1.0 + 1e-3 * (threadIdx.x + k);
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < NCHAIN; ++k) x[k] = fma(x[k], a, b);
    }
    double s = 0;
#pragma unroll
    for (int k = 0; k < NCHAIN; ++k) s += x[k];
    if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s; // never true: keeps the chains alive
}

template <int NCHAIN>
double run(int blocks, int iters, double* d)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k_dfma<NCHAIN><<<blocks, 256>>>(d, iters / 8, 0.999999, 1e-7); // warm-up
    std::vector<float> ms;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0);
        k_dfma<NCHAIN><<<blocks, 256>>>(d, iters, 0.999999, 1e-7);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float t; cudaEventElapsedTime(&t, e0, e1); ms.push_back(t);
    }
    const double best = *std::min_element(ms.begin(), ms.end());
    return 2.0 * NCHAIN * (double)iters * 256.0 * blocks / (best * 1e-3) / 1e12;
}

int main(int argc, char** argv)
{
    const int dev = argc > 1 ? atoi(argv[1]) : 0;
    if (cudaSetDevice(dev) != cudaSuccess) { printf("{\"error\": \"no CUDA device\"}\n"); return 1; }
    cudaDeviceProp p; cudaGetDeviceProperties(&p, dev);
    int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, dev);
    double* d; cudaMalloc(&d, 1 << 26);
    const int blocks = p.multiProcessorCount * 8 * 4, iters = 1 << 14;
    const double t4 = run<4>(blocks, iters, d), t8 = run<8>(blocks, iters, d), t16 = run<16>(blocks, iters, d);
    const double best = std::max(t4, std::max(t8, t16));
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"sm_clock_mhz\": %.0f, \"fp64_tflops\": %.3f, \"by_chains\": {\"4\": %.3f, \"8\": %.3f, \"16\": %.3f}, "
           "\"fma_per_clk_per_sm\": %.1f, \"how\": \"dependent DFMA chains, 256 threads x %d blocks, best of 5, CUDA events\"}\n",
        p.name, p.multiProcessorCount, clk / 1e3, best, t4, t8, t16, best * 1e12 / 2.0 / (p.multiProcessorCount * (clk * 1e3)), blocks);
    return 0;
}
