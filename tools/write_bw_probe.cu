// Write-only HBM bandwidth probe: which store flavour gets closest to the part's write ceiling?
// (diagnostic for k_hessian_expand, whose traffic is 88% writes).  nvcc -gencode arch=compute_100a,code=sm_100a
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__global__ void k_st16(int4* __restrict__ d, size_t n)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) d[i] = make_int4((int)i, 1, 2, 3);
}
__global__ void k_st16_cs(int4* __restrict__ d, size_t n)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) __stcs(d + i, make_int4((int)i, 1, 2, 3));
}
__global__ void k_st16_wt(int4* __restrict__ d, size_t n)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) __stwt(d + i, make_int4((int)i, 1, 2, 3));
}
// 32-byte store per thread (sm_100: st.global.v4.b64)
__global__ void k_st32(int4* __restrict__ d, size_t n)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (2 * i + 1 < n) {
        unsigned long long a = i, b = 1, c = 2, e = 3;
        asm volatile("st.global.v4.b64 [%0], {%1,%2,%3,%4};" ::"l"(d + 2 * i), "l"(a), "l"(b), "l"(c), "l"(e) : "memory");
    }
}
// 4 x 16-byte stores per thread, each warp-coalesced (thread t writes words t, t+32*..)
template <int U, bool CS>
__global__ void k_st16_unroll(int4* __restrict__ d, size_t n)
{
    size_t base = ((size_t)blockIdx.x * blockDim.x) * U + threadIdx.x;
#pragma unroll
    for (int u = 0; u < U; ++u) {
        size_t i = base + (size_t)u * blockDim.x;
        if (i < n) { if (CS) __stcs(d + i, make_int4((int)i, 1, 2, 3)); else d[i] = make_int4((int)i, 1, 2, 3); }
    }
}
// persistent grid-stride
__global__ void k_st16_persist(int4* __restrict__ d, size_t n)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) d[i] = make_int4((int)i, 1, 2, 3);
}
// TMA 1-D bulk store: fill a shared tile, one thread issues cp.async.bulk.global.shared::cta
template <int TILE_BYTES>
__global__ void k_bulk(int4* __restrict__ d, size_t n)
{
    extern __shared__ __align__(128) unsigned char sm[];
    int4* s = reinterpret_cast<int4*>(sm);
    constexpr int W = TILE_BYTES / 16;
    const size_t tiles = n / W;
    for (int t = threadIdx.x; t < W; t += blockDim.x) s[t] = make_int4(t, 1, 2, 3);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
        for (size_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
            unsigned sa = (unsigned)__cvta_generic_to_shared(s);
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(d + tile * W), "r"(sa), "r"(TILE_BYTES) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 4;" ::: "memory");
        }
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
}

template <class F>
float best_ms(F f, int reps = 6)
{
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b); if (r && ms < best) best = ms;
    }
    return best;
}

int main()
{
    const size_t bytes = (size_t)12 << 30, n = bytes / 16;
    int4* d; CK(cudaMalloc(&d, bytes));
    auto rep = [&](const char* name, float ms) { printf("%-34s %8.3f ms  %8.1f GB/s\n", name, ms, bytes / ms / 1e6); };
    rep("cudaMemset", best_ms([&] { cudaMemsetAsync(d, 0, bytes); }));
    rep("st.16B one/thread", best_ms([&] { k_st16<<<(unsigned)((n + 255) / 256), 256>>>(d, n); }));
    rep("st.cs 16B one/thread", best_ms([&] { k_st16_cs<<<(unsigned)((n + 255) / 256), 256>>>(d, n); }));
    rep("st.wt 16B one/thread", best_ms([&] { k_st16_wt<<<(unsigned)((n + 255) / 256), 256>>>(d, n); }));
    rep("st.32B one/thread", best_ms([&] { k_st32<<<(unsigned)((n / 2 + 255) / 256), 256>>>(d, n); }));
    rep("st.16B x4/thread", best_ms([&] { k_st16_unroll<4, false><<<(unsigned)((n + 1023) / 1024), 256>>>(d, n); }));
    rep("st.cs 16B x4/thread", best_ms([&] { k_st16_unroll<4, true><<<(unsigned)((n + 1023) / 1024), 256>>>(d, n); }));
    rep("st.16B x9/thread", best_ms([&] { k_st16_unroll<9, false><<<(unsigned)((n + 2303) / 2304), 256>>>(d, n); }));
    rep("st.16B persistent 148x8x256", best_ms([&] { k_st16_persist<<<148 * 8, 256>>>(d, n); }));
    rep("st.16B persistent 148x4x512", best_ms([&] { k_st16_persist<<<148 * 4, 512>>>(d, n); }));
    cudaFuncSetAttribute(k_bulk<16384>, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384);
    cudaFuncSetAttribute(k_bulk<32768>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);
    rep("TMA bulk store 16 KB tiles 148x4", best_ms([&] { k_bulk<16384><<<148 * 4, 128, 16384>>>(d, n); }));
    rep("TMA bulk store 32 KB tiles 148x4", best_ms([&] { k_bulk<32768><<<148 * 4, 128, 32768>>>(d, n); }));
    rep("TMA bulk store 16 KB tiles 148x8", best_ms([&] { k_bulk<16384><<<148 * 8, 128, 16384>>>(d, n); }));
    // copy for reference (read + write bytes)
    {
        float ms = best_ms([&] { cudaMemcpyAsync(d, d + n / 2, bytes / 2, cudaMemcpyDeviceToDevice); });
        printf("%-34s %8.3f ms  %8.1f GB/s (read+write)\n", "cudaMemcpy D2D 6+6 GiB", ms, bytes / ms / 1e6);
    }
    CK(cudaDeviceSynchronize());
    CK(cudaGetLastError());
    return 0;
}
