// util.cuh -- error handling, grow-only device buffers, launch accounting.
#pragma once
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>
#include <map>
#include <stdexcept>

namespace cipc {

struct CudaError : std::runtime_error {
    explicit CudaError(const std::string& s) : std::runtime_error(s) {}
};

#define CIPC_CUDA(expr)                                                                           \
    do {                                                                                          \
        cudaError_t _e = (expr);                                                                  \
        if (_e != cudaSuccess) {                                                                  \
            char _b[512];                                                                         \
            snprintf(_b, sizeof(_b), "%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            throw ::cipc::CudaError(_b);                                                          \
        }                                                                                         \
    } while (0)

// Grow-only typed device buffer (capacity in elements).  Contents are NOT preserved on growth
// unless keep=true.
template <class T>
struct DevBuf {
    T* p = nullptr;
    size_t cap = 0;
    ~DevBuf() { if (p) cudaFree(p); }
    DevBuf() {}
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    void reserve(size_t n, cudaStream_t s = 0, bool keep = false)
    {
        if (n <= cap) return;
        size_t ncap = n + n / 4 + 64;
        T* q = nullptr;
        CIPC_CUDA(cudaMalloc(&q, ncap * sizeof(T)));
        if (keep && p && cap) CIPC_CUDA(cudaMemcpyAsync(q, p, cap * sizeof(T), cudaMemcpyDeviceToDevice, s));
        if (p) { CIPC_CUDA(cudaStreamSynchronize(s)); cudaFree(p); }
        p = q;
        cap = ncap;
    }
    size_t bytes() const { return cap * sizeof(T); }
};

// Pinned host staging buffer (grow-only)
struct PinnedBuf {
    void* p = nullptr;
    size_t cap = 0;
    ~PinnedBuf() { if (p) cudaFreeHost(p); }
    void* reserve(size_t bytes)
    {
        if (bytes > cap) {
            if (p) cudaFreeHost(p);
            p = nullptr;
            cap = bytes + bytes / 4 + 4096;
            CIPC_CUDA(cudaMallocHost(&p, cap));
        }
        return p;
    }
};

static inline unsigned div_up(size_t a, size_t b) { return (unsigned)((a + b - 1) / b); }

} // namespace cipc
