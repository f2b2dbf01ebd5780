// hostpool.h -- host-side helpers of the C ABI's delivery paths: a persistent worker pool (the caller's containers are
// filled by all host cores: triplet expansion, copies into pageable std::vector storage, content hashes) and staged
// transfers between device memory and PAGEABLE host memory.
//
// Why: the reference's containers (std::vector, Cabana AoSoA) are pageable.  cudaMemcpy to / from pageable memory is
// staged by the driver on one thread at ~10 GB/s; here the staging buffer is ours (pinned, two 16 MiB halves) and the
// pageable side is moved by the pool at memory bandwidth while the other half is on the wire.
#pragma once
#include <cuda_runtime.h>
#include <sys/mman.h>
#include <atomic>
#include <condition_variable>
#include <cstring>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

namespace cipc {

class HostPool {
public:
    static HostPool& get()
    {
        static HostPool* p = new HostPool(); // intentionally leaked: worker threads may outlive static destructors
        return *p;
    }
    int size() const { return n_; }
    // runs f(tid) once on every thread of the pool (tid 0 = the caller); returns when all are done
    void run(const std::function<void(int)>& f)
    {
        if (n_ == 1) { f(0); return; }
        std::unique_lock<std::mutex> serial(serial_); // one parallel region at a time
        {
            std::lock_guard<std::mutex> lk(m_);
            job_ = &f;
            pending_ = n_ - 1;
            ++gen_;
        }
        cv_.notify_all();
        f(0);
        std::unique_lock<std::mutex> lk(m_);
        done_.wait(lk, [&] { return pending_ == 0; });
        job_ = nullptr;
    }
    // dynamic scheduling of nItems work items over the pool
    template <class F>
    void for_each(size_t nItems, F f)
    {
        if (nItems == 0) return;
        std::atomic<size_t> next(0);
        run([&](int) { for (size_t k; (k = next.fetch_add(1)) < nItems;) f(k); });
    }
    void copy(void* dst, const void* src, size_t bytes)
    {
        const size_t PIECE = (size_t)256 << 10;
        if (bytes <= 2 * PIECE) { memcpy(dst, src, bytes); return; }
        for_each((bytes + PIECE - 1) / PIECE, [&](size_t k) {
            const size_t o = k * PIECE;
            memcpy((char*)dst + o, (const char*)src + o, std::min(PIECE, bytes - o));
        });
    }

private:
    HostPool()
    {
        int nt = (int)std::thread::hardware_concurrency();
        if (const char* e = getenv("CIPC_HOST_THREADS")) nt = atoi(e);
        n_ = std::max(1, std::min(nt, 256));
        for (int t = 1; t < n_; ++t) std::thread([this, t] { worker(t); }).detach();
    }
    void worker(int tid)
    {
        unsigned seen = 0;
        for (;;) {
            const std::function<void(int)>* job;
            {
                std::unique_lock<std::mutex> lk(m_);
                cv_.wait(lk, [&] { return gen_ != seen; });
                seen = gen_;
                job = job_;
            }
            (*job)(tid);
            {
                std::lock_guard<std::mutex> lk(m_);
                if (--pending_ == 0) done_.notify_one();
            }
        }
    }
    int n_ = 1;
    std::mutex m_, serial_;
    std::condition_variable cv_, done_;
    const std::function<void(int)>* job_ = nullptr;
    unsigned gen_ = 0;
    int pending_ = 0;
};

inline bool host_ptr_is_pageable(const void* p)
{
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return true; }
    return a.type == cudaMemoryTypeUnregistered;
}
// asks for transparent huge pages on the interior of a large destination range before it is first touched (the box runs
// THP in `madvise` mode): 512x fewer page faults while the delivery threads fill a freshly allocated std::vector
inline void advise_huge(void* p, size_t bytes)
{
#ifdef MADV_HUGEPAGE
    const uintptr_t H = (uintptr_t)2 << 20, a = ((uintptr_t)p + H - 1) & ~(H - 1), b = ((uintptr_t)p + bytes) & ~(H - 1);
    if (b > a + H) madvise((void*)a, b - a, MADV_HUGEPAGE);
#else
    (void)p; (void)bytes;
#endif
}

// First touch of a freshly allocated destination in the background (one helper thread drives the pool): the caller reserves
// its output vector from the previous iteration's size BEFORE the device work starts, the page faults (~30 ms for 2 GB on 16
// cores) then overlap the kernels instead of the delivery.  wait() before the destination is written.
class Prefault {
public:
    static Prefault& get() { static Prefault* p = new Prefault(); return *p; }
    void start(void* p, size_t bytes)
    {
        wait();
        if (!p || bytes < ((size_t)8 << 20)) return;
        active_ = true;
        th_ = std::thread([p, bytes] {
            advise_huge(p, bytes);
            const size_t CH = (size_t)4 << 20, n = (bytes + CH - 1) / CH;
            HostPool::get().for_each(n, [&](size_t k) {
                volatile char* q = (volatile char*)p + k * CH;
                const size_t len = std::min(CH, bytes - k * CH);
                for (size_t o = 0; o < len; o += 4096) q[o] = 0;
            });
        });
    }
    void wait() { if (active_) { th_.join(); active_ = false; } }
private:
    std::thread th_;
    bool active_ = false;
};

// two pinned 16 MiB halves per context
struct StageRing {
    static constexpr size_t HALF = (size_t)16 << 20;
    char* p = nullptr;
    cudaEvent_t ev[2] = {nullptr, nullptr};
    ~StageRing() { if (p) cudaFreeHost(p); for (auto e : ev) if (e) cudaEventDestroy(e); }
    void init()
    {
        if (p) return;
        if (cudaMallocHost(&p, 2 * HALF) != cudaSuccess) throw std::runtime_error("cudaMallocHost failed for the staging ring");
        for (auto& e : ev) cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
    }
};
// device -> host; pageable destinations go through the ring with the pool copying one half while the other is in flight.
// Returns after the data has landed in dst.
inline void staged_d2h(StageRing& r, cudaStream_t st, void* dst, const void* src, size_t bytes)
{
    if (!bytes) return;
    if (bytes < ((size_t)1 << 20) || !host_ptr_is_pageable(dst)) {
        if (cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, st) != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess)
            throw std::runtime_error("device-to-host copy failed");
        return;
    }
    r.init();
    HostPool& pool = HostPool::get();
    const size_t H = StageRing::HALF, n = (bytes + H - 1) / H;
    bool ok = true;
    for (size_t k = 0; k <= n; ++k) {
        if (k < n) {
            const size_t o = k * H, len = std::min(H, bytes - o);
            ok = ok && cudaMemcpyAsync(r.p + (k & 1) * H, (const char*)src + o, len, cudaMemcpyDeviceToHost, st) == cudaSuccess;
            ok = ok && cudaEventRecord(r.ev[k & 1], st) == cudaSuccess;
        }
        if (k > 0) {
            const size_t o = (k - 1) * H, len = std::min(H, bytes - o);
            ok = ok && cudaEventSynchronize(r.ev[(k - 1) & 1]) == cudaSuccess;
            if (ok) pool.copy((char*)dst + o, r.p + ((k - 1) & 1) * H, len);
        }
    }
    if (!ok) throw std::runtime_error("staged device-to-host copy failed");
}
// host -> device; returns when the source may be reused (the last half may still be on the wire: stream-ordered)
inline void staged_h2d(StageRing& r, cudaStream_t st, void* dst, const void* src, size_t bytes)
{
    if (!bytes) return;
    if (bytes < ((size_t)1 << 20) || !host_ptr_is_pageable(src)) {
        if (cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, st) != cudaSuccess) throw std::runtime_error("host-to-device copy failed");
        return;
    }
    r.init();
    HostPool& pool = HostPool::get();
    const size_t H = StageRing::HALF, n = (bytes + H - 1) / H;
    bool ok = true;
    for (size_t k = 0; k < n; ++k) {
        const size_t o = k * H, len = std::min(H, bytes - o);
        if (k >= 2) ok = ok && cudaEventSynchronize(r.ev[k & 1]) == cudaSuccess; // the half is free again
        pool.copy(r.p + (k & 1) * H, (const char*)src + o, len);
        ok = ok && cudaMemcpyAsync((char*)dst + o, r.p + (k & 1) * H, len, cudaMemcpyHostToDevice, st) == cudaSuccess;
        ok = ok && cudaEventRecord(r.ev[k & 1], st) == cudaSuccess;
    }
    // the ring must not be refilled by a later call before these copies have read it
    ok = ok && cudaEventSynchronize(r.ev[(n - 1) & 1]) == cudaSuccess && (n < 2 || cudaEventSynchronize(r.ev[(n - 2) & 1]) == cudaSuccess);
    if (!ok) throw std::runtime_error("staged host-to-device copy failed");
}

} // namespace cipc
