// prims.cuh -- device-wide primitives written for this path: exclusive scan (warp-shuffle based),
// stable LSD radix sort of (u32 key, u32 value) pairs, deterministic fixed-tree reductions.
// Grids are sized from the tile count; all kernels are bandwidth-bound streaming kernels.
#pragma once
#include "util.cuh"
#include <cstdlib>

namespace cipc {

extern long g_launches; // kernels launched by this library (reported as gpu_launches)

#define CIPC_LAUNCH(kernel, grid, block, smem, stream, ...)                                         \
    do {                                                                                          \
        kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);                               \
        ++::cipc::g_launches;                                                                     \
        CIPC_CUDA(cudaGetLastError());                                                            \
    } while (0)

typedef unsigned int u32;
typedef unsigned long long u64;

// ------------------------------------------------------------------ warp / block scan
__device__ __forceinline__ u32 warp_incl_scan(u32 v)
{
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const u32 t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
    }
    return v;
}
// exclusive scan across a block of BT threads; returns exclusive prefix, total in `total`
template <int BT>
__device__ __forceinline__ u32 block_excl_scan(u32 v, u32& total)
{
    __shared__ u32 wsum[BT / 32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const u32 inc = warp_incl_scan(v);
    if (lane == 31) wsum[w] = inc;
    __syncthreads();
    if (w == 0) {
        u32 s = (lane < BT / 32) ? wsum[lane] : 0;
        s = warp_incl_scan(s);
        if (lane < BT / 32) wsum[lane] = s;
    }
    __syncthreads();
    const u32 base = (w == 0) ? 0 : wsum[w - 1];
    total = wsum[BT / 32 - 1];
    __syncthreads();
    return base + inc - v;
}

// ------------------------------------------------------------------ device exclusive scan (u32)
constexpr int SCAN_BT = 256, SCAN_IT = 8, SCAN_TILE = SCAN_BT * SCAN_IT;

__global__ void k_scan_tile_sums(const u32* __restrict__ in, u32* __restrict__ tileSum, size_t n)
{
    const size_t base = (size_t)blockIdx.x * SCAN_TILE;
    u32 s = 0;
#pragma unroll
    for (int i = 0; i < SCAN_IT; ++i) {
        const size_t idx = base + (size_t)i * SCAN_BT + threadIdx.x;
        if (idx < n) s += in[idx];
    }
    u32 total;
    block_excl_scan<SCAN_BT>(s, total);
    if (threadIdx.x == 0) tileSum[blockIdx.x] = total;
}
// single block: exclusive scan of tileSum[nt] in place; total -> *total
__global__ void k_scan_tile_offsets(u32* tileSum, int nt, u32* total)
{
    __shared__ u32 carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < nt; base += 1024) {
        const int i = base + threadIdx.x;
        const u32 v = (i < nt) ? tileSum[i] : 0;
        u32 tot;
        const u32 ex = block_excl_scan<1024>(v, tot);
        const u32 c = carry;
        if (i < nt) tileSum[i] = ex + c;
        __syncthreads();
        if (threadIdx.x == 0) carry = c + tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry;
}
__global__ void k_scan_apply(const u32* __restrict__ in, u32* __restrict__ out, const u32* __restrict__ tileOff, size_t n)
{
    // thread t owns the SCAN_IT consecutive items [base + t*IT, base + (t+1)*IT)
    const size_t base = (size_t)blockIdx.x * SCAN_TILE + (size_t)threadIdx.x * SCAN_IT;
    u32 v[SCAN_IT], s = 0;
#pragma unroll
    for (int i = 0; i < SCAN_IT; ++i) {
        v[i] = (base + i < n) ? in[base + i] : 0;
        s += v[i];
    }
    u32 total;
    u32 ex = block_excl_scan<SCAN_BT>(s, total) + tileOff[blockIdx.x];
#pragma unroll
    for (int i = 0; i < SCAN_IT; ++i) {
        if (base + i < n) out[base + i] = ex;
        ex += v[i];
    }
}

// Single-pass scan (decoupled look-back): tiles take a ticket in launch order, publish their aggregate, and resolve
// their exclusive prefix from the descriptors of the preceding tiles (flag in the high word: 1 = aggregate only,
// 2 = inclusive prefix).  One kernel and one read of the input instead of three kernels and two reads; the hash build
// runs a dozen scans per contact stage, most of them launch-latency bound.
__global__ void __launch_bounds__(SCAN_BT) k_scan_onepass(const u32* __restrict__ in, u32* __restrict__ out, size_t n, u64* desc, u32* ticket,
    u32* total, u32 nt)
{
    __shared__ u32 sTile, sPrefix;
    if (threadIdx.x == 0) sTile = atomicAdd(ticket, 1u);
    __syncthreads();
    const u32 tile = sTile;
    const size_t base = (size_t)tile * SCAN_TILE + (size_t)threadIdx.x * SCAN_IT;
    u32 v[SCAN_IT], s = 0;
#pragma unroll
    for (int i = 0; i < SCAN_IT; ++i) {
        v[i] = (base + i < n) ? in[base + i] : 0;
        s += v[i];
    }
    u32 tileTotal;
    u32 ex = block_excl_scan<SCAN_BT>(s, tileTotal);
    if (threadIdx.x < 32) { // warp 0 publishes the aggregate and looks back 32 descriptors at a time
        volatile u64* d = desc;
        const int lane = threadIdx.x;
        u32 prefix = 0;
        if (lane == 0) d[tile] = ((u64)(tile == 0 ? 2 : 1) << 32) | tileTotal;
        if (tile != 0) {
            for (int j0 = (int)tile - 1;; j0 -= 32) {
                const int j = j0 - lane;
                u64 w = (u64)2 << 32; // before the first tile: inclusive prefix 0
                if (j >= 0) do { w = d[j]; } while ((w >> 32) == 0);
                const u32 isP = __ballot_sync(0xffffffffu, (w >> 32) == 2);
                const int firstP = isP ? __ffs(isP) - 1 : 31; // nearest tile that already knows its inclusive prefix
                u32 c = (lane <= firstP) ? (u32)w : 0u;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
                prefix += c;
                if (isP) break;
            }
            if (lane == 0) d[tile] = ((u64)2 << 32) | (u32)(prefix + tileTotal);
        }
        if (lane == 0) {
            sPrefix = prefix;
            if (tile == nt - 1) *total = prefix + tileTotal;
        }
    }
    __syncthreads();
    ex += sPrefix;
#pragma unroll
    for (int i = 0; i < SCAN_IT; ++i) {
        if (base + i < n) out[base + i] = ex;
        ex += v[i];
    }
}

struct ScanWork {
    DevBuf<u32> tileSum;
    DevBuf<u32> total;
    DevBuf<u64> desc; // [0] = ticket counter (low word), [1..] = tile descriptors
};
// out[i] = sum_{j<i} in[j]; *total (device) = sum of all.  in may alias out.
static inline void device_excl_scan(const u32* in, u32* out, size_t n, ScanWork& wk, cudaStream_t s)
{
    wk.total.reserve(1, s);
    if (n == 0) { CIPC_CUDA(cudaMemsetAsync(wk.total.p, 0, sizeof(u32), s)); return; }
    const int nt = div_up(n, SCAN_TILE);
    // default: the one-pass look-back kernel (one launch + one memset instead of three launches: the small scans of the hash
    // build are launch-latency bound; equal speed on the large ones); CIPC_SCAN_3PASS=1 selects the three streaming kernels
    static const bool threePass = getenv("CIPC_SCAN_3PASS") != nullptr;
    if (!threePass) {
        wk.desc.reserve((size_t)nt + 1, s);
        CIPC_CUDA(cudaMemsetAsync(wk.desc.p, 0, ((size_t)nt + 1) * sizeof(u64), s));
        CIPC_LAUNCH(k_scan_onepass, nt, SCAN_BT, 0, s, in, out, n, wk.desc.p + 1, (u32*)wk.desc.p, wk.total.p, (u32)nt);
        return;
    }
    wk.tileSum.reserve(nt, s);
    CIPC_LAUNCH(k_scan_tile_sums, nt, SCAN_BT, 0, s, in, wk.tileSum.p, n);
    CIPC_LAUNCH(k_scan_tile_offsets, 1, 1024, 0, s, wk.tileSum.p, nt, wk.total.p);
    CIPC_LAUNCH(k_scan_apply, nt, SCAN_BT, 0, s, in, out, wk.tileSum.p, n);
}

// ------------------------------------------------------------------ LSD radix sort (u32 key, u32 value)
constexpr int RS_BT = 256, RS_IT = 8, RS_TILE = RS_BT * RS_IT, RS_WARPS = RS_BT / 32;

__global__ void k_radix_hist(const u32* __restrict__ keys, u32* __restrict__ hist, size_t n, int shift, int nb)
{
    __shared__ u32 h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    const size_t base = (size_t)blockIdx.x * RS_TILE;
#pragma unroll
    for (int i = 0; i < RS_IT; ++i) {
        const size_t idx = base + (size_t)i * RS_BT + threadIdx.x;
        if (idx < n) atomicAdd(&h[(keys[idx] >> shift) & 255u], 1u);
    }
    __syncthreads();
    hist[(size_t)threadIdx.x * nb + blockIdx.x] = h[threadIdx.x];
}
// stable scatter: order inside a tile is (item i, warp w, lane)
__global__ void k_radix_scatter(const u32* __restrict__ kin, const u32* __restrict__ vin, u32* __restrict__ kout,
    u32* __restrict__ vout, const u32* __restrict__ histScan, size_t n, int shift, int nb)
{
    __shared__ unsigned short cnt[RS_IT * RS_WARPS][256];
    for (int i = threadIdx.x; i < RS_IT * RS_WARPS * 256 / 2; i += RS_BT) ((u32*)cnt)[i] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const u32 lt = (1u << lane) - 1u;
    const size_t base = (size_t)blockIdx.x * RS_TILE;
    u32 key[RS_IT], val[RS_IT], rank[RS_IT];
#pragma unroll
    for (int i = 0; i < RS_IT; ++i) {
        const size_t idx = base + (size_t)i * RS_BT + threadIdx.x;
        const bool valid = idx < n;
        key[i] = valid ? kin[idx] : 0u;
        val[i] = valid ? vin[idx] : 0u;
        const u32 digit = valid ? ((key[i] >> shift) & 255u) : 256u;
        const u32 m = __match_any_sync(0xffffffffu, digit);
        rank[i] = __popc(m & lt);
        if (valid && rank[i] == 0) cnt[i * RS_WARPS + w][digit] = (unsigned short)__popc(m);
    }
    __syncthreads();
    {
        const int d = threadIdx.x; // one digit per thread
        u32 run = 0;
#pragma unroll 8
        for (int g = 0; g < RS_IT * RS_WARPS; ++g) {
            const u32 t = cnt[g][d];
            cnt[g][d] = (unsigned short)run;
            run += t;
        }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < RS_IT; ++i) {
        const size_t idx = base + (size_t)i * RS_BT + threadIdx.x;
        if (idx < n) {
            const u32 digit = (key[i] >> shift) & 255u;
            const size_t pos = (size_t)histScan[(size_t)digit * nb + blockIdx.x] + cnt[i * RS_WARPS + w][digit] + rank[i];
            kout[pos] = key[i];
            vout[pos] = val[i];
        }
    }
}

struct SortWork {
    DevBuf<u32> k2, v2, hist;
    ScanWork scan;
};
// Sorts (keys, vals) by the low `bits` bits of the key, ascending, stable.  Result is left in
// (keys, vals): after an odd number of passes it is copied back (two D2D copies cost less than a fourth pass).
static inline void device_radix_sort(u32* keys, u32* vals, size_t n, int bits, SortWork& wk, cudaStream_t s)
{
    if (n <= 1) return;
    int passes = (bits + 7) / 8;
    if (passes < 1) passes = 1;
    const int nb = div_up(n, RS_TILE);
    wk.k2.reserve(n, s); wk.v2.reserve(n, s); wk.hist.reserve((size_t)256 * nb, s);
    u32 *ki = keys, *vi = vals, *ko = wk.k2.p, *vo = wk.v2.p;
    for (int p = 0; p < passes; ++p) {
        const int shift = 8 * p;
        CIPC_LAUNCH(k_radix_hist, nb, RS_BT, 0, s, ki, wk.hist.p, n, shift, nb);
        device_excl_scan(wk.hist.p, wk.hist.p, (size_t)256 * nb, wk.scan, s);
        CIPC_LAUNCH(k_radix_scatter, nb, RS_BT, 0, s, ki, vi, ko, vo, wk.hist.p, n, shift, nb);
        u32* t = ki; ki = ko; ko = t;
        t = vi; vi = vo; vo = t;
    }
    if (ki != keys) {
        CIPC_CUDA(cudaMemcpyAsync(keys, ki, n * sizeof(u32), cudaMemcpyDeviceToDevice, s));
        CIPC_CUDA(cudaMemcpyAsync(vals, vi, n * sizeof(u32), cudaMemcpyDeviceToDevice, s));
    }
}

// ------------------------------------------------------------------ deterministic reductions
// Fixed launch shape (RED_GRID x RED_BT) and fixed tree => bitwise reproducible sums.
constexpr int RED_GRID = 592, RED_BT = 256; // 4 CTAs per SM on 148 SMs

template <int BT>
__device__ __forceinline__ double block_sum(double v)
{
    __shared__ double ws[BT / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = v;
    __syncthreads();
    double r = 0;
    if (threadIdx.x < 32) {
        r = (threadIdx.x < BT / 32) ? ws[threadIdx.x] : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) r += __shfl_down_sync(0xffffffffu, r, o);
    }
    __syncthreads();
    return r; // valid in thread 0
}
__global__ void k_final_sum(const double* __restrict__ partial, int n, double* out, double scale)
{
    double s = 0;
    for (int i = threadIdx.x; i < n; i += RED_BT) s += partial[i];
    s = block_sum<RED_BT>(s);
    if (threadIdx.x == 0) *out = s * scale;
}

// atomic min / max on doubles through their ordered integer image
__device__ __forceinline__ long long dbl_ordered(double d)
{
    const long long b = __double_as_longlong(d);
    return b >= 0 ? b : (b ^ 0x7fffffffffffffffLL);
}
__device__ __forceinline__ double ordered_dbl(long long o)
{
    return __longlong_as_double(o >= 0 ? o : (o ^ 0x7fffffffffffffffLL));
}
__host__ __device__ __forceinline__ long long dbl_ordered_h(double d)
{
    long long b;
    memcpy(&b, &d, 8);
    return b >= 0 ? b : (b ^ 0x7fffffffffffffffLL);
}

} // namespace cipc
