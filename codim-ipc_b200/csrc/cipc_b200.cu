// cipc_b200.cu -- B200-native (sm_100a, fp64) C-IPC contact hot path behind the C ABI of
// include/cipc_b200.h.  See DESIGN.md for the data layout and the per-kernel rooflines.
//
// Pipeline (one CUDA stream, everything resident in HBM):
//   positions X / rest X0 / search dir P : double4 per node (32 B = one DRAM sector per gather)
//   broad phase   : per-primitive integer voxel boxes -> (cell,kind | prim) entries -> LSD radix sort
//                   -> per-cell kind ranges -> pair enumeration with the "min-corner" rule (each
//                   overlapping pair is visited in exactly one cell) + topology filters + exact AABB test
//   constraint set: closest-feature classification + d < dHat^2 per candidate, PP/PE de-duplication
//                   through an index hash table, output in the reference's stencil encoding
//   barrier       : E (tree reduction), g (fp64 atomics), H (12x12 / 9x9 / 6x6 blocks, PSD projection)
//   step size     : swept voxel boxes mirroring SPATIAL_HASH.h:432-622, ACCD per pair, atomic min
#include "../../include/cipc_b200.h"
#include "geom.cuh"
#include "eig.cuh"
#include "hess.cuh"
#include "prims.cuh"
#include "hostpool.h"

#include <cooperative_groups.h>
#include <immintrin.h>
#include <atomic>
#include <chrono>
#include <thread>
#include <algorithm>
#include <cmath>
#include <memory>

namespace cg = cooperative_groups;

namespace cipc {

long g_launches = 0;

// ===================================================================== device-side views
struct Topo {
    int nV, nBN, nBE, nBT, nRod, codim0, codim1;
    const int* BN;
    const int2* BE;
    const int4* BT;
    const uint8_t* flags; // bit0: DBC, bit1: node has NNExclusion entries
    const int* v2sv;      // vertex -> (last) boundary-node slot, as SPATIAL_HASH.h:100 / :489
    const u64* nnx;       // sorted (key<<32 | member)
    int nNnx;
};
struct GridDesc {
    double ox, oy, oz, inv; // cell = floor((x - o) * inv)
    int gx, gy, gz;         // cells per axis (linear index = ix + gx*(iy + gy*iz))
    int fsh;                // log2 of the fine steps per voxel (quantised-AABB pre-test)
};

__device__ __forceinline__ xv3 ldx(const double4* __restrict__ A, int v)
{
    const double4 q = A[v];
    return xv3(xd(q.x), xd(q.y), xd(q.z));
}
__device__ __forceinline__ dv3 ldd(const double4* __restrict__ A, int v)
{
    const double4 q = A[v];
    return dv3(q.x, q.y, q.z);
}
__device__ __forceinline__ bool nnx_has(const Topo& T, int v, int a)
{
    const u64 key = ((u64)(u32)v << 32) | (u32)a;
    int lo = 0, hi = T.nNnx;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        const u64 m = T.nnx[mid];
        if (m < key) lo = mid + 1;
        else hi = mid;
    }
    return lo < T.nNnx && T.nnx[lo] == key;
}
// IPC.h:171-181 / :1984-1996
__device__ __forceinline__ bool pt_pair_ok(const Topo& T, int vI, const int4& t)
{
    if (vI == t.x || vI == t.y || vI == t.z) return false;
    const uint8_t fv = T.flags[vI];
    if ((fv & 1) && (T.flags[t.x] & 1) && (T.flags[t.y] & 1) && (T.flags[t.z] & 1)) return false;
    if ((fv & 2) && (nnx_has(T, vI, t.x) || nnx_has(T, vI, t.y) || nnx_has(T, vI, t.z))) return false;
    return true;
}
// IPC.h:384-397 / :2202-2216
__device__ __forceinline__ bool ee_pair_ok(const Topo& T, const int2& a, const int2& b)
{
    if (a.x == b.x || a.x == b.y || a.y == b.x || a.y == b.y) return false;
    const uint8_t f0 = T.flags[a.x], f1 = T.flags[a.y];
    if ((f0 & 1) && (f1 & 1) && (T.flags[b.x] & 1) && (T.flags[b.y] & 1)) return false;
    if ((f0 & 2) && (nnx_has(T, a.x, b.x) || nnx_has(T, a.x, b.y))) return false;
    if ((f1 & 2) && (nnx_has(T, a.y, b.x) || nnx_has(T, a.y, b.y))) return false;
    return true;
}

// packed voxel coordinates: 21 bits per axis
__device__ __forceinline__ u64 pack3(int x, int y, int z) { return (u64)(u32)x | ((u64)(u32)y << 21) | ((u64)(u32)z << 42); }
__device__ __forceinline__ int ux(u64 p) { return (int)(p & 0x1fffffu); }
__device__ __forceinline__ int uy(u64 p) { return (int)((p >> 21) & 0x1fffffu); }
__device__ __forceinline__ int uz(u64 p) { return (int)((p >> 42) & 0x1fffffu); }
__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
// quantised-AABB pre-test resolution: GridDesc::fs = 2^fsh steps per voxel, 256 unless the grid is so elongated that
// fs x finer coordinates would not fit the 21 bits per axis of the packed boxes (size_grid)

// ===================================================================== reductions for the grids
__global__ void k_edge_len_partial(const double4* __restrict__ X, const int2* __restrict__ BE, int nBE, double* partial)
{
    double s = 0;
    for (int e = blockIdx.x * RED_BT + threadIdx.x; e < nBE; e += RED_GRID * RED_BT) {
        const int2 ed = BE[e];
        s += sqrt((double)norm2(ldx(X, ed.x) - ldx(X, ed.y)));
    }
    s = block_sum<RED_BT>(s);
    if (threadIdx.x == 0) partial[blockIdx.x] = s;
}
// sum over boundary-node slots of |p_x|+|p_y|+|p_z| (SPATIAL_HASH.h:466-474)
__global__ void k_psize_partial(const double4* __restrict__ P, const int* __restrict__ BN, int nBN, double* partial)
{
    double s = 0;
    for (int i = blockIdx.x * RED_BT + threadIdx.x; i < nBN; i += RED_GRID * RED_BT) {
        const double4 p = P[BN[i]];
        s += (fabs(p.x) + fabs(p.y)) + fabs(p.z);
    }
    s = block_sum<RED_BT>(s);
    if (threadIdx.x == 0) partial[blockIdx.x] = s;
}
// bbox over boundary nodes of x (and x + alpha p when P != nullptr); out[0..2]=min, out[3..5]=max (ordered ints)
__global__ void k_bbox(const double4* __restrict__ X, const double4* __restrict__ P, double alpha, const int* __restrict__ BN,
    int nBN, long long* out)
{
    double mn[3] = {1e300, 1e300, 1e300}, mx[3] = {-1e300, -1e300, -1e300};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nBN; i += gridDim.x * blockDim.x) {
        const int v = BN[i];
        const double4 x = X[v];
        double c[3] = {x.x, x.y, x.z};
        for (int d = 0; d < 3; ++d) { mn[d] = fmin(mn[d], c[d]); mx[d] = fmax(mx[d], c[d]); }
        if (P) {
            const double4 p = P[v];
            const double e[3] = {__dadd_rn(x.x, __dmul_rn(alpha, p.x)), __dadd_rn(x.y, __dmul_rn(alpha, p.y)),
                __dadd_rn(x.z, __dmul_rn(alpha, p.z))};
            for (int d = 0; d < 3; ++d) { mn[d] = fmin(mn[d], e[d]); mx[d] = fmax(mx[d], e[d]); }
        }
    }
    // warp -> block (shared memory) -> six atomics per block: the six addresses are shared by the whole grid
    __shared__ double sm[32][6];
    for (int d = 0; d < 3; ++d) {
        for (int o = 16; o > 0; o >>= 1) {
            mn[d] = fmin(mn[d], __shfl_down_sync(0xffffffffu, mn[d], o));
            mx[d] = fmax(mx[d], __shfl_down_sync(0xffffffffu, mx[d], o));
        }
        if ((threadIdx.x & 31) == 0) { sm[threadIdx.x >> 5][d] = mn[d]; sm[threadIdx.x >> 5][3 + d] = mx[d]; }
    }
    __syncthreads();
    if (threadIdx.x < 6) {
        const int nw = (blockDim.x + 31) >> 5;
        double v = sm[0][threadIdx.x];
        for (int w = 1; w < nw; ++w) v = threadIdx.x < 3 ? fmin(v, sm[w][threadIdx.x]) : fmax(v, sm[w][threadIdx.x]);
        if (threadIdx.x < 3) atomicMin(&out[threadIdx.x], dbl_ordered(v));
        else atomicMax(&out[threadIdx.x], dbl_ordered(v));
    }
}

// ===================================================================== voxel boxes
// Constraint-set grid (our own; any superset of the AABB-gap test is valid):
// every primitive's AABB inflated by r (~dHat/2).  prim ids: nodes [0,nBN), edges, triangles.
__global__ void k_boxes_ccs(Topo T, const double4* __restrict__ X, GridDesc G, double r, u64* boxLo, u64* boxHi, ulonglong2* fine, u32* cnt)
{
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    const int nP = T.nBN + T.nBE + T.nBT;
    if (g >= nP) return;
    double lo[3], hi[3];
    auto acc = [&](int v, bool first) {
        const double4 q = X[v];
        const double c[3] = {q.x, q.y, q.z};
        for (int d = 0; d < 3; ++d) {
            lo[d] = first ? c[d] : fmin(lo[d], c[d]);
            hi[d] = first ? c[d] : fmax(hi[d], c[d]);
        }
    };
    if (g < T.nBN) acc(T.BN[g], true);
    else if (g < T.nBN + T.nBE) { const int2 e = T.BE[g - T.nBN]; acc(e.x, true); acc(e.y, false); }
    else { const int4 t = T.BT[g - T.nBN - T.nBE]; acc(t.x, true); acc(t.y, false); acc(t.z, false); }
    const double o[3] = {G.ox, G.oy, G.oz};
    const int gd[3] = {G.gx, G.gy, G.gz};
    int l[3], h[3], fl[3], fh[3];
    for (int d = 0; d < 3; ++d) {
        // fs-times finer coordinates of the same inflated box; the voxel index is the fine index / fs
        const double fs = (double)(1 << G.fsh);
        fl[d] = clampi((int)floor((lo[d] - r - o[d]) * G.inv * fs), 0, (gd[d] << G.fsh) - 1);
        fh[d] = clampi((int)floor((hi[d] + r - o[d]) * G.inv * fs), 0, (gd[d] << G.fsh) - 1);
        l[d] = fl[d] >> G.fsh;
        h[d] = fh[d] >> G.fsh;
    }
    boxLo[g] = pack3(l[0], l[1], l[2]);
    boxHi[g] = pack3(h[0], h[1], h[2]);
    fine[g] = make_ulonglong2(pack3(fl[0], fl[1], fl[2]), pack3(fh[0], fh[1], fh[2]));
    cnt[g] = (u32)(h[0] - l[0] + 1) * (u32)(h[1] - l[1] + 1) * (u32)(h[2] - l[2] + 1);
}
// Swept grid of the step-size search: per boundary-node slot, the voxel range of
// [min(x, x+a p) - xi/2, max(x, x+a p) + xi/2]  (SPATIAL_HASH.h:522-529), arithmetic order kept.
__global__ void k_node_boxes_ccd(Topo T, const double4* __restrict__ X, const double4* __restrict__ P, double alpha, double halfXi,
    double guard, GridDesc G, u64* nodeLo, u64* nodeHi, ulonglong2* nodeFine)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= T.nBN) return;
    const int v = T.BN[i];
    const double4 x = X[v], p = P[v];
    const double c[3] = {x.x, x.y, x.z};
    const double e[3] = {__dadd_rn(x.x, __dmul_rn(alpha, p.x)), __dadd_rn(x.y, __dmul_rn(alpha, p.y)), __dadd_rn(x.z, __dmul_rn(alpha, p.z))};
    const double f[3] = {__dadd_rn(x.x, p.x), __dadd_rn(x.y, p.y), __dadd_rn(x.z, p.z)}; // full step, as the swept AABB test uses it
    const double o[3] = {G.ox, G.oy, G.oz};
    const int gd[3] = {G.gx, G.gy, G.gz};
    int l[3], h[3], fl[3], fh[3];
    for (int d = 0; d < 3; ++d) {
        const double mn = __dsub_rn(fmin(c[d], e[d]), halfXi), mx = __dadd_rn(fmax(c[d], e[d]), halfXi);
        l[d] = clampi((int)floor(__dmul_rn(__dsub_rn(mn, o[d]), G.inv)), 0, gd[d] - 1);
        h[d] = clampi((int)floor(__dmul_rn(__dsub_rn(mx, o[d]), G.inv)), 0, gd[d] - 1);
        // conservative quantised full-step box (any pair passing the swept AABB test with gap xi overlaps here)
        const double fs = (double)(1 << G.fsh);
        fl[d] = clampi((int)floor((fmin(c[d], f[d]) - halfXi - guard - o[d]) * G.inv * fs), 0, (gd[d] << G.fsh) - 1);
        fh[d] = clampi((int)floor((fmax(c[d], f[d]) + halfXi + guard - o[d]) * G.inv * fs), 0, (gd[d] << G.fsh) - 1);
    }
    nodeLo[i] = pack3(l[0], l[1], l[2]);
    nodeHi[i] = pack3(h[0], h[1], h[2]);
    nodeFine[i] = make_ulonglong2(pack3(fl[0], fl[1], fl[2]), pack3(fh[0], fh[1], fh[2]));
}
// edges / triangles: union of their vertices' node boxes (SPATIAL_HASH.h:555-592)
__global__ void k_prim_boxes_ccd(Topo T, const u64* __restrict__ nodeLo, const u64* __restrict__ nodeHi, const ulonglong2* __restrict__ nodeFine,
    u64* boxLo, u64* boxHi, ulonglong2* fine, u32* cnt)
{
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    const int nP = T.nBN + T.nBE + T.nBT;
    if (g >= nP) return;
    int l[3], h[3], fl[3], fh[3];
    auto acc = [&](int sv, bool first) {
        const u64 a = nodeLo[sv], b = nodeHi[sv];
        const ulonglong2 q = nodeFine[sv];
        const int al[3] = {ux(a), uy(a), uz(a)}, bh[3] = {ux(b), uy(b), uz(b)};
        const int ql[3] = {ux(q.x), uy(q.x), uz(q.x)}, qh[3] = {ux(q.y), uy(q.y), uz(q.y)};
        for (int d = 0; d < 3; ++d) {
            l[d] = first ? al[d] : min(l[d], al[d]);
            h[d] = first ? bh[d] : max(h[d], bh[d]);
            fl[d] = first ? ql[d] : min(fl[d], ql[d]);
            fh[d] = first ? qh[d] : max(fh[d], qh[d]);
        }
    };
    if (g < T.nBN) acc(g, true);
    else if (g < T.nBN + T.nBE) { const int2 e = T.BE[g - T.nBN]; acc(T.v2sv[e.x], true); acc(T.v2sv[e.y], false); }
    else { const int4 t = T.BT[g - T.nBN - T.nBE]; acc(T.v2sv[t.x], true); acc(T.v2sv[t.y], false); acc(T.v2sv[t.z], false); }
    boxLo[g] = pack3(l[0], l[1], l[2]);
    boxHi[g] = pack3(h[0], h[1], h[2]);
    fine[g] = make_ulonglong2(pack3(fl[0], fl[1], fl[2]), pack3(fh[0], fh[1], fh[2]));
    cnt[g] = (u32)(h[0] - l[0] + 1) * (u32)(h[1] - l[1] + 1) * (u32)(h[2] - l[2] + 1);
}
// one (cell<<2|kind, local id) entry per covered voxel
// Multi-GPU: a rank owns a slab [s0, s1) of voxel indices along `axis`; only entries inside it are emitted and sorted.
struct Slab {
    int axis, s0, s1;
};
__device__ __forceinline__ int uax(u64 p, int axis) { return axis == 0 ? ux(p) : (axis == 1 ? uy(p) : uz(p)); }
// entries per voxel layer along the slab axis (summed over all primitives): the host balances slabs on it.
// The few hundred bins would serialise millions of global atomics on a handful of L2 lines, so each CTA first
// accumulates in shared memory (consecutive primitives are spatial neighbours: a CTA touches a few bins) and flushes
// only its non-zero bins.
constexpr int SLAB_SMEM_BINS = 8192;
__global__ void __launch_bounds__(256) k_slab_hist(int nP, const u64* __restrict__ boxLo, const u64* __restrict__ boxHi, int axis, int nl, u32* hist)
{
    __shared__ u32 sh[SLAB_SMEM_BINS];
    const bool priv = nl <= SLAB_SMEM_BINS;
    if (priv) {
        for (int i = threadIdx.x; i < nl; i += blockDim.x) sh[i] = 0;
        __syncthreads();
    }
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g < nP) {
        const u64 a = boxLo[g], b = boxHi[g];
        const u32 ex[3] = {(u32)(ux(b) - ux(a) + 1), (u32)(uy(b) - uy(a) + 1), (u32)(uz(b) - uz(a) + 1)};
        const u32 per = (axis == 0) ? ex[1] * ex[2] : (axis == 1 ? ex[0] * ex[2] : ex[0] * ex[1]);
        for (int i = uax(a, axis); i <= uax(b, axis); ++i) atomicAdd(priv ? &sh[i] : &hist[i], per);
    }
    if (priv) {
        __syncthreads();
        for (int i = threadIdx.x; i < nl; i += blockDim.x) {
            const u32 v = sh[i];
            if (v) atomicAdd(&hist[i], v);
        }
    }
}
__global__ void k_clip_counts(int nP, const u64* __restrict__ boxLo, const u64* __restrict__ boxHi, Slab sl, u32* cnt)
{
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= nP) return;
    const u64 a = boxLo[g], b = boxHi[g];
    int l[3] = {ux(a), uy(a), uz(a)}, h[3] = {ux(b), uy(b), uz(b)};
    l[sl.axis] = max(l[sl.axis], sl.s0);
    h[sl.axis] = min(h[sl.axis], sl.s1 - 1);
    cnt[g] = (h[sl.axis] < l[sl.axis]) ? 0u : (u32)(h[0] - l[0] + 1) * (u32)(h[1] - l[1] + 1) * (u32)(h[2] - l[2] + 1);
}
// one (cell<<2|kind, local id) entry per covered voxel (inside the rank's slab)
__global__ void k_emit_entries(Topo T, GridDesc G, const u64* __restrict__ boxLo, const u64* __restrict__ boxHi,
    const u32* __restrict__ off, Slab sl, u32* keys, u32* vals)
{
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    const int nP = T.nBN + T.nBE + T.nBT;
    if (g >= nP) return;
    u32 kind, id;
    if (g < T.nBN) { kind = 0; id = g; }
    else if (g < T.nBN + T.nBE) { kind = 1; id = g - T.nBN; }
    else { kind = 2; id = g - T.nBN - T.nBE; }
    const u64 a = boxLo[g], b = boxHi[g];
    int l[3] = {ux(a), uy(a), uz(a)}, h[3] = {ux(b), uy(b), uz(b)};
    l[sl.axis] = max(l[sl.axis], sl.s0);
    h[sl.axis] = min(h[sl.axis], sl.s1 - 1);
    u32 o = off[g];
    for (int iz = l[2]; iz <= h[2]; ++iz)
        for (int iy = l[1]; iy <= h[1]; ++iy)
            for (int ix = l[0]; ix <= h[0]; ++ix) {
                const u32 cell = (u32)ix + (u32)G.gx * ((u32)iy + (u32)G.gy * (u32)iz);
                keys[o] = (cell << 2) | kind;
                vals[o] = id;
                ++o;
            }
}
// ---- dense cell table (counting sort): when the grid is small enough to index every cell (DENSE_CELL_LIMIT), the
// cell-sorted entry lists are built without a radix sort: count entries per (cell, kind) bin with atomics, scan the bins
// -- the scanned table IS the kind-range table ks[cell*4 + k] -- and scatter every entry to its bin with a running
// per-bin counter, writing the cell-local pair code on the way.  Entries of one (cell, kind) run come out in arbitrary
// order, which the pair enumeration does not depend on (pairs are normalised when they are emitted).  Three launches and
// one host round trip replace ~25 launches and two round trips.
constexpr long DENSE_CELL_LIMIT = 16L << 20;
__global__ void k_cell_count(Topo T, GridDesc G, const u64* __restrict__ boxLo, const u64* __restrict__ boxHi, Slab sl, u32* __restrict__ bins)
{
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    const int nP = T.nBN + T.nBE + T.nBT;
    if (g >= nP) return;
    const u32 kind = g < T.nBN ? 0u : (g < T.nBN + T.nBE ? 1u : 2u);
    const u64 a = boxLo[g], b = boxHi[g];
    int l[3] = {ux(a), uy(a), uz(a)}, h[3] = {ux(b), uy(b), uz(b)};
    l[sl.axis] = max(l[sl.axis], sl.s0);
    h[sl.axis] = min(h[sl.axis], sl.s1 - 1);
    for (int iz = l[2]; iz <= h[2]; ++iz)
        for (int iy = l[1]; iy <= h[1]; ++iy)
            for (int ix = l[0]; ix <= h[0]; ++ix) {
                const u32 cell = (u32)ix + (u32)G.gx * ((u32)iy + (u32)G.gy * (u32)iz);
                atomicAdd(&bins[cell * 4u + kind], 1u);
            }
}
__device__ __forceinline__ uint2 entry_code(const int* c, const int* l, const int* fl, const int* fh, int fsh);
__global__ void k_cell_fill(Topo T, GridDesc G, const u64* __restrict__ boxLo, const u64* __restrict__ boxHi, const ulonglong2* __restrict__ fine,
    Slab sl, u32* __restrict__ next, u32* __restrict__ vals, uint2* __restrict__ codes)
{
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    const int nP = T.nBN + T.nBE + T.nBT;
    if (g >= nP) return;
    u32 kind, id;
    if (g < T.nBN) { kind = 0; id = g; }
    else if (g < T.nBN + T.nBE) { kind = 1; id = g - T.nBN; }
    else { kind = 2; id = g - T.nBN - T.nBE; }
    const u64 a = boxLo[g], b = boxHi[g];
    const int l0[3] = {ux(a), uy(a), uz(a)};
    int l[3] = {l0[0], l0[1], l0[2]}, h[3] = {ux(b), uy(b), uz(b)};
    l[sl.axis] = max(l[sl.axis], sl.s0);
    h[sl.axis] = min(h[sl.axis], sl.s1 - 1);
    if (l[sl.axis] > h[sl.axis]) return; // outside this rank's slab: leave before touching the fine coordinates
    const ulonglong2 f = fine[g];
    const int fl[3] = {ux(f.x), uy(f.x), uz(f.x)}, fh[3] = {ux(f.y), uy(f.y), uz(f.y)};
    for (int iz = l[2]; iz <= h[2]; ++iz)
        for (int iy = l[1]; iy <= h[1]; ++iy)
            for (int ix = l[0]; ix <= h[0]; ++ix) {
                const u32 cell = (u32)ix + (u32)G.gx * ((u32)iy + (u32)G.gy * (u32)iz);
                const u32 pos = atomicAdd(&next[cell * 4u + kind], 1u);
                const int c[3] = {ix, iy, iz};
                vals[pos] = id;
                codes[pos] = entry_code(c, l0, fl, fh, G.fsh);
            }
}
__global__ void k_cell_heads(const u32* __restrict__ keys, u32 n, u32* heads)
{
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    heads[i] = (i == 0 || (keys[i] >> 2) != (keys[i - 1] >> 2)) ? 1u : 0u;
}
// ks[c*4 + k] = first entry of cell c whose kind is >= k (k = 3: end of the cell run)
__global__ void k_kind_starts(const u32* __restrict__ keys, const u32* __restrict__ headScan, const u32* __restrict__ heads, u32 n, u32* ks)
{
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n) return;
    const bool end = (i == n);
    const u32 ki = end ? 0 : (keys[i] & 3u);
    const u32 ci = end ? 0 : (headScan[i] + heads[i] - 1u);
    if (i == 0) {
        for (u32 k = 0; k <= ki; ++k) ks[ci * 4 + k] = 0;
        return;
    }
    const u32 kp = keys[i - 1] & 3u;
    const u32 cp = headScan[i - 1] + heads[i - 1] - 1u;
    if (end || heads[i]) {
        for (u32 k = kp + 1; k <= 3; ++k) ks[cp * 4 + k] = i;
        if (!end) for (u32 k = 0; k <= ki; ++k) ks[ci * 4 + k] = i;
    }
    else if (ki != kp) {
        for (u32 k = kp + 1; k <= ki; ++k) ks[ci * 4 + k] = i;
    }
}

// ===================================================================== candidate pairs
struct CandOut {
    int2* buf[4]; // 0 PT (slot, tri)  1 EE (eI, eJ)  2 PE (slot, edge)  3 PP (slotI, slotJ)
    u32* count;   // 4 counters
    u32 cap[4];
};
__device__ __forceinline__ void cand_push(const CandOut& o, int which, int a, int b)
{
    cg::coalesced_group g = cg::coalesced_threads();
    u32 base = 0;
    if (g.thread_rank() == 0) base = atomicAdd(&o.count[which], g.size());
    base = g.shfl(base, 0);
    const u32 idx = base + g.thread_rank();
    if (idx < o.cap[which]) o.buf[which][idx] = make_int2(a, b);
}
// ---- per-entry pair codes.  A pair of boxes that share a cell is handled only in the cell that is the minimum corner
// of the boxes' intersection ("min-corner rule": no pair is visited twice, no de-duplication pass).  Both cheap pair
// tests only need data that is LOCAL to the (cell, primitive) entry:
//   * min corner:  cell == max(loA, loB) per axis  <=>  per axis, A or B STARTS in this cell (both cover it)  -> 3 bits
//   * quantised-AABB overlap (fs <= 256 steps per voxel): with both fine intervals clamped to the cell's own fine
//     range [fs c, fs c + fs-1] the test is unchanged, because each interval reaches into the cell from both sides
//     (cell(fine lo) <= box lo <= c <= box hi <= cell(fine hi))                                                   -> 6 x 8 bits
// code.x = lo fields (x at bit 0, y at bit 10, z at bit 20, 8 bits each); code.y = hi fields at the same positions, a
// guard bit above each field (bits 8, 18, 28) and the start bits at 29..31.  "lo_A <= hi_B on all three axes" is then one
// subtraction: (hi_B | guards) - lo_A keeps every guard bit exactly when no field borrows.  One coalesced 8-byte load
// and ~7 integer instructions per pair replace the 24 bytes of gathers of the global test (byte-wise SIMD compares are
// emulated with a dozen instructions each on this architecture and were 70% of the kernel's instruction count).
constexpr u32 CODE_GUARDS = (1u << 8) | (1u << 18) | (1u << 28);
__global__ void k_entry_codes(const u32* __restrict__ keys, const u32* __restrict__ vals, u32 nE, int nBN, int nBE,
    const u64* __restrict__ boxLo, const ulonglong2* __restrict__ fine, GridDesc G, uint2* __restrict__ codes)
{
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nE) return;
    const u32 key = keys[i], cell = key >> 2, kind = key & 3u;
    const u32 g = vals[i] + (kind == 0 ? 0u : (kind == 1 ? (u32)nBN : (u32)(nBN + nBE)));
    const u32 cxy = (u32)G.gx * (u32)G.gy;
    const u32 cz = cell / cxy, rem = cell - cz * cxy, cy = rem / (u32)G.gx, cx = rem - cy * (u32)G.gx;
    const u64 lo = boxLo[g];
    const ulonglong2 f = fine[g];
    const int c[3] = {(int)cx, (int)cy, (int)cz};
    const int l[3] = {ux(lo), uy(lo), uz(lo)}, fl[3] = {ux(f.x), uy(f.x), uz(f.x)}, fh[3] = {ux(f.y), uy(f.y), uz(f.y)};
    codes[i] = entry_code(c, l, fl, fh, G.fsh);
}
// pair code of primitive (voxel box lower corner l, fine box fl..fh) as an entry of cell c
__device__ __forceinline__ uint2 entry_code(const int* c, const int* l, const int* fl, const int* fh, int fsh)
{
    u32 a = 0, b = 0;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        a |= (u32)clampi(fl[d] - (c[d] << fsh), 0, (1 << fsh) - 1) << (10 * d);
        b |= (u32)clampi(fh[d] - (c[d] << fsh), 0, (1 << fsh) - 1) << (10 * d);
        if (l[d] == c[d]) b |= 1u << (29 + d);
    }
    return make_uint2(a, b | CODE_GUARDS);
}
__device__ __forceinline__ bool code_pair_ok(const uint2 a, const uint2 b)
{
    const u32 t = (b.y - a.x) & (a.y - b.x) & CODE_GUARDS; // guards survive <=> lo_A <= hi_B and lo_B <= hi_A per axis
    return t == CODE_GUARDS && (a.y | b.y) >= 0xE0000000u; // and on every axis one of the two boxes starts in this cell
}

// Work items of the pair enumeration: a cell's point queries and edge queries are cut into tasks of `qch` queries,
// one warp per task, so a crowded cell (hundreds of particles in one voxel of a granular pile) is spread over many
// warps instead of serialising on one.  cnt[c] = tasks of cell c; desc[] = (cell, task index within the cell).
// qch is chosen per hash build (pairs_qch): few, crowded cells are cut finely; with many cells a task is a whole cell
// side, because every task ends with a partially filled batch of the expensive filters.
__device__ __forceinline__ u32 cell_point_tasks(u32 p0, u32 e0, u32 qch) { return (e0 - p0 + qch - 1) / qch; }
__device__ __forceinline__ u32 cell_edge_tasks(u32 e0, u32 t0, u32 qch) { return t0 > e0 + 1 ? (t0 - e0 - 1 + qch - 1) / qch : 0u; }
// Point tasks of all cells come first, then the edge tasks (cnt has 2 nCells entries, one scan): the warps of a CTA then
// work on tasks of one kind and similar length.
__global__ void k_task_counts(const u32* __restrict__ ks, u32 nCells, u32 qch, u32* __restrict__ cnt)
{
    const u32 c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nCells) return;
    const u32 p0 = ks[c * 4], e0 = ks[c * 4 + 1], t0 = ks[c * 4 + 2];
    cnt[c] = cell_point_tasks(p0, e0, qch);
    cnt[nCells + c] = cell_edge_tasks(e0, t0, qch);
}
__global__ void k_task_desc(const u32* __restrict__ ks, u32 nCells, u32 qch, const u32* __restrict__ off, uint2* __restrict__ desc)
{
    const u32 c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nCells) return;
    const u32 p0 = ks[c * 4], e0 = ks[c * 4 + 1], t0 = ks[c * 4 + 2];
    const u32 np = cell_point_tasks(p0, e0, qch), ne = cell_edge_tasks(e0, t0, qch);
    const u32 op = off[c], oe = off[nCells + c];
    for (u32 k = 0; k < np; ++k) desc[op + k] = make_uint2(c, k);
    for (u32 k = 0; k < ne; ++k) desc[oe + k] = make_uint2(c, k | 0x80000000u);
}
// One WARP per task (<= qch point queries or edge queries of one voxel cell).  Phase 1 (cheap, all lanes busy): the
// queries are walked in order and the 32 lanes test 32 targets of the run at a time on the entry codes; survivors
// are appended to a per-warp ring buffer in shared memory.  Phase 2 (expensive, all lanes busy): whenever the buffer
// holds 32 pairs, each lane takes one: topology filters, coordinate gathers and the reference's exact AABB test.
// Without the queue the expensive path ran for every chunk with ~4 of 32 lanes active.
// CCD=false: constraint-set pass (gap test with dist = dHat).  CCD=true: step-size pass (swept AABB test with
// dist = thickness, full search direction).
constexpr int PAIRS_WARPS = 8;
template <bool CCD>
__global__ void __launch_bounds__(PAIRS_WARPS * 32) k_pairs(Topo T, const double4* __restrict__ X, const double4* __restrict__ P, double dist_,
    const u32* __restrict__ vals, const u32* __restrict__ ks, const uint2* __restrict__ codes, const uint2* __restrict__ desc,
    const u32* __restrict__ nTasks, u32 qch, CandOut out)
{
    __shared__ uint2 sq[PAIRS_WARPS][64];
    __shared__ int2 so[PAIRS_WARPS][96];
    const u32 warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const u32 task = blockIdx.x * PAIRS_WARPS + warp;
    if (task >= *nTasks) return; // whole warp leaves; no block-level barrier below
    const uint2 td = desc[task];
    const u32 cellIdx = td.x;
    uint2* q = sq[warp];
    int2* ob = so[warp];
    u32 head = 0, tail = 0; // warp-uniform ring indices of the pair queue
    u32 on = 0;             // warp-uniform fill of the output buffer
    const u32 p0 = ks[cellIdx * 4], e0 = ks[cellIdx * 4 + 1], t0 = ks[cellIdx * 4 + 2], end = ks[cellIdx * 4 + 3];
    const bool pointTask = (td.y & 0x80000000u) == 0u;
    const u32 tk = td.y & 0x7fffffffu;
    // this task's query range [qa, qb): point entries of the cell, or edge entries (the last edge has no later partner)
    const u32 qa = (pointTask ? p0 : e0) + tk * qch;
    const u32 qb = pointTask ? min(qa + qch, e0) : min(qa + qch, t0 - 1);
    const xd dist(dist_);
    const u32 ltmask = (1u << lane) - 1u;
    auto push = [&](bool pass, u32 a, u32 b) { // called by all 32 lanes
        const u32 m = __ballot_sync(0xffffffffu, pass);
        if (pass) q[(tail + __popc(m & ltmask)) & 63u] = make_uint2(a, b);
        tail += __popc(m);
    };
    // Candidates are collected per warp in shared memory and handed to the global list 64 or more at a time: one
    // atomic on the list counter per flush instead of one per batch (millions of same-address atomics serialise in L2).
    auto flush = [&](int which) { // called by all 32 lanes
        if (on == 0) return;
        u32 base = 0;
        if (lane == 0) base = atomicAdd(&out.count[which], on);
        base = __shfl_sync(0xffffffffu, base, 0);
        __syncwarp();
        for (u32 k = lane; k < on; k += 32)
            if (base + k < out.cap[which]) out.buf[which][base + k] = ob[k];
        __syncwarp();
        on = 0;
    };
    auto emit = [&](bool ok, int a, int b, int which) { // called by all 32 lanes
        const u32 m = __ballot_sync(0xffffffffu, ok);
        if (ok) ob[on + __popc(m & ltmask)] = make_int2(a, b);
        on += __popc(m);
        if (on >= 64u) flush(which);
    };
    // (a, b) = sorted-entry indices of a point and a triangle
    auto do_pt = [&](u32 a, u32 b, int& ra, int& rb) -> bool {
        const int svI = (int)vals[a], t = (int)vals[b];
        const int vI = T.BN[svI];
        const int4 tri = T.BT[t];
        ra = svI; rb = t;
        if (!pt_pair_ok(T, vI, tri)) return false;
        const xv3 p = ldx(X, vI), q0 = ldx(X, tri.x), q1 = ldx(X, tri.y), q2 = ldx(X, tri.z);
        if (CCD) return pt_ccd_broadphase(p, q0, q1, q2, ldx(P, vI), ldx(P, tri.x), ldx(P, tri.y), ldx(P, tri.z), dist);
        return pt_cd_broadphase(p, q0, q1, q2, dist);
    };
    auto do_ee = [&](u32 a, u32 b, int& ra, int& rb) -> bool {
        const int e0_ = (int)vals[a], e1_ = (int)vals[b];
        const int eI = min(e0_, e1_), eJ = max(e0_, e1_); // the reference pairs eI with eJ > eI; runs may be unordered
        const int2 ea = T.BE[eI], eb = T.BE[eJ];
        ra = eI; rb = eJ;
        if (!ee_pair_ok(T, ea, eb)) return false;
        const xv3 a0 = ldx(X, ea.x), a1 = ldx(X, ea.y), b0 = ldx(X, eb.x), b1 = ldx(X, eb.y);
        if (CCD) return ee_ccd_broadphase(a0, a1, b0, b1, ldx(P, ea.x), ldx(P, ea.y), ldx(P, eb.x), ldx(P, eb.y), dist);
        return ee_cd_broadphase(a0, a1, b0, b1, dist);
    };
    auto drain = [&](auto process, int which, bool all) {
        while (tail - head >= 32u || (all && tail != head)) {
            const u32 n = min(32u, tail - head);
            __syncwarp();
            bool ok = false;
            int ra = 0, rb = 0;
            uint2 e = make_uint2(0u, 0u);
            if (lane < n) e = q[(head + lane) & 63u];
            __syncwarp(); // the slots may be overwritten by the next push
            if (lane < n) ok = process(e.x, e.y, ra, rb);
            head += n;
            emit(ok, ra, rb, which);
        }
    };
    if (pointTask) {
        // ---- point queries against the triangle run
        for (u32 i = qa; i < qb; ++i) {
            const uint2 cA = codes[i];
            for (u32 jb = t0; jb < end; jb += 32) {
                const u32 j = jb + lane;
                push(j < end && code_pair_ok(cA, codes[j]), i, j);
                drain(do_pt, 0, false);
            }
        }
        drain(do_pt, 0, true);
        flush(0);
        // ---- codimensional extras: rod / particle points against rod edges (IPC.h:271-326; step size: particles only,
        //      :2098-2135) and particles against later boundary-node slots (IPC.h:328-352, :2137-2163)
        if (T.nRod > 0 || T.codim1 < T.nBN) {
            for (u32 i = qa; i < qb; ++i) {
                const int svI = (int)vals[i];
                if (svI < min(T.codim0, T.codim1)) continue; // warp-uniform
                const int vI = T.BN[svI];
                const uint2 cA = codes[i];
                const xv3 p = ldx(X, vI);
                xv3 dp;
                if (CCD) dp = ldx(P, vI);
                if (T.nRod > 0 && svI >= (CCD ? T.codim1 : T.codim0)) {
                    for (u32 jb = e0; jb < t0; jb += 32) {
                        const u32 j = jb + lane;
                        bool ok = j < t0 && code_pair_ok(cA, codes[j]);
                        int e = 0;
                        if (ok) {
                            e = (int)vals[j];
                            ok = e >= T.nBE - T.nRod;
                        }
                        if (ok) {
                            const int2 ed = T.BE[e];
                            ok = !(vI == ed.x || vI == ed.y) && !((T.flags[vI] & 1) && (T.flags[ed.x] & 1) && (T.flags[ed.y] & 1));
                            if (ok) {
                                const xv3 q0 = ldx(X, ed.x), q1 = ldx(X, ed.y);
                                if (CCD) ok = pe_ccd_broadphase(p, q0, q1, dp, ldx(P, ed.x), ldx(P, ed.y), dist);
                                else ok = pe_cd_broadphase(p, q0, q1, dist);
                            }
                        }
                        emit(ok, svI, e, 2);
                    }
                    flush(2);
                }
                if (svI >= T.codim1) { // slots ascend inside a run
                    for (u32 jb = i + 1; jb < e0; jb += 32) {
                        const u32 j = jb + lane;
                        bool ok = j < e0 && code_pair_ok(cA, codes[j]);
                        int svJ = 0;
                        if (ok) {
                            svJ = (int)vals[j];
                            const int vJ = T.BN[svJ];
                            ok = svJ >= T.codim1 && !((T.flags[vI] & 1) && (T.flags[vJ] & 1)); // particle-particle only (runs may be unordered)
                            if (CCD && ok) ok = pp_ccd_broadphase(p, ldx(X, vJ), dp, ldx(P, vJ), dist);
                        }
                        emit(ok, min(svI, svJ), max(svI, svJ), 3);
                    }
                    if (T.nRod > 0) flush(3); // the buffer holds one candidate kind at a time
                }
            }
            flush(3);
        }
        return;
    }
    // ---- edge queries: edge ids ascend inside a run, so j > i <=> eJ > eI
    for (u32 i = qa; i < qb; ++i) {
        const uint2 cA = codes[i];
        for (u32 jb = i + 1; jb < t0; jb += 32) {
            const u32 j = jb + lane;
            push(j < t0 && code_pair_ok(cA, codes[j]), i, j);
            drain(do_ee, 1, false);
        }
    }
    drain(do_ee, 1, true);
    flush(1);
}

// ===================================================================== constraint-set narrow phase
struct NarrowOut {
    int4* pass;   // PT / EE / mollified stencils (never duplicated)
    int4* raw;    // PP / PE stencils before de-duplication (c3 = -1)
    u32* count;   // [0] pass, [1] raw
};
__device__ __forceinline__ void push4(int4* buf, u32* ctr, int4 v)
{
    cg::coalesced_group g = cg::coalesced_threads();
    u32 base = 0;
    if (g.thread_rank() == 0) base = atomicAdd(ctr, g.size());
    base = g.shfl(base, 0);
    buf[base + g.thread_rank()] = v;
}
// Block-wide slot reservation in one of NL append-only lists: which in [0, NL) or -1 (nothing to append).  One atomic
// per list and CTA instead of one per warp -- the list counters are single addresses, and same-address atomics retire at
// about one per clock in L2, which bounded the narrow phase (a million warps, two counters).  EVERY thread of the CTA
// must call it.  Returns the element index in list `which`.
template <int NL, int BT>
__device__ __forceinline__ u32 block_slot(int which, u32* counters)
{
    __shared__ u32 wcnt[NL][BT / 32];
    __shared__ u32 sbase[NL];
    const u32 lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    u32 myRank = 0;
#pragma unroll
    for (int l = 0; l < NL; ++l) {
        const u32 m = __ballot_sync(0xffffffffu, which == l);
        if (which == l) myRank = __popc(m & ((1u << lane) - 1u));
        if (lane == 0) wcnt[l][warp] = __popc(m);
    }
    __syncthreads();
    if (threadIdx.x < NL) {
        u32 tot = 0;
        for (int w = 0; w < BT / 32; ++w) { const u32 t = wcnt[threadIdx.x][w]; wcnt[threadIdx.x][w] = tot; tot += t; }
        sbase[threadIdx.x] = tot ? atomicAdd(&counters[threadIdx.x], tot) : 0u;
    }
    __syncthreads();
    const u32 r = which >= 0 ? sbase[which] + wcnt[which][warp] + myRank : 0u;
    __syncthreads();
    return r;
}
constexpr int NARROW_BT = 256;
__device__ __forceinline__ void narrow_append(const NarrowOut& out, int which, const int4 r)
{
    const u32 slot = block_slot<2, NARROW_BT>(which, out.count);
    if (which == 0) out.pass[slot] = r;
    else if (which == 1) out.raw[slot] = r;
}
// IPC.h:189-257.  The closest-feature type selects the operands first, so a warp runs at most three distance bodies
// (point-point, point-edge, point-triangle) instead of one per type; the arithmetic of each body is unchanged.
__global__ void __launch_bounds__(NARROW_BT, 4) k_narrow_pt(Topo T, const double4* __restrict__ X, const int2* __restrict__ cand, u32 n, double dHat2_,
    NarrowOut out)
{
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    int which = -1;
    int4 r = make_int4(0, 0, 0, 0);
    if (i < n) {
        const int2 c = cand[i];
        const int vI = T.BN[c.x];
        const int4 t = T.BT[c.y];
        const xv3 p = ldx(X, vI), t0 = ldx(X, t.x), t1 = ldx(X, t.y), t2 = ldx(X, t.z);
        const xd dHat2(dHat2_);
        const int ty = pt_type(p, t0, t1, t2);
        xd d;
        if (ty == 6) { d = pt_dist2(p, t0, t1, t2); r = make_int4(-vI - 1, t.x, t.y, t.z); }
        else if (ty >= 3) {
            const bool e3 = ty == 3, e4 = ty == 4;
            const xv3 e0 = e3 ? t0 : (e4 ? t1 : t2), e1 = e3 ? t1 : (e4 ? t2 : t0);
            d = pe_dist2(p, e0, e1);
            r = make_int4(-vI - 1, e3 ? t.x : (e4 ? t.y : t.z), e3 ? t.y : (e4 ? t.z : t.x), -1);
        }
        else {
            const xv3 q = ty == 0 ? t0 : (ty == 1 ? t1 : t2);
            d = pp_dist2(p, q);
            r = make_int4(-vI - 1, ty == 0 ? t.x : (ty == 1 ? t.y : t.z), -1, -1);
        }
        if (d < dHat2) which = (ty == 6) ? 0 : 1;
    }
    narrow_append(out, which, r);
}
// rest length^2 per boundary edge: the mollifier threshold 1e-3 |ea|^2 |eb|^2 (EDGE_EDGE_MOLLIFIER.h:582-592) only
// depends on it, which saves the four rest-position gathers per edge-edge candidate
__global__ void k_rest_len2(const double4* __restrict__ X0, const int2* __restrict__ BE, int nBE, double* __restrict__ out)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nBE) return;
    const int2 ed = BE[e];
    out[e] = norm2(ldx(X0, ed.x) - ldx(X0, ed.y)).v;
}
// IPC.h:414-564
__global__ void __launch_bounds__(NARROW_BT, 4) k_narrow_ee(Topo T, const double4* __restrict__ X, const double* __restrict__ restLen2,
    const int2* __restrict__ cand, u32 n, double dHat2_, NarrowOut out)
{
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    int which = -1;
    int4 r = make_int4(0, 0, 0, 0);
    if (i < n) {
        const int2 c = cand[i];
        const int2 a = T.BE[c.x], b = T.BE[c.y];
        const xv3 a0 = ldx(X, a.x), a1 = ldx(X, a.y), b0 = ldx(X, b.x), b1 = ldx(X, b.y);
        const xd dHat2(dHat2_);
        const xd cn2 = ee_cross_norm2(a0, a1, b0, b1);
        const xd eps_x = xd(1.0e-3) * xd(restLen2[c.x]) * xd(restLen2[c.y]);
        const bool mol = cn2 < eps_x;
        const int ty = ee_type(a0, a1, b0, b1);
        xd d;
        if (ty == 8) {
            d = ee_dist2(a0, a1, b0, b1);
            r = mol ? make_int4(a.x, a.y, -b.x - 1, b.y) : make_int4(a.x, a.y, b.x, b.y);
        }
        else if (ty == 2 || ty >= 5) { // point-edge: 2 (a0; b), 5 (a1; b), 6 (b0; a), 7 (b1; a)
            const bool onB = ty <= 5;   // the edge is b
            const xv3 p = ty == 2 ? a0 : (ty == 5 ? a1 : (ty == 6 ? b0 : b1));
            const int pv = ty == 2 ? a.x : (ty == 5 ? a.y : (ty == 6 ? b.x : b.y));
            const int po = ty == 2 ? a.y : (ty == 5 ? a.x : (ty == 6 ? b.y : b.x)); // the other end of the point's edge
            d = pe_dist2(p, onB ? b0 : a0, onB ? b1 : a1);
            const int e0 = onB ? b.x : a.x, e1 = onB ? b.y : a.y;
            r = mol ? make_int4(pv, e0, e1, -po - 1) : make_int4(-pv - 1, e0, e1, -1);
        }
        else { // point-point: 0 (a0,b0), 1 (a0,b1), 3 (a1,b0), 4 (a1,b1)
            const bool fa = ty < 3, fb = (ty == 0 || ty == 3);
            d = pp_dist2(fa ? a0 : a1, fb ? b0 : b1);
            const int pa = fa ? a.x : a.y, oa = fa ? a.y : a.x, pb = fb ? b.x : b.y, ob = fb ? b.y : b.x;
            r = mol ? make_int4(pa, pb, -oa - 1, -ob - 1) : make_int4(-pa - 1, pb, -1, -1);
        }
        if (d < dHat2) which = (r.x >= 0) ? 0 : 1;
    }
    narrow_append(out, which, r);
}
// IPC.h:281-322
__global__ void __launch_bounds__(NARROW_BT) k_narrow_pe(Topo T, const double4* __restrict__ X, const int2* __restrict__ cand, u32 n, double dHat2_,
    NarrowOut out)
{
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    int which = -1;
    int4 r = make_int4(0, 0, 0, 0);
    if (i < n) {
        const int2 c = cand[i];
        const int vI = T.BN[c.x];
        const int2 e = T.BE[c.y];
        const xv3 p = ldx(X, vI), e0 = ldx(X, e.x), e1 = ldx(X, e.y);
        xd d;
        switch (pe_type(p, e0, e1)) {
        case 0: d = pp_dist2(p, e0); r = make_int4(-vI - 1, e.x, -1, -1); break;
        case 1: d = pp_dist2(p, e1); r = make_int4(-vI - 1, e.y, -1, -1); break;
        default: d = pe_dist2(p, e0, e1); r = make_int4(-vI - 1, e.x, e.y, -1); break;
        }
        if (d < xd(dHat2_)) which = 1;
    }
    narrow_append(out, which, r);
}
// IPC.h:336-349
__global__ void __launch_bounds__(NARROW_BT) k_narrow_pp(Topo T, const double4* __restrict__ X, const int2* __restrict__ cand, u32 n, double dHat2_,
    NarrowOut out)
{
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    int which = -1;
    int4 r = make_int4(0, 0, 0, 0);
    if (i < n) {
        const int2 c = cand[i];
        const int vI = T.BN[c.x], vJ = T.BN[c.y];
        const xd d = pp_dist2(ldx(X, vI), ldx(X, vJ));
        r = make_int4(-vI - 1, vJ, -1, -1);
        if (d < xd(dHat2_)) which = 1;
    }
    narrow_append(out, which, r);
}

// PP/PE de-duplication (IPC.h:599-654): same raw 4-tuple => one stencil with multiplicity.
// Index hash table: a slot holds the index of the first record that claimed it.
__device__ __forceinline__ u32 hash3(int a, int b, int c)
{
    u32 h = (u32)a * 0x9E3779B1u;
    h ^= (u32)b * 0x85EBCA77u + (h << 6) + (h >> 2);
    h ^= (u32)c * 0xC2B2AE3Du + (h << 6) + (h >> 2);
    h ^= h >> 15; h *= 0x2C1B3C6Du; h ^= h >> 12;
    return h;
}
// own[i] = the slot record i claimed (it is the representative of its stencil), 0xffffffff for a duplicate: the emit pass then
// walks the records (a few per unique stencil) instead of every slot of the table
__global__ void k_dedup_insert(const int4* __restrict__ raw, u32 n, u32* slots, u32* slotCnt, u32 mask, u32* __restrict__ own)
{
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int4 k = raw[i];
    u32 h = hash3(k.x, k.y, k.z) & mask;
    while (true) {
        const u32 prev = atomicCAS(&slots[h], 0xffffffffu, i);
        if (prev == 0xffffffffu) { atomicAdd(&slotCnt[h], 1u); own[i] = h; return; }
        const int4 o = raw[prev];
        if (o.x == k.x && o.y == k.y && o.z == k.z) { atomicAdd(&slotCnt[h], 1u); own[i] = 0xffffffffu; return; }
        h = (h + 1) & mask;
    }
}
__global__ void __launch_bounds__(256) k_dedup_emit_records(const int4* __restrict__ raw, const u32* __restrict__ own, const u32* __restrict__ slotCnt, u32 n,
    int4* out, u32* outCount)
{
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    const u32 h = i < n ? own[i] : 0xffffffffu;
    const u32 o = block_slot<1, 256>(h != 0xffffffffu ? 0 : -1, outCount);
    if (h == 0xffffffffu) return;
    const int4 k = raw[i];
    out[o] = make_int4(k.x, k.y, k.z, -(int)slotCnt[h]);
}
__global__ void k_fill_info(double2* info, u32 n, double w, double dHat2)
{
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) info[i] = make_double2(w, dHat2);
}

// ===================================================================== stencil decoding
enum Kind { K_PT = 0, K_PE = 1, K_PP = 2, K_EE = 3, K_EE_M = 4, K_PE_M = 5, K_PP_M = 6 };
struct Stencil {
    int kind, mult;
    int v[4];
};
__device__ __forceinline__ Stencil decode(const int4 c)
{
    Stencil s;
    s.mult = 1;
    if (c.x >= 0) {
        if (c.w >= 0 && c.z >= 0) { s.kind = K_EE; s.v[0] = c.x; s.v[1] = c.y; s.v[2] = c.z; s.v[3] = c.w; }
        else if (c.w >= 0) { s.kind = K_EE_M; s.v[0] = c.x; s.v[1] = c.y; s.v[2] = -c.z - 1; s.v[3] = c.w; }
        else if (c.z >= 0) { s.kind = K_PE_M; s.v[0] = c.x; s.v[1] = -c.w - 1; s.v[2] = c.y; s.v[3] = c.z; }
        else { s.kind = K_PP_M; s.v[0] = c.x; s.v[1] = -c.z - 1; s.v[2] = c.y; s.v[3] = -c.w - 1; }
    }
    else {
        s.v[0] = -c.x - 1; s.v[1] = c.y; s.v[2] = c.z; s.v[3] = c.w;
        if (c.w >= 0) s.kind = K_PT;
        else if (c.z >= 0) { s.kind = K_PE; s.mult = -c.w; }
        else { s.kind = K_PP; s.mult = -c.w; }
    }
    return s;
}
__device__ __forceinline__ double stencil_dist2(const double4* __restrict__ X, const Stencil& s)
{
    switch (s.kind) {
    case K_PT: return pt_dist2(ldx(X, s.v[0]), ldx(X, s.v[1]), ldx(X, s.v[2]), ldx(X, s.v[3]));
    case K_PE: return pe_dist2(ldx(X, s.v[0]), ldx(X, s.v[1]), ldx(X, s.v[2]));
    case K_PP: return pp_dist2(ldx(X, s.v[0]), ldx(X, s.v[1]));
    case K_EE: case K_EE_M: return ee_dist2(ldx(X, s.v[0]), ldx(X, s.v[1]), ldx(X, s.v[2]), ldx(X, s.v[3]));
    case K_PE_M: return pe_dist2(ldx(X, s.v[0]), ldx(X, s.v[2]), ldx(X, s.v[3]));
    default: return pp_dist2(ldx(X, s.v[0]), ldx(X, s.v[2]));
    }
}
__device__ __forceinline__ int block_dim_of(const int4 c) { return (c.x >= 0 || c.w >= 0) ? 12 : (c.z >= 0 ? 9 : 6); }

} // namespace cipc
#include "merge.cuh"
namespace cipc {

struct BarrierParams {
    double dHat2;      // already offset: dHat2 + 2 sqrt(dHat2) xi   (IPC.h:756-757)
    double thickness2; // xi^2
    double k0;
    int elastic;
};
// mollifier value e and threshold for a mollified stencil (rest positions X0)
__device__ __forceinline__ void mollifier_terms(const double4* __restrict__ X, const double4* __restrict__ X0, const Stencil& s,
    double& eps_x, double& cn2)
{
    eps_x = ee_mollifier_threshold(ldx(X0, s.v[0]), ldx(X0, s.v[1]), ldx(X0, s.v[2]), ldx(X0, s.v[3]));
    cn2 = ee_cross_norm2(ldx(X, s.v[0]), ldx(X, s.v[1]), ldx(X, s.v[2]), ldx(X, s.v[3]));
}

// ===================================================================== Compute_Barrier (IPC.h:742-941)
__global__ void __launch_bounds__(RED_BT) k_barrier_energy(const double4* __restrict__ X, const double4* __restrict__ X0,
    const int4* __restrict__ cs, const double2* __restrict__ info, u32 n, BarrierParams bp, double* partial, int* errFlag)
{
    double acc = 0;
    for (u32 i = blockIdx.x * RED_BT + threadIdx.x; i < n; i += RED_GRID * RED_BT) {
        const int4 c = cs[i];
        const Stencil s = decode(c);
        const double d = stencil_dist2(X, s) - bp.thickness2;
        if (d <= 0) { *errFlag = CIPC_ERR_NONPOSITIVE_DIST; continue; }
        double b = barrier_b(bp.elastic, d, bp.dHat2, bp.k0);
        if (s.kind >= K_EE_M) {
            double eps_x, cn2;
            mollifier_terms(X, X0, s, eps_x, cn2);
            if (cn2 < eps_x) b *= eem(cn2, eps_x);
        }
        else if (c.x < 0 && c.w < -1) b *= (double)(-c.w);
        acc += b * info[i].x;
    }
    acc = block_sum<RED_BT>(acc);
    if (threadIdx.x == 0) partial[blockIdx.x] = acc;
}

// ===================================================================== Compute_Barrier_Gradient (IPC.h:943-1256)
__device__ __forceinline__ void gadd(double* g, int v, const double* d3, double sc)
{
    atomicAdd(&g[3 * v], sc * d3[0]);
    atomicAdd(&g[3 * v + 1], sc * d3[1]);
    atomicAdd(&g[3 * v + 2], sc * d3[2]);
}
__global__ void __launch_bounds__(128) k_barrier_gradient(const double4* __restrict__ X, const double4* __restrict__ X0,
    const int4* __restrict__ cs, const double2* __restrict__ info, u32 n, BarrierParams bp, double* g, const u32* __restrict__ idx)
{
    const u32 k_ = blockIdx.x * blockDim.x + threadIdx.x;
    if (k_ >= n) return;
    const u32 i = idx ? idx[k_] : k_; // optional index list (mollified stencils of the fused gradient + Hessian call)
    const Stencil s = decode(cs[i]);
    const double w = info[i].x;
    const double d = stencil_dist2(X, s) - bp.thickness2;
    const double bG = barrier_g(bp.elastic, d, bp.dHat2, bp.k0);
    double dg[12];
    const int rows3[3] = {0, 1, 2};
    if (s.kind < K_EE_M) {
        int nb;
        if (s.kind == K_PT || s.kind == K_EE) {
            const dv3 x[4] = {ldd(X, s.v[0]), ldd(X, s.v[1]), ldd(X, s.v[2]), ldd(X, s.v[3])};
            d4_derivs(s.kind == K_EE, x, dg, nullptr, 0.0);
            nb = 4;
        }
        else if (s.kind == K_PE) { pe_derivs(ldd(X, s.v[0]), ldd(X, s.v[1]), ldd(X, s.v[2]), dg, nullptr, 9, rows3, 0.0); nb = 3; }
        else { pp_derivs(ldd(X, s.v[0]), ldd(X, s.v[1]), dg, nullptr, 6, rows3, 0.0); nb = 2; }
        const double sc = (double)s.mult * w * bG;
        for (int k = 0; k < nb; ++k) gadd(g, s.v[k], dg + 3 * k, sc);
        return;
    }
    const dv3 x[4] = {ldd(X, s.v[0]), ldd(X, s.v[1]), ldd(X, s.v[2]), ldd(X, s.v[3])};
    const double b = barrier_b(bp.elastic, d, bp.dHat2, bp.k0);
    double eps_x, cn2;
    mollifier_terms(X, X0, s, eps_x, cn2);
    double e = 1.0;
    if (cn2 < eps_x) {
        e = eem(cn2, eps_x);
        double eg[12];
        eecn2_derivs(x, eg, nullptr, 0.0);
        const double q = eem_g(cn2, eps_x) * w * b;
        for (int k = 0; k < 4; ++k) gadd(g, s.v[k], eg + 3 * k, q);
    }
    const double sc = e * w * bG;
    if (s.kind == K_EE_M) {
        d4_derivs(true, x, dg, nullptr, 0.0);
        for (int k = 0; k < 4; ++k) gadd(g, s.v[k], dg + 3 * k, sc);
    }
    else if (s.kind == K_PE_M) {
        pe_derivs(x[0], x[2], x[3], dg, nullptr, 9, rows3, 0.0);
        gadd(g, s.v[0], dg, sc); gadd(g, s.v[2], dg + 3, sc); gadd(g, s.v[3], dg + 6, sc);
    }
    else {
        pp_derivs(x[0], x[2], dg, nullptr, 6, rows3, 0.0);
        gadd(g, s.v[0], dg, sc); gadd(g, s.v[2], dg + 3, sc);
    }
}

// ===================================================================== Compute_Barrier_Hessian (IPC.h:1258-1731)
__global__ void k_block_sizes(const int4* __restrict__ cs, u32 n, u32* sz)
{
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int d = block_dim_of(cs[i]);
    sz[i] = (u32)(d * d / 9); // offsets are kept in units of 9 triplets (144 = 16*9, 81 = 9*9, 36 = 4*9): 32 bits reach 38G triplets
}
// Builds the n x n block of one stencil into H (row-major, n = 3*nb), returns nb and vertex ids.
__device__ void stencil_hessian(const double4* __restrict__ X, const double4* __restrict__ X0, const int4 c, double w,
    const BarrierParams& bp, bool projectSPD, double* H, int* vids, int& nb)
{
    const Stencil s = decode(c);
    const double d = stencil_dist2(X, s) - bp.thickness2;
    const double bG = barrier_g(bp.elastic, d, bp.dHat2, bp.k0);
    const double bH = barrier_H(bp.elastic, d, bp.dHat2, bp.k0);
    double dg[12];
    const int rows3[3] = {0, 1, 2};
    if (s.kind < K_EE_M) {
        const double m = (double)s.mult;
        if (s.kind == K_PT || s.kind == K_EE) {
            nb = 4;
            for (int i = 0; i < 144; ++i) H[i] = 0.0;
            const dv3 x[4] = {ldd(X, s.v[0]), ldd(X, s.v[1]), ldd(X, s.v[2]), ldd(X, s.v[3])};
            d4_derivs(s.kind == K_EE, x, dg, H, w * m * bG);
        }
        else if (s.kind == K_PE) {
            nb = 3;
            for (int i = 0; i < 81; ++i) H[i] = 0.0;
            pe_derivs(ldd(X, s.v[0]), ldd(X, s.v[1]), ldd(X, s.v[2]), dg, H, 9, rows3, w * m * bG);
        }
        else {
            nb = 2;
            for (int i = 0; i < 36; ++i) H[i] = 0.0;
            pp_derivs(ldd(X, s.v[0]), ldd(X, s.v[1]), dg, H, 6, rows3, w * m * bG);
        }
        const int n = 3 * nb;
        const double q = w * m * bH;
        for (int i = 0; i < n; ++i)
            for (int j = 0; j < n; ++j) H[i * n + j] += q * dg[i] * dg[j];
        for (int k = 0; k < nb; ++k) vids[k] = s.v[k];
        (void)projectSPD; // the caller projects (shared-memory Jacobi on the translation-reduced block)
        return;
    }
    // mollified stencils: 12 x 12 in the (ea0, ea1, eb0, eb1) frame (IPC.h:1472-1475, 1526-1537, 1588-1599)
    nb = 4;
    for (int k = 0; k < 4; ++k) vids[k] = s.v[k];
    for (int i = 0; i < 144; ++i) H[i] = 0.0;
    const dv3 x[4] = {ldd(X, s.v[0]), ldd(X, s.v[1]), ldd(X, s.v[2]), ldd(X, s.v[3])};
    const double b = barrier_b(bp.elastic, d, bp.dHat2, bp.k0);
    double eps_x, cn2;
    mollifier_terms(X, X0, s, eps_x, cn2);
    double e = 1.0;
    double eg[12];
    for (int i = 0; i < 12; ++i) eg[i] = 0.0;
    if (cn2 < eps_x) {
        e = eem(cn2, eps_x);
        const double qg = eem_g(cn2, eps_x), qH = eem_H(eps_x);
        double cg_[12];
        eecn2_derivs(x, cg_, H, w * b * qg); // b * eH, first part: q_g * d2(cn2)
        for (int i = 0; i < 12; ++i)
            for (int j = 0; j < 12; ++j) H[i * 12 + j] += w * b * qH * cg_[i] * cg_[j];
        for (int i = 0; i < 12; ++i) eg[i] = qg * cg_[i];
    }
    double G[12];
    for (int i = 0; i < 12; ++i) G[i] = 0.0;
    if (s.kind == K_EE_M) {
        d4_derivs(true, x, dg, H, w * e * bG);
        for (int i = 0; i < 12; ++i) G[i] = dg[i];
    }
    else if (s.kind == K_PE_M) {
        const int rows[3] = {0, 2, 3};
        pe_derivs(x[0], x[2], x[3], dg, H, 12, rows, w * e * bG);
        for (int k = 0; k < 3; ++k) for (int r = 0; r < 3; ++r) G[3 * rows[k] + r] = dg[3 * k + r];
    }
    else {
        const int rows[2] = {0, 2};
        pp_derivs(x[0], x[2], dg, H, 12, rows, w * e * bG);
        for (int k = 0; k < 2; ++k) for (int r = 0; r < 3; ++r) G[3 * rows[k] + r] = dg[3 * k + r];
    }
    for (int i = 0; i < 12; ++i)
        for (int j = 0; j < 12; ++j)
            H[i * 12 + j] += w * (bG * (G[i] * eg[j] + eg[i] * G[j]) + (e * bH) * G[i] * G[j]);
}
// class of a stencil for the Hessian pass: 0 = PT / EE, 1 = PE, 2 = PP (low-rank fast paths), 3 = mollified (dense path)
__global__ void __launch_bounds__(256) k_classify(const int4* __restrict__ cs, u32 n, u32* idx0, u32* idx1, u32* idx2, u32* idx3, u32* counts)
{
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    int cls = -1;
    if (i < n) {
        const int4 c = cs[i];
        if (c.x >= 0) cls = (c.w >= 0 && c.z >= 0) ? 0 : 3;
        else cls = (c.w >= 0) ? 0 : (c.z >= 0 ? 1 : 2);
    }
    const u32 slot = block_slot<4, 256>(cls, counts);
    u32* lists[4] = {idx0, idx1, idx2, idx3};
#pragma unroll
    for (int k = 0; k < 4; ++k)
        if (cls == k) lists[k][slot] = i;
}
__device__ __forceinline__ void put_triplet(cipc_triplet* o, int row, int col, double val)
{
    // one 16-byte store per triplet
    int4 q;
    q.x = row; q.y = col;
    const long long b = __double_as_longlong(val);
    q.z = (int)(b & 0xffffffffLL); q.w = (int)(b >> 32);
    *reinterpret_cast<int4*>(o) = q;
}
// CLS 0: PT / EE (rank-5 path), 1: PE (rank-4 path), 2: PP (closed form).  One thread per stencil.
template <int CLS>
__global__ void __launch_bounds__(128) k_hessian_lowrank(const double4* __restrict__ X, const int4* __restrict__ cs,
    const double2* __restrict__ info, const u32* __restrict__ off, const u32* __restrict__ idx, u32 n, BarrierParams bp,
    int projectSPD, cipc_triplet* trip, int blkMode)
{
    const u32 k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const u32 i = idx[k];
    const int4 c = cs[i];
    const Stencil s = decode(c);
    const double wm = info[i].x * (double)s.mult;
    const double d = stencil_dist2(X, s) - bp.thickness2;
    const double alpha = wm * barrier_H(bp.elastic, d, bp.dHat2, bp.k0);
    const double beta = wm * barrier_g(bp.elastic, d, bp.dHat2, bp.k0);
    cipc_triplet* o = trip + (size_t)off[i] * 9;
    constexpr int NB = (CLS == 0) ? 4 : (CLS == 1 ? 3 : 2);
    constexpr int NN = 3 * NB;
    int vid[4] = {s.v[0], s.v[1], s.v[2], s.v[3]};
    double* ob = reinterpret_cast<double*>(trip) + (size_t)off[i] * 9; // block mode: off[] counts blocks, 9 doubles each
    auto emit = [&](int I, int J, const double* B) {
        if (blkMode) { // upper block triangle, stored for (min vertex, max vertex) (merge.cuh)
            if (J < I) return;
            double* q = ob + 9 * (I * NB - I * (I - 1) / 2 + (J - I));
            const bool sw = vid[I] > vid[J];
#pragma unroll
            for (int a = 0; a < 3; ++a)
#pragma unroll
                for (int b2 = 0; b2 < 3; ++b2) q[3 * a + b2] = sw ? B[3 * b2 + a] : B[3 * a + b2];
            return;
        }
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int b2 = 0; b2 < 3; ++b2)
                put_triplet(o + (3 * I + a) * NN + 3 * J + b2, vid[I] * 3 + a, vid[J] * 3 + b2, B[3 * a + b2]);
    };
    if (CLS == 0) {
        const dv3 x[4] = {ldd(X, s.v[0]), ldd(X, s.v[1]), ldd(X, s.v[2]), ldd(X, s.v[3])};
        hess4_lowrank(s.kind == K_EE, x, alpha, beta, projectSPD != 0, emit);
    }
    else if (CLS == 1) hess_pe_lowrank(ldd(X, s.v[0]), ldd(X, s.v[1]), ldd(X, s.v[2]), alpha, beta, projectSPD != 0, emit);
    else hess_pp_closed(ldd(X, s.v[0]), ldd(X, s.v[1]), alpha, beta, projectSPD != 0, emit);
}
// ---- two-phase projected Hessian: (A) per stencil, factor the PSD block as sum_k y_k y_k^T (k <= 3 / 2 / 1),
//      (B) one thread per triplet expands y y^T into the (row, col, value) stream with fully coalesced 16-byte stores.
struct __align__(32) YHdr {
    u32 off;      // first triplet of the stencil's block (units of 9), 0xffffffff = handled by the dense path
    int v[4];     // vertex ids in block order
    int pad[3];   // pad[0] != 0: the block is -(sum_k y_k y_k^T)
};
template <int CLS> struct YShape { static constexpr int NB = (CLS == 0) ? 4 : (CLS == 1 ? 3 : 2); static constexpr int NY = (CLS == 0) ? 3 : (CLS == 1 ? 2 : 1); };

template <int CLS>
__global__ void __launch_bounds__(128, 4) k_hessian_factor(const double4* __restrict__ X, const int4* __restrict__ cs,
    const double2* __restrict__ info, const u32* __restrict__ off, const u32* __restrict__ idx, u32 n, BarrierParams bp,
    double* __restrict__ Yout, YHdr* __restrict__ hdr, u32* denseList, u32* denseCount)
{
    const u32 q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n) return;
    const u32 i = idx[q];
    const Stencil s = decode(cs[i]);
    const double wm = info[i].x * (double)s.mult;
    const double d = stencil_dist2(X, s) - bp.thickness2;
    const double alpha = wm * barrier_H(bp.elastic, d, bp.dHat2, bp.k0);
    const double beta = wm * barrier_g(bp.elastic, d, bp.dHat2, bp.k0);
    constexpr int NN = 3 * YShape<CLS>::NB, NY = YShape<CLS>::NY;
    YHdr h;
    h.off = off[i];
    h.v[0] = s.v[0]; h.v[1] = s.v[1]; h.v[2] = s.v[2]; h.v[3] = s.v[3];
    h.pad[0] = h.pad[1] = h.pad[2] = 0;
    if (!(beta < 0.0) || !(alpha > 0.0)) {
        // outside the barrier's support (stale constraint set) the inertia argument does not hold: dense path
        h.off = 0xffffffffu;
        hdr[q] = h;
        denseList[atomicAdd(denseCount, 1u)] = i;
        return;
    }
    double Y[NY * NN];
    if (CLS == 0) {
        const dv3 x[4] = {ldd(X, s.v[0]), ldd(X, s.v[1]), ldd(X, s.v[2]), ldd(X, s.v[3])};
        hess4_factor(s.kind == K_EE, x, alpha, beta, Y);
    }
    else if (CLS == 1) hess_pe_factor(ldd(X, s.v[0]), ldd(X, s.v[1]), ldd(X, s.v[2]), alpha, beta, Y);
    else hess_pp_factor(ldd(X, s.v[0]), ldd(X, s.v[1]), alpha, beta, Y);
    double2* o = reinterpret_cast<double2*>(Yout + (size_t)q * (NY * NN));
#pragma unroll
    for (int k = 0; k < NY * NN / 2; ++k) o[k] = make_double2(Y[2 * k], Y[2 * k + 1]);
    hdr[q] = h;
}
template <int NB, int NY>
__global__ void __launch_bounds__(256) k_hessian_expand(const double* __restrict__ Yin, const YHdr* __restrict__ hdr, u64 total,
    cipc_triplet* __restrict__ trip)
{
    constexpr int NN = 3 * NB, PER = NN * NN;
    // one 64-bit division per thread for the block base, everything else in small 32-bit arithmetic
    const u64 base = (u64)blockIdx.x * 256u;
    if (base + threadIdx.x >= total) return;
    const u32 q0 = (u32)(base / PER);
    const u32 x = (u32)(base - (u64)q0 * PER) + threadIdx.x; // < PER + 256
    const u32 dq = x / PER;
    const u32 q = q0 + dq;
    const int e = (int)(x - dq * PER);
    const int r = e / NN, c = e - r * NN;
    const YHdr h = hdr[q];
    if (h.off == 0xffffffffu) return;
    const double* y = Yin + (size_t)q * (NY * NN);
    double v = 0.0;
#pragma unroll
    for (int k = 0; k < NY; ++k) v += y[k * NN + r] * y[k * NN + c];
    if (h.pad[0]) v = -v;
    const int ri = r / 3, ci = c / 3;
    put_triplet(trip + (size_t)h.off * 9 + e, h.v[ri] * 3 + (r - 3 * ri), h.v[ci] * 3 + (c - 3 * ci), v);
}
// Tiled expansion: a CTA stages the factors and headers of G consecutive stencils in shared memory with coalesced
// 16-byte loads, then every thread forms G*PER/256 triplets from shared memory and issues that many independent
// 16-byte stores back to back (a warp writes 512 contiguous bytes per instruction).  Versus one thread per triplet
// this removes the load -> store dependency from the critical path: the write stream is limited by HBM, not by the
// latency of the factor gathers.  STREAM selects st.global.cs (the triplet stream is never re-read by this library).
template <int NB, int NY, int G, bool STREAM>
__global__ void __launch_bounds__(256) k_hessian_expand_tiled(const double* __restrict__ Yin, const YHdr* __restrict__ hdr, u32 n,
    cipc_triplet* __restrict__ trip)
{
    constexpr int NN = 3 * NB, PER = NN * NN, YD = NY * NN;
    static_assert((G * YD) % 2 == 0, "factor tile must be a whole number of 16-byte words");
    __shared__ __align__(16) double sY[G * YD];
    __shared__ __align__(16) int sH[G * 8];
    const u32 q0 = blockIdx.x * G;
    const u32 g = min((u32)G, n - q0);
    {
        const double2* src = reinterpret_cast<const double2*>(Yin + (size_t)q0 * YD);
        double2* dst = reinterpret_cast<double2*>(sY);
        for (u32 t = threadIdx.x; t < g * YD / 2; t += 256) dst[t] = __ldcs(src + t);
        const int4* hs = reinterpret_cast<const int4*>(hdr + q0);
        int4* hd = reinterpret_cast<int4*>(sH);
        for (u32 t = threadIdx.x; t < g * 2; t += 256) hd[t] = __ldcs(hs + t);
    }
    __syncthreads();
    const u32 total = g * PER;
#pragma unroll 4
    for (u32 t = threadIdx.x; t < total; t += 256) {
        const u32 q = t / PER;
        const int e = (int)(t - q * PER);
        const int r = e / NN, c = e - r * NN;
        const int* h = sH + q * 8;
        const u32 off = (u32)h[0];
        if (off == 0xffffffffu) continue;
        const double* y = sY + q * YD;
        double v = 0.0;
#pragma unroll
        for (int k = 0; k < NY; ++k) v += y[k * NN + r] * y[k * NN + c];
        if (h[5]) v = -v; // YHdr::pad[0]: the block is -(sum y y^T) (friction with a negative lagged normal force)
        const int ri = r / 3, ci = c / 3;
        int4 o;
        o.x = h[1 + ri] * 3 + (r - 3 * ri); o.y = h[1 + ci] * 3 + (c - 3 * ci);
        const long long b = __double_as_longlong(v);
        o.z = (int)(b & 0xffffffffLL); o.w = (int)(b >> 32);
        int4* dst = reinterpret_cast<int4*>(trip + (size_t)off * 9 + e);
        if (STREAM) __stcs(dst, o); else *dst = o;
    }
}
// Expansion, warp by warp: a warp expands the (up to) 32 stencils whose factors its own lanes parked in shared memory
// (stencil q of the CTA: factors at sY + q*YS, header {offset, v0..v3, sign} at sH + q*8).  No CTA barrier is involved:
// warps stay decoupled, so one warp's store stream overlaps the others' arithmetic.  A lane owns one column of the block and
// a group of its rows (see below).
// sR: per stencil, the NN global row indices 3 v[r/3] + r%3 (written by the factoring thread).  SIGNED: honour the header's
// sign flag (friction only; the barrier factors are never negated).
// ROWTAB = false: no table, the indices are formed from the header's vertex ids per triplet (friction: a pure store stream
// that prefers the extra resident CTA the table's 6 KB would cost).
template <int NN, int NY, int YS, bool SIGNED, bool ROWTAB = true>
__device__ __forceinline__ void warp_expand_stencils(const double* sY, const int* sH, const int* sR, u32 wq0, u32 g, u32 lane,
    cipc_triplet* __restrict__ trip)
{
    // Lane = (column c, row group): NG groups of NN lanes, a group owns RPG consecutive rows.  The lane's column of the
    // factors and its column index stay in registers for the whole stencil, so a triplet costs NY shared-memory loads (warp
    // broadcasts of y[k][r]) + 1 index load instead of 2 NY + 2; a store instruction writes NG runs of NN x 16 contiguous bytes.
    constexpr int NG = (NN == 12) ? 2 : 3, RPG = NN / NG;
    static_assert(NG * NN <= 32 && NG * RPG == NN, "lane layout");
    const int c = (int)lane % NN, grp = (int)lane / NN;
    const bool active = grp < NG;
    const int r0 = grp * RPG;
    const u32 wn = (wq0 < g) ? min(32u, g - wq0) : 0u;
    if (!active) return; // no warp-level primitives below
    for (u32 qq = 0; qq < wn; ++qq) {
        const int* h = sH + (wq0 + qq) * 8;
        const u32 o = (u32)h[0];
        if (o == 0xffffffffu) continue;
        const double* y = sY + (wq0 + qq) * YS;
        const int* rows = sR + (wq0 + qq) * NN;
        const bool neg = SIGNED && h[5] != 0;
        double yc[NY];
#pragma unroll
        for (int k = 0; k < NY; ++k) { yc[k] = y[k * NN + c]; if (SIGNED && neg) yc[k] = -yc[k]; }
        const int colIdx = ROWTAB ? rows[c] : h[1 + c / 3] * 3 + c % 3;
        cipc_triplet* dst = trip + (size_t)o * 9 + r0 * NN + c;
#pragma unroll
        for (int j = 0; j < RPG; ++j) {
            const int r = r0 + j;
            double v = 0.0;
#pragma unroll
            for (int k = 0; k < NY; ++k) v += y[k * NN + r] * yc[k];
            int4 w;
            w.x = ROWTAB ? rows[r] : h[1 + r / 3] * 3 + r % 3;
            w.y = colIdx;
            const long long bb = __double_as_longlong(v);
            w.z = (int)(bb & 0xffffffffLL); w.w = (int)(bb >> 32);
            __stcs(reinterpret_cast<int4*>(dst + j * NN), w);
        }
    }
}
// Fused factor + expansion for the device-resident triplet stream: a CTA of 128 threads factors 128 stencils (FP64
// pipe, one stencil per thread), parks the factors in shared memory, and then all threads turn them into triplets with
// coalesced streaming 16-byte stores (HBM).  CTAs resident on one SM are in different phases, so the FP64 work of one
// overlaps the store stream of another, and the factors never travel through HBM (-3.3 GB per launch at 1M triangles).
// The factor-only kernel above stays in use when the triplets are delivered to the HOST (compact factors cross PCIe).
constexpr int FUSED_BD = 128;
#ifndef FUSED_MINB
#define FUSED_MINB 4
#endif
#ifndef FUSED_MINB12
#define FUSED_MINB12 6 // PE / PP kernels: small factors, bound by the latency of the position gathers -> more resident warps
#endif
#ifndef FUSED_STASH
#define FUSED_STASH 1
#endif
template <int CLS> struct FusedShape {
    static constexpr int NN = 3 * YShape<CLS>::NB, NY = YShape<CLS>::NY, YD = NY * NN;
    static constexpr int YS = (YD % 2 == 0) ? YD + 1 : YD; // odd stride: conflict-free 8-byte accesses, one stencil per lane
    static constexpr int SMEM = FUSED_BD * YS * 8 + FUSED_BD * 32 + FUSED_BD * NN * 4; // factors, headers, row indices
};
// BLK: the stencil's upper 3x3 blocks (9 doubles each, merge.cuh) are written instead of 16-byte triplets -- 720 instead of
// 2304 bytes per PT/EE stencil -- and `off` counts blocks.
template <int CLS, bool BLK>
__global__ void __launch_bounds__(FUSED_BD, (CLS == 0 ? FUSED_MINB : FUSED_MINB12)) k_hessian_fused(const double4* __restrict__ X, const int4* __restrict__ cs,
    const double2* __restrict__ info, const u32* __restrict__ off, const u32* __restrict__ idx, u32 n, BarrierParams bp,
    void* __restrict__ outp, u32* denseList, u32* denseCount, double* gOut)
{
    constexpr int NN = FusedShape<CLS>::NN, NY = FusedShape<CLS>::NY, YS = FusedShape<CLS>::YS;
    extern __shared__ __align__(16) unsigned char fused_sm[];
    double* sY = reinterpret_cast<double*>(fused_sm);
    int* sH = reinterpret_cast<int*>(fused_sm + FUSED_BD * YS * 8);
    int* sR = sH + FUSED_BD * 8;
    const u32 q0 = blockIdx.x * FUSED_BD;
    const u32 g = min((u32)FUSED_BD, n - q0);
    if (threadIdx.x < g) {
        const u32 i = idx[q0 + threadIdx.x];
        const Stencil s = decode(cs[i]);
        const double wm = info[i].x * (double)s.mult;
        const double d = stencil_dist2(X, s) - bp.thickness2;
        const double alpha = wm * barrier_H(bp.elastic, d, bp.dHat2, bp.k0);
        const double beta = wm * barrier_g(bp.elastic, d, bp.dHat2, bp.k0);
        int* h = sH + threadIdx.x * 8;
        h[1] = s.v[0]; h[2] = s.v[1]; h[3] = s.v[2]; h[4] = s.v[3]; h[5] = 0;
        if (BLK) h[6] = swap_mask<YShape<CLS>::NB>(s.v);
        else {
#pragma unroll
            for (int r = 0; r < NN; ++r) sR[threadIdx.x * NN + r] = 3 * s.v[r / 3] + r % 3;
        }
        if (gOut) { // barrier gradient of the same stencil (cipc_barrier_gradient_hessian_dev): w m b' grad d, as k_barrier_gradient
            double dg[12];
            const int rows3[3] = {0, 1, 2};
            if (CLS == 0) {
                const dv3 x[4] = {ldd(X, s.v[0]), ldd(X, s.v[1]), ldd(X, s.v[2]), ldd(X, s.v[3])};
                d4_derivs(s.kind == K_EE, x, dg, nullptr, 0.0);
            }
            else if (CLS == 1) pe_derivs(ldd(X, s.v[0]), ldd(X, s.v[1]), ldd(X, s.v[2]), dg, nullptr, 9, rows3, 0.0);
            else pp_derivs(ldd(X, s.v[0]), ldd(X, s.v[1]), dg, nullptr, 6, rows3, 0.0);
            constexpr int NBL = YShape<CLS>::NB;
#pragma unroll
            for (int k = 0; k < NBL; ++k) gadd(gOut, s.v[k], dg + 3 * k, beta);
        }
        if (!(beta < 0.0) || !(alpha > 0.0)) { // outside the barrier's support: dense path (see k_hessian_factor)
            h[0] = (int)0xffffffffu;
            denseList[atomicAdd(denseCount, 1u)] = i;
        }
        else {
            h[0] = (int)off[i];
            double* Y = sY + threadIdx.x * YS; // the factor routines write every entry exactly once: straight into the thread's slot
            if (CLS == 0) {
                const dv3 x[4] = {ldd(X, s.v[0]), ldd(X, s.v[1]), ldd(X, s.v[2]), ldd(X, s.v[3])};
                hess4_factor<FUSED_STASH != 0>(s.kind == K_EE, x, alpha, beta, Y); // the slot doubles as the parking space of the iteration
            }
            else if (CLS == 1) hess_pe_factor(ldd(X, s.v[0]), ldd(X, s.v[1]), ldd(X, s.v[2]), alpha, beta, Y);
            else hess_pp_factor(ldd(X, s.v[0]), ldd(X, s.v[1]), alpha, beta, Y);
        }
    }
    __syncwarp();
    if (BLK) warp_expand_blocks<YShape<CLS>::NB, NY, YS, false>(sY, sH, threadIdx.x & ~31u, g, threadIdx.x & 31u, reinterpret_cast<double*>(outp));
    else warp_expand_stencils<NN, NY, YS, false>(sY, sH, sR, threadIdx.x & ~31u, g, threadIdx.x & 31u, reinterpret_cast<cipc_triplet*>(outp));
}
// dense path (mollified stencils; also usable for every stencil as a cross-check: idx == nullptr)
constexpr int DENSE_BD = 64;                              // threads per block of the dense path
constexpr int DENSE_SMEM = 81 * DENSE_BD * 8;              // the 9x9 matrix the eigen-solver works on, per thread
#ifndef DENSE_MINB
#define DENSE_MINB 4
#endif
__global__ void __launch_bounds__(DENSE_BD, DENSE_MINB) k_barrier_hessian(const double4* __restrict__ X, const double4* __restrict__ X0,
    const int4* __restrict__ cs, const double2* __restrict__ info, const u32* __restrict__ off, const u32* __restrict__ idx, u32 n,
    const u32* __restrict__ nDev, BarrierParams bp, int projectSPD, cipc_triplet* trip, int blkMode, u32 first = 0)
{
    extern __shared__ double dense_sm[];
    if (nDev) n = *nDev; // list length produced on the device (fallbacks of the factor kernels)
    for (u32 k = first + blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        const u32 i = idx ? idx[k] : k;
        double H[144];
        int vids[4], nb;
        stencil_hessian(X, X0, cs[i], info[i].x, bp, false, H, vids, nb);
        if (projectSPD) psd_project_reduced_smem(H, nb, dense_sm, threadIdx.x, DENSE_BD);
        const int nn = 3 * nb;
        if (blkMode) { store_blocks_dense(H, nb, vids, reinterpret_cast<double*>(trip) + (size_t)off[i] * 9); continue; }
        cipc_triplet* o = trip + (size_t)off[i] * 9;
        for (int I = 0; I < nb; ++I)
            for (int a = 0; a < 3; ++a)
                for (int J = 0; J < nb; ++J)
                    for (int b2 = 0; b2 < 3; ++b2)
                        put_triplet(o + (I * 3 + a) * nn + J * 3 + b2, vids[I] * 3 + a, vids[J] * 3 + b2, H[(I * 3 + a) * nn + J * 3 + b2]);
    }
}
// dense-path blocks (mollified / rejected stencils) packed for the host: block k at dst + 144 k, meta[k] = (offset, count)
__global__ void k_gather_dense(const int4* __restrict__ cs, const u32* __restrict__ off, const u32* __restrict__ list, const u32* __restrict__ nDev,
    const cipc_triplet* __restrict__ trip, cipc_triplet* __restrict__ dst, uint2* __restrict__ meta)
{
    const u32 n = *nDev;
    for (u32 k = blockIdx.x; k < n; k += gridDim.x) {
        const u32 i = list[k];
        const int d = block_dim_of(cs[i]);
        const u32 cnt = (u32)(d * d), o = off[i];
        const size_t o9 = (size_t)o * 9;
        if (threadIdx.x == 0) meta[k] = make_uint2(o, cnt);
        const int4* src = reinterpret_cast<const int4*>(trip + o9);
        int4* out = reinterpret_cast<int4*>(dst + (size_t)k * 144);
        for (u32 t = threadIdx.x; t < cnt; t += blockDim.x) out[t] = src[t];
    }
}
template <int N>
__global__ void k_test_make_pd(double* H, int count)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    double A[N * N];
    for (int k = 0; k < N * N; ++k) A[k] = H[(size_t)i * N * N + k];
    psd_project_jacobi<N>(A);
    for (int k = 0; k < N * N; ++k) H[(size_t)i * N * N + k] = A[k];
}

// ===================================================================== Compute_Min_Dist2 (IPC.h:2246-2388)
__global__ void k_min_dist(const double4* __restrict__ X, const int4* __restrict__ cs, u32 n, double* dist2, long long* minBits)
{
    double m = 1e300;
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const double d = stencil_dist2(X, decode(cs[i]));
        if (dist2) dist2[i] = d;
        m = fmin(m, d);
    }
    for (int o = 16; o > 0; o >>= 1) m = fmin(m, __shfl_down_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) atomicMin(minBits, dbl_ordered(m));
}

// ===================================================================== ACCD over candidate pairs
// alphaBits holds the running minimum step as the bit pattern of a positive double (monotone).
struct AccdOut {
    u64* alphaBits;
    int* errFlag;
};
__device__ __forceinline__ void accd_commit(const AccdOut& o, bool hit, xd toc, bool checkZero)
{
    if (hit) atomicMin(o.alphaBits, (u64)__double_as_longlong(toc.v));
    if (checkZero && hit && toc.v == 0.0) *o.errFlag = CIPC_ERR_ZERO_STEP;
}
__global__ void __launch_bounds__(128) k_accd_pt(Topo T, const double4* __restrict__ X, const double4* __restrict__ P,
    const int2* __restrict__ cand, u32 n, double thickness, double bound, AccdOut out)
{
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int2 c = cand[i];
    const int vI = T.BN[c.x];
    const int4 t = T.BT[c.y];
    xd toc;
    const bool hit = pt_accd(ldx(X, vI), ldx(X, t.x), ldx(X, t.y), ldx(X, t.z), ldx(P, vI), ldx(P, t.x), ldx(P, t.y), ldx(P, t.z),
        xd(0.1), xd(thickness), xd(bound), toc, (const volatile double*)out.alphaBits);
    accd_commit(out, hit, toc, true);
}
__global__ void __launch_bounds__(128) k_accd_ee(Topo T, const double4* __restrict__ X, const double4* __restrict__ P,
    const int2* __restrict__ cand, u32 n, double thickness, double bound, AccdOut out)
{
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int2 c = cand[i];
    const int2 a = T.BE[c.x], b = T.BE[c.y];
    xd toc;
    const bool hit = ee_accd(ldx(X, a.x), ldx(X, a.y), ldx(X, b.x), ldx(X, b.y), ldx(P, a.x), ldx(P, a.y), ldx(P, b.x), ldx(P, b.y),
        xd(0.1), xd(thickness), xd(bound), toc, (const volatile double*)out.alphaBits);
    accd_commit(out, hit, toc, false);
}
__global__ void __launch_bounds__(128) k_accd_pe(Topo T, const double4* __restrict__ X, const double4* __restrict__ P,
    const int2* __restrict__ cand, u32 n, double thickness, double bound, AccdOut out)
{
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int2 c = cand[i];
    const int vI = T.BN[c.x];
    const int2 e = T.BE[c.y];
    xd toc;
    const bool hit = pe_accd(ldx(X, vI), ldx(X, e.x), ldx(X, e.y), ldx(P, vI), ldx(P, e.x), ldx(P, e.y), xd(0.1), xd(thickness),
        xd(bound), toc, (const volatile double*)out.alphaBits);
    accd_commit(out, hit, toc, false);
}
__global__ void __launch_bounds__(128) k_accd_pp(Topo T, const double4* __restrict__ X, const double4* __restrict__ P,
    const int2* __restrict__ cand, u32 n, double thickness, double bound, AccdOut out)
{
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int2 c = cand[i];
    const int vI = T.BN[c.x], vJ = T.BN[c.y];
    xd toc;
    const bool hit = pp_accd(ldx(X, vI), ldx(X, vJ), ldx(P, vI), ldx(P, vJ), xd(0.1), xd(thickness), xd(bound), toc,
        (const volatile double*)out.alphaBits);
    accd_commit(out, hit, toc, false);
}

// ===================================================================== marshalling helpers
__global__ void k_expand3(const double* __restrict__ src, double4* dst, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = make_double4(src[3 * i], src[3 * i + 1], src[3 * i + 2], 0.0);
}
__global__ void k_pack_edges(const int* __restrict__ src, int stride, int n, int2* dst)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = make_int2(src[(size_t)i * stride], src[(size_t)i * stride + 1]);
}
__global__ void k_pack_tris(const int* __restrict__ src, int stride, int n, int4* dst)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = make_int4(src[(size_t)i * stride], src[(size_t)i * stride + 1], src[(size_t)i * stride + 2], 0);
}
// line-search trial positions x = xprev + alpha p (Shell/IMPLICIT_EULER.h:102-109), rounded like the unfused host expression
__global__ void k_step_positions(const double4* __restrict__ Xprev, const double4* __restrict__ P, double alpha, int n, double4* __restrict__ X)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double4 a = Xprev[i], p = P[i];
    X[i] = make_double4(__dadd_rn(a.x, __dmul_rn(alpha, p.x)), __dadd_rn(a.y, __dmul_rn(alpha, p.y)), __dadd_rn(a.z, __dmul_rn(alpha, p.z)), 0.0);
}
__global__ void k_fill_u32(u32* p, size_t n, u32 v)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

} // namespace cipc

#include "friction.cuh"
#include "csr.cuh"
#include "boundary.cuh"

struct cipc_ctx;
#include "multidev.h"

// ========================================================================================= context
using namespace cipc;

struct StageEv {
    std::string name;
    cudaEvent_t a, b;
};

struct cipc_ctx {
    int dev = 0, rank = 0, world = 1;
    std::unique_ptr<cipc_multi> multi; // cipc_create_multi: this context only fans calls out to multi->sub (multidev.h)
    u32 nPassLast = 0;                 // stencils of the last constraint set that needed no de-duplication (PT / EE / mollified)
    cudaStream_t st = nullptr, ownSt = nullptr;
    cipc::DevBuf<cipc::u32> dedupOwn;              // per raw PP / PE record: the hash slot it claimed, or 0xffffffff (duplicate)
    cipc::u32 dedupLastRaw = 0, dedupLastUnique = 0; // duplicate ratio of the previous constraint set (table sizing only)
    cudaStream_t sideSt = nullptr;                 // dense-path launches that overlap the fused Hessian kernels (fork / join by events)
    cudaEvent_t evFork = nullptr, evJoin = nullptr;
    cudaEvent_t userEv[64] = {};
    std::string err;
    // topology
    Topo T{};
    u64 topoHash = 0;
    DevBuf<int> BN, v2sv, stageI;
    DevBuf<int2> BE;
    DevBuf<int4> BT;
    DevBuf<uint8_t> flags;
    DevBuf<u64> nnx;
    DevBuf<double> BNArea, BEArea, BTArea;
    // state
    DevBuf<double4> X, X0, P, Xprev;
    bool haveX = false, haveX0 = false, haveP = false, haveXprev = false;
    u64 tagX = 0, tagX0 = 0, tagP = 0, tagXn = 0; // content tags of the resident uploads (upload_vec3); 0 = unknown
    struct SlabCache { bool valid = false; int age = 0, gx = 0, gy = 0, gz = 0, nP = 0; Slab sl{0, 0, 1 << 30}; };
    SlabCache slabCache[2]; // [0] constraint-set grid, [1] swept step-size grid (multi-GPU slab bounds, build_cell_lists)
    u64 xVersion = 1, edgeLenVersion = 0; // bumped whenever the resident positions or the topology change
    double edgeLenCached = 0.0;
    bool edgeLenPending = false;
    DevBuf<double> stageD, restLen2;
    // hash
    DevBuf<u64> boxLo, boxHi, nodeLo, nodeHi;
    DevBuf<ulonglong2> fine, nodeFine;
    DevBuf<u32> slabHist;
    DevBuf<u32> cnt, keys, vals, heads, headScan, ks;
    DevBuf<uint2> codes; // per sorted entry: cell-local pair code (k_entry_codes)
    DevBuf<u32> taskCnt, nTasksDev, binNext;
    DevBuf<uint2> taskDesc; // pair-enumeration tasks: (cell, task index within the cell)
    SortWork sortwk;
    ScanWork scanwk;
    DevBuf<double> partial, scal;  // scal: [0] E partial, [1] alpha, [2] min dist2, [3..] scratch
    DevBuf<long long> bbox;        // 6 ordered ints
    DevBuf<u32> counters;          // 16 device counters
    DevBuf<int> errFlag;
    // candidates
    DevBuf<int2> cand[4];
    // constraints
    DevBuf<int4> cs, raw;
    DevBuf<double2> info;
    u32 nC = 0;
    DevBuf<u32> slots, slotCnt;
    // outputs
    DevBuf<double> g, dist2;
    DevBuf<cipc_triplet> trip;
    DevBuf<u32> tripOff, clsIdx[4];
    DevBuf<double> Y;
    bool factorValid = false, expanded = true;
    double *Y0 = nullptr, *Y1 = nullptr, *Y2 = nullptr;
    void *h0 = nullptr, *h1 = nullptr, *h2 = nullptr;
    u32 nk[4] = {0, 0, 0, 0};
    int ny[3] = {3, 2, 1};  // factor vectors per stencil of class 0/1/2: barrier {3,2,1}, friction {2,2,2}
    DevBuf<cipc_triplet> denseBuf;
    DevBuf<uint2> denseMeta;
    PinnedBuf pinY, pinH, pinD, pinM, pinSmall, pinScal;
    DevBuf<double4> yhdr; // YHdr records (32 B each)
    int64_t nTrip = 0;
    PinnedBuf pin;
    // triplets -> CSR assembly (csr.cuh)
    const int4* hessStencils = nullptr; // stencil array (cs or fcs) and count of the Hessian whose triplets are resident
    u32 hessN = 0;
    DevBuf<double> blkVal, csrVal;
    DevBuf<u32> keyI, keyJ, csrKey, csrIds, csrColS, csrHeads, csrScan, urow, ucol, ustart, browCnt;
    DevBuf<int> csrRowPtr, csrColIdx;
    size_t nBlk = 0;
    u32 nU = 0;
    bool csrValid = false;
    // block-mode Hessian + device-side merge (merge.cuh): unique upper blocks (mRow, mCol, mVal) of the Hessian computed last
    DevBuf<u64> ent;
    DevBuf<u32> rowStart, rowCur, mHeads, mScan, mRow, mCol, mStart;
    DevBuf<double> mVal;
    u32 nBlkLast = 0, nUm = 0, nDiag = 0;
    bool mergedValid = false;
    PinnedBuf pinMK, pinMV;
    // boundary-primitive construction (boundary.cuh): device scratch + the assembled host lists of the last build
    DevBuf<int4> bdTri;
    DevBuf<double> bdThird, bdBTArea, bdUeVal, bdBEArea, bdNodeSum, bdBNArea;
    DevBuf<u32> bdLo, bdHi, bdV, bdIds, bdKey, bdHeads, bdScan, bdUeA, bdUeB, bdIds2, bdNodeV, bdKeep, bdKeepScan;
    DevBuf<int2> bdBE;
    DevBuf<int> bdBN;
    std::vector<int> hBN;
    std::vector<int2> hBE;
    std::vector<double> hBNArea, hBEArea, hBTArea;
    std::vector<int> hTri; // 3 per triangle
    u64 bdHash = 0;
    int bdCodim[2] = {0, 0};
    StageRing ring;              // staging for transfers from / to pageable host memory (hostpool.h)
    bool infoUniform = false;    // stencilInfo of the resident set is (infoW, infoD) for every constraint (non-elastic sets)
    double infoW = 1.0, infoD = 0.0;
    // lagged friction (FEM/FRICTION.h): the friction set lives on the device between calls
    DevBuf<double4> Xn;
    bool haveXn = false;
    DevBuf<int4> fcs;
    DevBuf<double2> fcp;
    DevBuf<double> fB, fnf, muComp;
    DevBuf<int> compRange;
    DevBuf<u32> fslot;
    u32 nF = 0;
    // timing
    std::vector<StageEv> stages;
    std::vector<cudaEvent_t> evPool;
    size_t evUsed = 0;
    std::map<std::string, int64_t> ctr;

    cudaEvent_t ev()
    {
        if (evUsed == evPool.size()) {
            cudaEvent_t e;
            CIPC_CUDA(cudaEventCreate(&e));
            evPool.push_back(e);
        }
        return evPool[evUsed++];
    }
    void begin_call() { stages.clear(); evUsed = 0; }
    bool timing = true; // cipc_set_timing: stage scopes record CUDA events only while on
    struct Scope {
        cipc_ctx* c;
        size_t idx;
        Scope(cipc_ctx* c_, const char* name) : c(c_), idx((size_t)-1)
        {
            if (!c->timing) return;
            StageEv s;
            s.name = name; s.a = c->ev(); s.b = c->ev();
            CIPC_CUDA(cudaEventRecord(s.a, c->st));
            c->stages.push_back(s);
            idx = c->stages.size() - 1;
        }
        ~Scope() { if (idx != (size_t)-1) cudaEventRecord(c->stages[idx].b, c->st); }
    };
};

namespace {

const int TB = 256;

template <class F>
int guarded(cipc_ctx* ctx, F f)
{
    if (!ctx) return CIPC_ERR_ARG;
    try {
        CIPC_CUDA(cudaSetDevice(ctx->dev));
        return f();
    }
    catch (const CudaError& e) {
        ctx->err = e.what();
        return CIPC_ERR_CUDA;
    }
    catch (const std::exception& e) {
        ctx->err = e.what();
        return CIPC_ERR_ARG;
    }
}

// change detector for host arrays (not a cryptographic hash): a Fletcher-style pair of running sums on four 64-bit lanes
// over 32-byte strides (a += w; b += a -- position dependent, two vector adds per 32 bytes: memory-bandwidth bound),
// folded through a multiply-xorshift mix
__attribute__((target("avx2"))) u64 fnv(const void* p, size_t n, u64 h)
{
    const unsigned char* c = (const unsigned char*)p;
    __m256i a0 = _mm256_set1_epi64x((long long)(h ^ 0x9E3779B97F4A7C15ULL)), b0 = _mm256_setzero_si256();
    __m256i a1 = _mm256_set1_epi64x((long long)(h + 0xC2B2AE3D27D4EB4FULL)), b1 = _mm256_setzero_si256();
    size_t i = 0;
    for (; i + 64 <= n; i += 64) {
        a0 = _mm256_add_epi64(a0, _mm256_loadu_si256((const __m256i*)(c + i))); b0 = _mm256_add_epi64(b0, a0);
        a1 = _mm256_add_epi64(a1, _mm256_loadu_si256((const __m256i*)(c + i + 32))); b1 = _mm256_add_epi64(b1, a1);
    }
    u64 l[16];
    _mm256_storeu_si256((__m256i*)l, a0); _mm256_storeu_si256((__m256i*)(l + 4), b0);
    _mm256_storeu_si256((__m256i*)(l + 8), a1); _mm256_storeu_si256((__m256i*)(l + 12), b1);
    for (int k = 0; k < 16; ++k) { h = (h ^ l[k]) * 0x9FB21C651E98DF25ULL; h ^= h >> 29; }
    for (; i < n; ++i) h = (h ^ c[i]) * 0x100000001b3ULL;
    return h ^ (h >> 32) ^ (u64)n;
}

// change detector for large host arrays (topology lists, the caller's constraint set): hashed in 1 MiB pieces by a few
// host threads, the piece hashes chained in order -- ~1 ms per 50 MB instead of ~5 ms single-threaded
struct HashSeg { const unsigned char* p; size_t n; };
u64 hash_segments(const HashSeg* segs, int nSeg, u64 h)
{
    const size_t PIECE = (size_t)1 << 20;
    std::vector<HashSeg> pieces;
    for (int k = 0; k < nSeg; ++k)
        for (size_t o = 0; o < segs[k].n; o += PIECE) pieces.push_back({segs[k].p + o, std::min(PIECE, segs[k].n - o)});
    std::vector<u64> ph(pieces.size());
    if (pieces.size() < 8) for (size_t k = 0; k < pieces.size(); ++k) ph[k] = fnv(pieces[k].p, pieces[k].n, 0x9E3779B97F4A7C15ULL + k);
    else HostPool::get().for_each(pieces.size(), [&](size_t k) { ph[k] = fnv(pieces[k].p, pieces[k].n, 0x9E3779B97F4A7C15ULL + k); });
    return fnv(ph.data(), ph.size() * sizeof(u64), h);
}

// Uploads nV x (x, y, z) records.  `tag` remembers (content hash, stride) of what is resident in `dst`: the contact stage
// passes the same X to up to eight consecutive calls (Shell/IMPLICIT_EULER.h:418-600), hashing 16 MB costs a fifth of
// moving it.  Returns true when the device copy changed.
bool upload_vec3(cipc_ctx* c, DevBuf<double4>& dst, const double* src, int stride_bytes, u64* tag)
{
    const int n = c->T.nV;
    if (stride_bytes != 32 && stride_bytes != 24) throw std::runtime_error("stride_bytes must be 24 or 32");
    u64 hsh = 0;
    if (tag) {
        const HashSeg sg{(const unsigned char*)src, (size_t)n * stride_bytes};
        hsh = hash_segments(&sg, 1, 0x51ED270B0000ULL + (u64)stride_bytes) | 1ULL;
        if (hsh == *tag && dst.cap >= (size_t)n) return false;
    }
    dst.reserve(n, c->st);
    if (stride_bytes == 32) {
        staged_h2d(c->ring, c->st, dst.p, src, (size_t)n * 32);
    }
    else {
        c->stageD.reserve((size_t)3 * n, c->st);
        staged_h2d(c->ring, c->st, c->stageD.p, src, (size_t)n * 24);
        CIPC_LAUNCH(k_expand3, div_up(n, TB), TB, 0, c->st, c->stageD.p, dst.p, n);
    }
    if (tag) *tag = hsh;
    return true;
}

// ---- spatial hash (shared by the constraint-set and the step-size passes)
struct HashInfo {
    GridDesc G;
    u32 nEntries = 0, nCells = 0, maxTasks = 0, qch = 16;
};

// mean boundary-edge length of the resident positions; cached until the positions (or the topology) change, so the
// step-size pass that follows a constraint-set pass on the same X saves the reduction and its host round trip.
// edge_len_begin launches the reduction and its copy without waiting; the value is valid after the next stream sync.
void edge_len_begin(cipc_ctx* c)
{
    if (c->edgeLenVersion == c->xVersion || !c->T.nBE) return;
    double* h = (double*)c->pinScal.reserve(64);
    CIPC_LAUNCH(k_edge_len_partial, RED_GRID, RED_BT, 0, c->st, c->X.p, c->BE.p, c->T.nBE, c->partial.p);
    CIPC_LAUNCH(k_final_sum, 1, RED_BT, 0, c->st, c->partial.p, RED_GRID, c->scal.p + 5, 1.0 / (double)c->T.nBE);
    CIPC_CUDA(cudaMemcpyAsync(h, c->scal.p + 5, sizeof(double), cudaMemcpyDeviceToHost, c->st));
    c->edgeLenPending = true;
}
double edge_len_end(cipc_ctx* c) // after a stream synchronisation
{
    if (c->edgeLenPending) {
        c->edgeLenCached = *(const double*)c->pinScal.p;
        c->edgeLenVersion = c->xVersion;
        c->edgeLenPending = false;
    }
    return c->edgeLenCached;
}
void bbox_to_host(cipc_ctx* c, const double4* P, double alpha, double* mn, double* mx)
{
    long long init[6];
    for (int d = 0; d < 3; ++d) { init[d] = dbl_ordered_h(1e300); init[3 + d] = dbl_ordered_h(-1e300); }
    CIPC_CUDA(cudaMemcpyAsync(c->bbox.p, init, sizeof(init), cudaMemcpyHostToDevice, c->st));
    CIPC_LAUNCH(k_bbox, RED_GRID, RED_BT, 0, c->st, c->X.p, P, alpha, c->BN.p, c->T.nBN, c->bbox.p);
    long long* out = (long long*)((char*)c->pinScal.reserve(64) + 8);
    CIPC_CUDA(cudaMemcpyAsync(out, c->bbox.p, 6 * sizeof(long long), cudaMemcpyDeviceToHost, c->st));
    CIPC_CUDA(cudaStreamSynchronize(c->st));
    for (int d = 0; d < 3; ++d) {
        long long a = out[d], b = out[3 + d];
        a = a >= 0 ? a : (a ^ 0x7fffffffffffffffLL);
        b = b >= 0 ? b : (b ^ 0x7fffffffffffffffLL);
        memcpy(&mn[d], &a, 8);
        memcpy(&mx[d], &b, 8);
    }
}
// after boxes + counts are in place: scan, emit, sort, cell table
void build_cell_lists(cipc_ctx* c, HashInfo& H, int which)
{
    const Topo& T = c->T;
    const int nP = T.nBN + T.nBE + T.nBT;
    Slab sl{0, 0, 1 << 30};
    cipc_ctx::SlabCache& sc = c->slabCache[which];
    if (c->world > 1 && sc.valid && sc.gx == H.G.gx && sc.gy == H.G.gy && sc.gz == H.G.gz && sc.nP == nP && sc.age < 8) {
        // The slab bounds only balance the load (any bounds give the same global candidate set, every rank uses the same
        // ones because every rank sees the same inputs); positions move little between Newton iterations, so the bounds of
        // the previous build on the same grid are kept for a few calls: no histogram pass, no host round trip.
        ++sc.age;
        sl = sc.sl;
    }
    else if (c->world > 1) {
        // this rank's slab along the axis with the most voxel layers, balanced on the number of hash entries
        const int gd[3] = {H.G.gx, H.G.gy, H.G.gz};
        sl.axis = (gd[0] >= gd[1] && gd[0] >= gd[2]) ? 0 : (gd[1] >= gd[2] ? 1 : 2);
        const int nl = gd[sl.axis];
        c->slabHist.reserve(nl, c->st);
        CIPC_CUDA(cudaMemsetAsync(c->slabHist.p, 0, (size_t)nl * 4, c->st));
        CIPC_LAUNCH(k_slab_hist, div_up(nP, TB), TB, 0, c->st, nP, c->boxLo.p, c->boxHi.p, sl.axis, nl, c->slabHist.p);
        const u32* hist = (const u32*)c->pinSmall.reserve((size_t)nl * 4);
        CIPC_CUDA(cudaMemcpyAsync((void*)hist, c->slabHist.p, (size_t)nl * 4, cudaMemcpyDeviceToHost, c->st));
        CIPC_CUDA(cudaStreamSynchronize(c->st));
        double total = 0;
        for (int i = 0; i < nl; ++i) total += hist[i];
        auto bound = [&](int r) { // first layer whose prefix reaches total * r / world
            if (r <= 0) return 0;
            if (r >= c->world) return nl;
            const double target = total * r / c->world;
            double acc = 0;
            for (int i = 0; i < nl; ++i) { acc += hist[i]; if (acc >= target) return i + 1; }
            return nl;
        };
        sl.s0 = bound(c->rank);
        sl.s1 = bound(c->rank + 1);
        sc.valid = true; sc.age = 0; sc.gx = H.G.gx; sc.gy = H.G.gy; sc.gz = H.G.gz; sc.nP = nP; sc.sl = sl;
    }
    const long gridCells = (long)H.G.gx * H.G.gy * H.G.gz;
    static const bool forceSort = getenv("CIPC_HASH_SORT") != nullptr; // cross-check switch: radix-sort path for every grid
    // dense cell table for grids of up to 4 cells per primitive, and for ANY grid of up to 2M cells: scanning a table that
    // size costs ~20 us, less than the radix-sort path's extra launches and round trip (sparse scenes such as a cloth over a
    // large obstacle: cfg1 0.28 -> 0.23 ms per build, tools/dense_min_cells.sh)
    static const long denseMinCells = getenv("CIPC_DENSE_MIN_CELLS") ? atol(getenv("CIPC_DENSE_MIN_CELLS")) : (2L << 20);
    if (gridCells <= DENSE_CELL_LIMIT && gridCells <= std::max(4L * nP, denseMinCells) && !forceSort) {
        // dense cell table (grids with at most a few cells per primitive): count -> scan (= kind-range table) -> tasks ->
        // fill; one host round trip for the entry and task totals
        const u32 nCells = (u32)gridCells;
        const size_t bins = (size_t)nCells * 4;
        u32 tot[2]; // entries, tasks
        u32* totDev = c->counters.p + 4;
        H.qch = std::min(256u, std::max(8u, (u32)(4L * nP / 50000L))); // entries are ~4 per primitive; the exact count is not known yet
        {
            cipc_ctx::Scope s1(c, "hb_count_scan");
            c->binNext.reserve(bins, c->st); c->ks.reserve(bins, c->st); c->taskCnt.reserve((size_t)2 * nCells, c->st);
            CIPC_CUDA(cudaMemsetAsync(c->binNext.p, 0, bins * 4, c->st));
            CIPC_LAUNCH(k_cell_count, div_up(nP, TB), TB, 0, c->st, T, H.G, c->boxLo.p, c->boxHi.p, sl, c->binNext.p);
            device_excl_scan(c->binNext.p, c->ks.p, bins, c->scanwk, c->st);
            CIPC_CUDA(cudaMemcpyAsync(totDev, c->scanwk.total.p, sizeof(u32), cudaMemcpyDeviceToDevice, c->st));
            CIPC_CUDA(cudaMemcpyAsync(c->binNext.p, c->ks.p, bins * 4, cudaMemcpyDeviceToDevice, c->st)); // running fill positions
            CIPC_LAUNCH(k_task_counts, div_up(nCells, TB), TB, 0, c->st, c->ks.p, nCells, H.qch, c->taskCnt.p);
            device_excl_scan(c->taskCnt.p, c->taskCnt.p, (size_t)2 * nCells, c->scanwk, c->st);
            CIPC_CUDA(cudaMemcpyAsync(c->nTasksDev.p, c->scanwk.total.p, sizeof(u32), cudaMemcpyDeviceToDevice, c->st));
            CIPC_CUDA(cudaMemcpyAsync(totDev + 1, c->scanwk.total.p, sizeof(u32), cudaMemcpyDeviceToDevice, c->st));
            CIPC_CUDA(cudaMemcpyAsync(tot, totDev, 2 * sizeof(u32), cudaMemcpyDeviceToHost, c->st));
        }
        CIPC_CUDA(cudaStreamSynchronize(c->st));
        const u32 nE = tot[0];
        H.nEntries = nE;
        H.nCells = nE ? nCells : 0;
        H.maxTasks = tot[1];
        c->ctr["hash_entries"] = nE;
        c->ctr["hash_cells"] = nCells;
        if (nE == 0) return;
        cipc_ctx::Scope s5(c, "hb_tables");
        c->vals.reserve(nE, c->st); c->codes.reserve(nE, c->st); c->taskDesc.reserve((size_t)tot[1] + 1, c->st);
        CIPC_LAUNCH(k_cell_fill, div_up(nP, TB), TB, 0, c->st, T, H.G, c->boxLo.p, c->boxHi.p, c->fine.p, sl, c->binNext.p, c->vals.p, c->codes.p);
        CIPC_LAUNCH(k_task_desc, div_up(nCells, TB), TB, 0, c->st, c->ks.p, nCells, H.qch, c->taskCnt.p, c->taskDesc.p);
        return;
    }
    // ---- sparse / very large grids: entries radix-sorted by (cell, kind)
    u32 nE;
    {
        cipc_ctx::Scope s1(c, "hb_count_scan");
        // per-primitive entry counts clipped to this rank's slab (the dense path counts per cell and never reads them)
        if (c->world > 1) CIPC_LAUNCH(k_clip_counts, div_up(nP, TB), TB, 0, c->st, nP, c->boxLo.p, c->boxHi.p, sl, c->cnt.p);
        device_excl_scan(c->cnt.p, c->cnt.p, nP, c->scanwk, c->st);
        CIPC_CUDA(cudaMemcpyAsync(&nE, c->scanwk.total.p, sizeof(u32), cudaMemcpyDeviceToHost, c->st));
    }
    CIPC_CUDA(cudaStreamSynchronize(c->st));
    H.nEntries = nE;
    c->keys.reserve(nE, c->st); c->vals.reserve(nE, c->st); c->heads.reserve(nE, c->st); c->headScan.reserve(nE, c->st);
    if (nE == 0) { H.nCells = 0; return; }
    {
        cipc_ctx::Scope s2(c, "hb_emit");
        CIPC_LAUNCH(k_emit_entries, div_up(nP, TB), TB, 0, c->st, T, H.G, c->boxLo.p, c->boxHi.p, c->cnt.p, sl, c->keys.p, c->vals.p);
    }
    const double cells = (double)H.G.gx * H.G.gy * H.G.gz;
    int bits = 2;
    while ((double)(1ULL << (bits - 2)) < cells) ++bits;
    {
        cipc_ctx::Scope s3(c, "hb_sort");
        device_radix_sort(c->keys.p, c->vals.p, nE, bits, c->sortwk, c->st);
    }
    u32 nCells;
    {
        cipc_ctx::Scope s4(c, "hb_heads");
        CIPC_LAUNCH(k_cell_heads, div_up(nE, TB), TB, 0, c->st, c->keys.p, nE, c->heads.p);
        device_excl_scan(c->heads.p, c->headScan.p, nE, c->scanwk, c->st);
        CIPC_CUDA(cudaMemcpyAsync(&nCells, c->scanwk.total.p, sizeof(u32), cudaMemcpyDeviceToHost, c->st));
    }
    CIPC_CUDA(cudaStreamSynchronize(c->st));
    cipc_ctx::Scope s5(c, "hb_tables");
    H.nCells = nCells;
    c->ks.reserve((size_t)nCells * 4, c->st);
    CIPC_LAUNCH(k_kind_starts, div_up((size_t)nE + 1, TB), TB, 0, c->st, c->keys.p, c->headScan.p, c->heads.p, nE, c->ks.p);
    // pair-enumeration tasks (count -> scan -> descriptors); the exact total stays on the device, the launch uses a bound
    H.qch = std::min(256u, std::max(8u, nE / 50000u)); // aim at >= ~50K tasks (5 per resident warp slot), whole cell sides beyond that
    H.maxTasks = 2 * nCells + nE / H.qch + 1;
    c->taskCnt.reserve((size_t)2 * nCells, c->st); c->taskDesc.reserve(H.maxTasks, c->st);
    CIPC_LAUNCH(k_task_counts, div_up(nCells, TB), TB, 0, c->st, c->ks.p, nCells, H.qch, c->taskCnt.p);
    device_excl_scan(c->taskCnt.p, c->taskCnt.p, (size_t)2 * nCells, c->scanwk, c->st);
    CIPC_CUDA(cudaMemcpyAsync(c->nTasksDev.p, c->scanwk.total.p, sizeof(u32), cudaMemcpyDeviceToDevice, c->st));
    CIPC_LAUNCH(k_task_desc, div_up(nCells, TB), TB, 0, c->st, c->ks.p, nCells, H.qch, c->taskCnt.p, c->taskDesc.p);
    c->codes.reserve(nE, c->st);
    CIPC_LAUNCH(k_entry_codes, div_up(nE, TB), TB, 0, c->st, c->keys.p, c->vals.p, nE, T.nBN, T.nBE, c->boxLo.p, c->fine.p, H.G, c->codes.p);
    c->ctr["hash_entries"] = nE;
    c->ctr["hash_cells"] = nCells;
}
// grid sizing shared by both passes (SPATIAL_HASH.h:71-86): voxelCount = ceil(range/voxel), total <= ~1e9
bool size_grid(const double* mn, const double* mx, double voxelSize, GridDesc& G, bool plusOne)
{
    const double range[3] = {mx[0] - mn[0], mx[1] - mn[1], mx[2] - mn[2]};
    double inv = 1.0 / voxelSize;
    double amt = 1;
    for (int d = 0; d < 3; ++d) amt *= std::max(1.0, std::ceil(range[d] * inv));
    if (amt > 1e9) {
        voxelSize *= std::pow(amt / 1.0e9, 1.0 / 3);
        inv = 1.0 / voxelSize;
    }
    long g[3];
    for (int d = 0; d < 3; ++d) g[d] = std::max(1L, (long)std::ceil(range[d] * inv)) + (plusOne ? 1 : 0);
    if ((double)g[0] * g[1] * g[2] >= (double)(1L << 30)) return false;
    const long gmax = std::max(g[0], std::max(g[1], g[2]));
    int fsh = 8; // 256 fine steps per voxel when they fit the 21 bits per axis of the packed boxes
    while (fsh > 0 && (gmax << fsh) > (1L << 21) - 1) --fsh;
    if ((gmax << fsh) > (1L << 21) - 1) return false;
    G.fsh = fsh;
    G.ox = mn[0]; G.oy = mn[1]; G.oz = mn[2]; G.inv = inv;
    G.gx = (int)g[0]; G.gy = (int)g[1]; G.gz = (int)g[2];
    return true;
}

// enumerate candidate pairs for this rank's slice of the sorted entries; grows buffers on overflow
template <bool CCD>
void run_pairs(cipc_ctx* c, const HashInfo& H, double dist, u32 counts[4])
{
    // the hash only holds this rank's slab of voxels (build_cell_lists), so every local cell is processed here
    const u32 e0 = 0, e1 = H.nCells;
    for (int k = 0; k < 4; ++k) counts[k] = 0;
    if (e1 <= e0 || H.maxTasks == 0) return;
    for (int attempt = 0; attempt < 3; ++attempt) {
        CandOut out;
        for (int k = 0; k < 4; ++k) {
            if (c->cand[k].cap == 0) c->cand[k].reserve(k < 2 ? (size_t)8 * (c->T.nBN + c->T.nBE) + 1024 : 1024, c->st);
            out.buf[k] = c->cand[k].p;
            out.cap[k] = (u32)std::min<size_t>(c->cand[k].cap, 0xfffffff0u);
        }
        out.count = c->counters.p;
        CIPC_CUDA(cudaMemsetAsync(c->counters.p, 0, 16 * sizeof(u32), c->st));
        CIPC_LAUNCH(k_pairs<CCD>, div_up(H.maxTasks, PAIRS_WARPS), PAIRS_WARPS * 32, 0, c->st, c->T, c->X.p, c->P.p, dist, c->vals.p, c->ks.p,
            c->codes.p, c->taskDesc.p, c->nTasksDev.p, H.qch, out);
        CIPC_CUDA(cudaMemcpyAsync(counts, c->counters.p, 4 * sizeof(u32), cudaMemcpyDeviceToHost, c->st));
        CIPC_CUDA(cudaStreamSynchronize(c->st));
        bool ok = true;
        for (int k = 0; k < 4; ++k)
            if (counts[k] > out.cap[k]) { ok = false; c->cand[k].reserve((size_t)counts[k] + counts[k] / 8, c->st); }
        if (ok) return;
    }
    throw std::runtime_error("candidate buffer kept overflowing");
}

BarrierParams make_bp(int elastic, double dHat2, const double* kappa, double thickness)
{
    if (elastic) thickness = 0;
    BarrierParams bp;
    bp.thickness2 = thickness * thickness;
    bp.dHat2 = dHat2 + 2 * std::sqrt(dHat2) * thickness;
    bp.k0 = kappa[0];
    bp.elastic = elastic;
    return bp;
}
int fetch_err(cipc_ctx* c)
{
    int e = 0;
    CIPC_CUDA(cudaMemcpyAsync(&e, c->errFlag.p, sizeof(int), cudaMemcpyDeviceToHost, c->st));
    CIPC_CUDA(cudaStreamSynchronize(c->st));
    return e;
}
void need(bool cond, const char* what)
{
    if (!cond) throw std::runtime_error(what);
}

// ---- stage bodies (results left on the device)
int do_barrier_energy(cipc_ctx* c, int elastic, double dHat2, const double* kappa, double thickness)
{
    need(c->haveX, "positions not set");
    const BarrierParams bp = make_bp(elastic, dHat2, kappa, thickness);
    CIPC_CUDA(cudaMemsetAsync(c->errFlag.p, 0, sizeof(int), c->st));
    {
        cipc_ctx::Scope sc(c, "barrier_E");
        CIPC_LAUNCH(k_barrier_energy, RED_GRID, RED_BT, 0, c->st, c->X.p, c->X0.p, c->cs.p, c->info.p, c->nC, bp, c->partial.p,
            c->errFlag.p);
        CIPC_LAUNCH(k_final_sum, 1, RED_BT, 0, c->st, c->partial.p, RED_GRID, c->scal.p + 0, 1.0);
    }
    return CIPC_OK;
}
int do_barrier_gradient(cipc_ctx* c, int elastic, double dHat2, const double* kappa, double thickness)
{
    need(c->haveX, "positions not set");
    const BarrierParams bp = make_bp(elastic, dHat2, kappa, thickness);
    c->g.reserve((size_t)3 * c->T.nV + 8, c->st);
    cipc_ctx::Scope sc(c, "barrier_g");
    CIPC_CUDA(cudaMemsetAsync(c->g.p, 0, (size_t)3 * c->T.nV * sizeof(double), c->st));
    if (c->nC) CIPC_LAUNCH(k_barrier_gradient, div_up(c->nC, 128), 128, 0, c->st, c->X.p, c->X0.p, c->cs.p, c->info.p, c->nC, bp, c->g.p, (const u32*)nullptr);
    return CIPC_OK;
}
int do_step_size(cipc_ctx* c, int elastic, double thickness, double stepIn)
{
    need(c->haveX && c->haveP, "positions / search direction not set");
    if (elastic) thickness = 0;
    const Topo& T = c->T;
    HashInfo H;
    double alpha = stepIn;
    {
        cipc_ctx::Scope sc(c, "ccd_hash_build");
        edge_len_begin(c); // rides on the search-direction reduction's round trip
        // span rule (SPATIAL_HASH.h:466-482).  The sum is a fixed-tree reduction; when the rule
        // fires the step is biased down by 4e-13 relative so that it never exceeds the sequential
        // CPU sum's result.
        // one host round trip for the search-direction reduction AND the swept bounding box: the box is launched for the
        // unshrunk step before the span rule is known and redone only when the rule fires (rare)
        CIPC_LAUNCH(k_psize_partial, RED_GRID, RED_BT, 0, c->st, c->P.p, c->BN.p, T.nBN, c->partial.p);
        CIPC_LAUNCH(k_final_sum, 1, RED_BT, 0, c->st, c->partial.p, RED_GRID, c->scal.p + 3, 1.0 / ((double)T.nBN * 3.0));
        double* hp = (double*)((char*)c->pinScal.reserve(128) + 64);
        CIPC_CUDA(cudaMemcpyAsync(hp, c->scal.p + 3, sizeof(double), cudaMemcpyDeviceToHost, c->st));
        double mn[3], mx[3];
        bbox_to_host(c, c->P.p, alpha, mn, mx); // synchronises the stream
        const double pSize = *hp;
        double voxelSize = 1.0;
        if (T.nBE) voxelSize *= edge_len_end(c);
        const double spanSize = alpha * pSize / voxelSize;
        if (spanSize > 1) {
            alpha = (alpha / spanSize) * (1.0 - 4e-13);
            bbox_to_host(c, c->P.p, alpha, mn, mx);
        }
        for (int d = 0; d < 3; ++d) { mn[d] -= thickness / 2; mx[d] += thickness / 2; }
        if (!size_grid(mn, mx, voxelSize, H.G, true)) return CIPC_ERR_GRID;
        const int nP = T.nBN + T.nBE + T.nBT;
        c->boxLo.reserve(nP, c->st); c->boxHi.reserve(nP, c->st); c->cnt.reserve(nP, c->st);
        c->nodeLo.reserve(T.nBN, c->st); c->nodeHi.reserve(T.nBN, c->st); c->nodeFine.reserve(T.nBN, c->st); c->fine.reserve(nP, c->st);
        double mag = 0;
        for (int d = 0; d < 3; ++d) mag = std::max(mag, std::max(std::fabs(mn[d]), std::fabs(mx[d])));
        const double guard = 16.0 * 2.220446049250313e-16 * (mag + 1.0 / H.G.inv);
        CIPC_LAUNCH(k_node_boxes_ccd, div_up(T.nBN, TB), TB, 0, c->st, T, c->X.p, c->P.p, alpha, thickness / 2, guard, H.G, c->nodeLo.p, c->nodeHi.p,
            c->nodeFine.p);
        CIPC_LAUNCH(k_prim_boxes_ccd, div_up(nP, TB), TB, 0, c->st, T, c->nodeLo.p, c->nodeHi.p, c->nodeFine.p, c->boxLo.p, c->boxHi.p, c->fine.p,
            c->cnt.p);
        build_cell_lists(c, H, 1);
    }
    u32 counts[4];
    {
        cipc_ctx::Scope sc(c, "ccd_pairs");
        run_pairs<true>(c, H, thickness, counts);
    }
    c->ctr["ccd_pairs"] = (int64_t)counts[0] + counts[1] + counts[2] + counts[3];
    c->ctr["ccd_pairs_ee"] = counts[1];
    {
        cipc_ctx::Scope sc(c, "ccd_accd");
        u64 bits;
        memcpy(&bits, &alpha, 8);
        CIPC_CUDA(cudaMemcpyAsync(c->scal.p + 1, &bits, 8, cudaMemcpyHostToDevice, c->st));
        CIPC_CUDA(cudaMemsetAsync(c->errFlag.p, 0, sizeof(int), c->st));
        AccdOut out{(u64*)(c->scal.p + 1), c->errFlag.p};
        {
            cipc_ctx::Scope sp(c, "ccd_accd_pt");
            if (counts[0]) CIPC_LAUNCH(k_accd_pt, div_up(counts[0], 128), 128, 0, c->st, T, c->X.p, c->P.p, c->cand[0].p, counts[0], thickness, alpha, out);
            if (counts[2]) CIPC_LAUNCH(k_accd_pe, div_up(counts[2], 128), 128, 0, c->st, T, c->X.p, c->P.p, c->cand[2].p, counts[2], thickness, alpha, out);
            if (counts[3]) CIPC_LAUNCH(k_accd_pp, div_up(counts[3], 128), 128, 0, c->st, T, c->X.p, c->P.p, c->cand[3].p, counts[3], thickness, alpha, out);
        }
        {
            cipc_ctx::Scope se(c, "ccd_accd_ee");
            if (counts[1]) CIPC_LAUNCH(k_accd_ee, div_up(counts[1], 128), 128, 0, c->st, T, c->X.p, c->P.p, c->cand[1].p, counts[1], thickness, alpha, out);
        }
    }
    return CIPC_OK;
}
int do_min_dist(cipc_ctx* c, bool wantDist)
{
    need(c->haveX, "positions not set");
    cipc_ctx::Scope sc(c, "min_dist");
    const long long init = dbl_ordered_h(1e300);
    CIPC_CUDA(cudaMemcpyAsync(c->scal.p + 2, &init, 8, cudaMemcpyHostToDevice, c->st));
    if (wantDist) c->dist2.reserve(c->nC, c->st);
    if (c->nC) CIPC_LAUNCH(k_min_dist, RED_GRID, RED_BT, 0, c->st, c->X.p, c->cs.p, c->nC, wantDist ? c->dist2.p : nullptr, (long long*)(c->scal.p + 2));
    return CIPC_OK;
}

enum HessOutF { HF_HOST = 0, HF_DEV = 1, HF_BLK = 2 };
// ---- device-side merge of the block stream of the Hessian computed last (merge.cuh): blkVal holds `tot` upper blocks of the
// `nSt` stencils `st` at the offsets in tripOff; afterwards (mRow, mCol, mVal) are the nUm unique upper blocks sorted by
// (row, col), nDiag of them diagonal, and nTrip = 9 (nUm + off-diagonal) is the number of MERGED triplets of the full matrix
void merge_blocks(cipc_ctx* c, const int4* st, u32 nSt, u32 tot)
{
    cipc_ctx::Scope sc(c, "hessian_merge");
    const int nV = c->T.nV;
    need(nV < (1 << 28), "block merge: more than 2^28 nodes");
    c->rowStart.reserve((size_t)nV + 1, c->st); c->rowCur.reserve((size_t)nV + 1, c->st);
    c->ent.reserve((size_t)tot + 1, c->st);
    u32* nRowsDev = c->counters.p + 7;
    {
        cipc_ctx::Scope s1(c, "mg_count");
        CIPC_CUDA(cudaMemsetAsync(c->rowCur.p, 0, ((size_t)nV + 1) * 4, c->st));
        CIPC_LAUNCH(k_blk_rows<false>, div_up(nSt, TB), TB, 0, c->st, st, c->tripOff.p, nSt, c->rowCur.p, (u64*)nullptr);
        device_excl_scan(c->rowCur.p, c->rowStart.p, (size_t)nV + 1, c->scanwk, c->st);
        CIPC_CUDA(cudaMemcpyAsync(c->rowCur.p, c->rowStart.p, ((size_t)nV + 1) * 4, cudaMemcpyDeviceToDevice, c->st));
    }
    {
        cipc_ctx::Scope s1(c, "mg_scatter");
        CIPC_LAUNCH(k_blk_rows<true>, div_up(nSt, TB), TB, 0, c->st, st, c->tripOff.p, nSt, c->rowCur.p, c->ent.p);
    }
    const u32 nGrp = div_up(nV, MRS_ROWS);
    c->mHeads.reserve((size_t)nGrp + 1, c->st); c->mScan.reserve((size_t)nGrp + 1, c->st);
    {
        cipc_ctx::Scope s1(c, "mg_sort");
        CIPC_CUDA(cudaMemsetAsync(nRowsDev, 0, 4, c->st));
        CIPC_LAUNCH(k_row_sort, nGrp, MRS_BT, 0, c->st, c->rowStart.p, nV, c->ent.p, c->mHeads.p, nRowsDev);
        device_excl_scan(c->mHeads.p, c->mScan.p, nGrp, c->scanwk, c->st);
    }
    u32 h[2];
    CIPC_CUDA(cudaMemcpyAsync(&h[0], c->scanwk.total.p, 4, cudaMemcpyDeviceToHost, c->st));
    CIPC_CUDA(cudaMemcpyAsync(&h[1], nRowsDev, 4, cudaMemcpyDeviceToHost, c->st));
    CIPC_CUDA(cudaStreamSynchronize(c->st));
    const u32 nU = h[0];
    c->mRow.reserve((size_t)nU + 1, c->st); c->mCol.reserve((size_t)nU + 1, c->st); c->mStart.reserve((size_t)nU + 1, c->st);
    c->mVal.reserve((size_t)nU * 9 + 16, c->st);
    {
        cipc_ctx::Scope s1(c, "mg_uniq");
        CIPC_LAUNCH(k_blk_uniq, nGrp, MRS_BT, 0, c->st, c->rowStart.p, nV, c->ent.p, c->mScan.p, c->mRow.p, c->mCol.p, c->mStart.p);
    }
    {
        cipc_ctx::Scope s1(c, "mg_sum");
        CIPC_LAUNCH(k_blk_sum, div_up(nU, 32), 288, 0, c->st, c->blkVal.p, c->ent.p, c->mStart.p, nU, tot, c->mVal.p);
    }
    c->nUm = nU;
    c->nDiag = h[1]; // every vertex of a stencil owns a diagonal block, which opens its row
    c->nTrip = (int64_t)9 * ((int64_t)nU + (int64_t)(nU - h[1]));
    c->mergedValid = true;
    c->ctr["merge_blocks_in"] = tot; c->ctr["merge_blocks_unique"] = nU; c->ctr["merge_blocks_diag"] = h[1];
}

// ---- friction stage bodies (FEM/FRICTION.h)
int do_friction_basis(cipc_ctx* c, int elastic, double dHat2, const double* kappa, double thickness)
{
    need(c->haveX, "positions not set");
    const BarrierParams bp = make_bp(elastic, dHat2, kappa, thickness);
    c->nF = 0;
    if (!c->nC) return CIPC_OK;
    cipc_ctx::Scope sc(c, "friction_basis");
    c->fslot.reserve(c->nC, c->st);
    CIPC_LAUNCH(k_friction_flags, div_up(c->nC, TB), TB, 0, c->st, c->cs.p, c->nC, c->fslot.p);
    device_excl_scan(c->fslot.p, c->fslot.p, c->nC, c->scanwk, c->st);
    u32 nF;
    CIPC_CUDA(cudaMemcpyAsync(&nF, c->scanwk.total.p, sizeof(u32), cudaMemcpyDeviceToHost, c->st));
    CIPC_CUDA(cudaStreamSynchronize(c->st));
    c->fcs.reserve((size_t)nF + 1, c->st); c->fcp.reserve((size_t)nF + 1, c->st);
    c->fB.reserve((size_t)nF * 6 + 2, c->st); c->fnf.reserve((size_t)nF + 1, c->st);
    CIPC_LAUNCH(k_friction_basis, div_up(c->nC, 128), 128, 0, c->st, c->X.p, c->cs.p, c->info.p, c->fslot.p, c->nC, bp, c->fcs.p, c->fcp.p,
        c->fB.p, c->fnf.p);
    c->nF = nF;
    c->ctr["friction_constraints"] = nF;
    return CIPC_OK;
}
int do_friction_energy(cipc_ctx* c, double epsvh2, double mu)
{
    need(c->haveX && c->haveXn, "positions / previous positions not set");
    cipc_ctx::Scope sc(c, "friction_E");
    CIPC_LAUNCH(k_friction_energy, RED_GRID, RED_BT, 0, c->st, c->X.p, c->Xn.p, c->fcs.p, c->fcp.p, c->fB.p, c->fnf.p, c->nF, std::sqrt(epsvh2),
        c->partial.p);
    CIPC_LAUNCH(k_final_sum, 1, RED_BT, 0, c->st, c->partial.p, RED_GRID, c->scal.p + 4, mu);
    return CIPC_OK;
}
int do_friction_gradient(cipc_ctx* c, double epsvh2, double mu, bool accumulate)
{
    need(c->haveX && c->haveXn, "positions / previous positions not set");
    c->g.reserve((size_t)3 * c->T.nV + 8, c->st, true);
    cipc_ctx::Scope sc(c, "friction_g");
    if (!accumulate) CIPC_CUDA(cudaMemsetAsync(c->g.p, 0, (size_t)3 * c->T.nV * sizeof(double), c->st));
    if (c->nF) CIPC_LAUNCH(k_friction_gradient, div_up(c->nF, 128), 128, 0, c->st, c->X.p, c->Xn.p, c->fcs.p, c->fcp.p, c->fB.p, c->fnf.p, c->nF,
        std::sqrt(epsvh2), mu, c->g.p);
    return CIPC_OK;
}
int do_friction_hessian(cipc_ctx* c, double epsvh2, double mu, HessOutF mode)
{
    need(c->haveX && c->haveXn, "positions / previous positions not set");
    const bool devTriplets = mode != HF_HOST, blk = mode == HF_BLK;
    c->nTrip = 0;
    c->factorValid = false;
    c->mergedValid = false;
    c->hessStencils = c->fcs.p; c->hessN = c->nF;
    if (!c->nF) return CIPC_OK;
    cipc_ctx::Scope sc(c, "friction_H");
    const u32 nF = c->nF;
    c->tripOff.reserve(nF, c->st);
    if (blk) CIPC_LAUNCH(k_blk_sizes, div_up(nF, TB), TB, 0, c->st, c->fcs.p, nF, c->tripOff.p);
    else CIPC_LAUNCH(k_block_sizes, div_up(nF, TB), TB, 0, c->st, c->fcs.p, nF, c->tripOff.p);
    device_excl_scan(c->tripOff.p, c->tripOff.p, nF, c->scanwk, c->st);
    for (int k = 0; k < 4; ++k) c->clsIdx[k].reserve(nF, c->st);
    CIPC_CUDA(cudaMemsetAsync(c->counters.p + 12, 0, 4 * sizeof(u32), c->st)); // also zeroes counters[15], the dense-list length
    CIPC_LAUNCH(k_classify, div_up(nF, TB), TB, 0, c->st, c->fcs.p, nF, c->clsIdx[0].p, c->clsIdx[1].p, c->clsIdx[2].p, c->clsIdx[3].p,
        c->counters.p + 12);
    u32 tot, nk[4];
    CIPC_CUDA(cudaMemcpyAsync(&tot, c->scanwk.total.p, sizeof(u32), cudaMemcpyDeviceToHost, c->st));
    CIPC_CUDA(cudaMemcpyAsync(nk, c->counters.p + 12, 4 * sizeof(u32), cudaMemcpyDeviceToHost, c->st));
    CIPC_CUDA(cudaStreamSynchronize(c->st));
    need(nk[3] == 0, "mollified stencil in the friction constraint set");
    void* outp;
    if (blk) {
        c->blkVal.reserve((size_t)tot * 9 + 16, c->st);
        c->nBlkLast = tot;
        outp = c->blkVal.p;
    }
    else {
        c->nTrip = (int64_t)tot * 9;
        c->trip.reserve((size_t)c->nTrip, c->st);
        outp = c->trip.p;
    }
    const double epsvh = std::sqrt(epsvh2);
    if (devTriplets) {
        // device-resident output: factors go through shared memory only (k_friction_fused)
        cipc_ctx::Scope sk(c, "k_friction_fused");
#define CIPC_FRF(CLS, B) CIPC_LAUNCH((k_friction_fused<CLS, B>), div_up(nk[CLS], FUSED_BD), FUSED_BD, FrFusedShape<CLS>::SMEM, c->st, c->X.p, c->Xn.p, \
    c->fcs.p, c->fcp.p, c->fB.p, c->fnf.p, c->tripOff.p, c->clsIdx[CLS].p, nk[CLS], epsvh, epsvh2, mu, outp)
        if (nk[0]) { if (blk) CIPC_FRF(0, true); else CIPC_FRF(0, false); }
        if (nk[1]) { if (blk) CIPC_FRF(1, true); else CIPC_FRF(1, false); }
        if (nk[2]) { if (blk) CIPC_FRF(2, true); else CIPC_FRF(2, false); }
#undef CIPC_FRF
        for (int k = 0; k < 4; ++k) c->nk[k] = nk[k];
    }
    else {
    const size_t ydoubles = (size_t)nk[0] * 24 + (size_t)nk[1] * 18 + (size_t)nk[2] * 12;
    c->Y.reserve(ydoubles + 4, c->st);
    c->yhdr.reserve((size_t)nk[0] + nk[1] + nk[2] + 1, c->st);
    double* Y0 = c->Y.p; double* Y1 = Y0 + (size_t)nk[0] * 24; double* Y2 = Y1 + (size_t)nk[1] * 18;
    YHdr* h0 = (YHdr*)c->yhdr.p; YHdr* h1 = h0 + nk[0]; YHdr* h2 = h1 + nk[1];
    {
        cipc_ctx::Scope sk(c, "k_friction_factor");
        if (nk[0]) CIPC_LAUNCH(k_friction_factor<0>, div_up(nk[0], 128), 128, 0, c->st, c->X.p, c->Xn.p, c->fcs.p, c->fcp.p, c->fB.p, c->fnf.p,
            c->tripOff.p, c->clsIdx[0].p, nk[0], epsvh, epsvh2, mu, Y0, h0);
        if (nk[1]) CIPC_LAUNCH(k_friction_factor<1>, div_up(nk[1], 128), 128, 0, c->st, c->X.p, c->Xn.p, c->fcs.p, c->fcp.p, c->fB.p, c->fnf.p,
            c->tripOff.p, c->clsIdx[1].p, nk[1], epsvh, epsvh2, mu, Y1, h1);
        if (nk[2]) CIPC_LAUNCH(k_friction_factor<2>, div_up(nk[2], 128), 128, 0, c->st, c->X.p, c->Xn.p, c->fcs.p, c->fcp.p, c->fB.p, c->fnf.p,
            c->tripOff.p, c->clsIdx[2].p, nk[2], epsvh, epsvh2, mu, Y2, h2);
    }
    c->factorValid = true;
    c->expanded = false;
    c->Y0 = Y0; c->Y1 = Y1; c->Y2 = Y2; c->h0 = h0; c->h1 = h1; c->h2 = h2;
    for (int k = 0; k < 4; ++k) c->nk[k] = nk[k];
    c->ny[0] = c->ny[1] = c->ny[2] = 2;
    }
    c->ctr["friction_4pt"] = nk[0]; c->ctr["friction_pe"] = nk[1]; c->ctr["friction_pp"] = nk[2];
    if (blk) merge_blocks(c, c->fcs.p, nF, tot);
    return CIPC_OK;
}

} // namespace

// ---- triplet delivery
// device side: expand the factors into the triplet stream (needed only when the triplets are consumed on the GPU)
template <int NB, int NY>
void launch_expand(cipc_ctx* c, const double* Y, const void* hdr, u32 n)
{
    if (!n) return;
    // product path: tiled expansion, 16 stencils per CTA, streaming stores (0.99 of the measured copy peak on cfg5_1m);
    // CIPC_EXPAND_VARIANT=0 selects the one-thread-per-triplet kernel as a cross-check
    static const int variant = getenv("CIPC_EXPAND_VARIANT") ? atoi(getenv("CIPC_EXPAND_VARIANT")) : 1;
    constexpr int PER = 9 * NB * NB;
    if (variant == 0) CIPC_LAUNCH((k_hessian_expand<NB, NY>), div_up((u64)n * PER, 256), 256, 0, c->st, Y, (const YHdr*)hdr, (u64)n * PER, c->trip.p);
    else CIPC_LAUNCH((k_hessian_expand_tiled<NB, NY, 16, true>), div_up(n, 16), 256, 0, c->st, Y, (const YHdr*)hdr, n, c->trip.p);
}
void expand_on_device(cipc_ctx* c)
{
    if (!c->factorValid || c->expanded) return;
    const u32* nk = c->nk;
    const bool fr = c->ny[0] == 2; // friction factors: two vectors per stencil in every class
    {
        cipc_ctx::Scope sk(c, "k_barrier_hessian"); // the dominant kernel: triplet expansion of the PT/EE blocks
        if (fr) launch_expand<4, 2>(c, c->Y0, c->h0, nk[0]); else launch_expand<4, 3>(c, c->Y0, c->h0, nk[0]);
    }
    launch_expand<3, 2>(c, c->Y1, c->h1, nk[1]);
    if (fr) launch_expand<2, 2>(c, c->Y2, c->h2, nk[2]); else launch_expand<2, 1>(c, c->Y2, c->h2, nk[2]);
    c->expanded = true;
}
// host side: y y^T expansion of one stencil with streaming 16-byte stores (the destination is never read back).
// Values of one row are formed two at a time in SSE registers; (row, col) pairs are pre-packed per stencil.
template <int NB, int NY>
inline void expand_one_host(const double* y, const YHdr& h, cipc_triplet* out, bool aligned)
{
    constexpr int NN = 3 * NB, NP = (NN + 1) / 2 * 2;
    unsigned idx[NP];
    for (int r = 0; r < NN; ++r) idx[r] = (unsigned)(h.v[r / 3] * 3 + r % 3);
    double yy[NY][NP];
    for (int k = 0; k < NY; ++k) {
        for (int c = 0; c < NN; ++c) yy[k][c] = y[k * NN + c];
        for (int c = NN; c < NP; ++c) yy[k][c] = 0.0;
    }
    const double sgn = h.pad[0] ? -1.0 : 1.0;
    __m128i* o = reinterpret_cast<__m128i*>(out + (size_t)h.off * 9);
    for (int r = 0; r < NN; ++r) {
        const long long rr = (long long)idx[r];
        __m128d yr[NY];
        for (int k = 0; k < NY; ++k) yr[k] = _mm_set1_pd(sgn * yy[k][r]);
        for (int c = 0; c < NN; c += 2) {
            __m128d v = _mm_mul_pd(yr[0], _mm_loadu_pd(&yy[0][c]));
            for (int k = 1; k < NY; ++k) v = _mm_add_pd(v, _mm_mul_pd(yr[k], _mm_loadu_pd(&yy[k][c])));
            const __m128i vi = _mm_castpd_si128(v);
            const __m128i t0 = _mm_unpacklo_epi64(_mm_cvtsi64_si128(((long long)idx[c] << 32) | rr), vi);
            if (aligned) _mm_stream_si128(o + r * NN + c, t0); else _mm_storeu_si128(o + r * NN + c, t0);
            if (c + 1 < NN) {
                const __m128i t1 = _mm_unpacklo_epi64(_mm_cvtsi64_si128(((long long)idx[c + 1] << 32) | rr), _mm_unpackhi_epi64(vi, vi));
                if (aligned) _mm_stream_si128(o + r * NN + c + 1, t1); else _mm_storeu_si128(o + r * NN + c + 1, t1);
            }
        }
    }
}
// Delivers the triplets of the last projected Hessian to host memory: ships the compact factors (8x fewer
// bytes over PCIe than the triplet stream) and expands them with all host threads.
void deliver_triplets_host(cipc_ctx* c, cipc_triplet* out)
{
    const u32* nk = c->nk;
    const size_t yd0 = (size_t)c->ny[0] * 12, yd1 = (size_t)c->ny[1] * 9, yd2 = (size_t)c->ny[2] * 6; // doubles per stencil
    const bool fr = c->ny[0] == 2;
    const size_t yD = (size_t)nk[0] * yd0 + (size_t)nk[1] * yd1 + (size_t)nk[2] * yd2, nH = (size_t)nk[0] + nk[1] + nk[2];
    double* hY = (double*)c->pinY.reserve(yD * 8 + 64);
    YHdr* hH = (YHdr*)c->pinH.reserve(nH * sizeof(YHdr) + 64);
    u32 nDense = 0;
    CIPC_CUDA(cudaMemcpyAsync(&nDense, c->counters.p + 15, 4, cudaMemcpyDeviceToHost, c->st));
    if (nH) CIPC_CUDA(cudaMemcpyAsync(hH, c->yhdr.p, nH * sizeof(YHdr), cudaMemcpyDeviceToHost, c->st));
    CIPC_CUDA(cudaStreamSynchronize(c->st));
    // the factors are copied in pieces; host threads start expanding as soon as the piece they need has landed
    const size_t PIECE = (size_t)8 << 20; // doubles (64 MiB)
    std::vector<cudaEvent_t> evs;
    std::vector<size_t> evEnd;
    for (size_t o = 0; o < yD; o += PIECE) {
        const size_t len = std::min(PIECE, yD - o);
        CIPC_CUDA(cudaMemcpyAsync(hY + o, c->Y.p + o, len * 8, cudaMemcpyDeviceToHost, c->st));
        cudaEvent_t e;
        CIPC_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        CIPC_CUDA(cudaEventRecord(e, c->st));
        evs.push_back(e);
        evEnd.push_back(o + len);
    }
    std::atomic<size_t> landed(0);
    cipc_triplet* hD = nullptr;
    uint2* hM = nullptr;
    if (nDense) {
        c->denseBuf.reserve((size_t)nDense * 144, c->st);
        c->denseMeta.reserve(nDense, c->st);
        CIPC_LAUNCH(k_gather_dense, std::min<u32>(nDense, 4736u), 64, 0, c->st, c->cs.p, c->tripOff.p, c->clsIdx[3].p, c->counters.p + 15, c->trip.p,
            c->denseBuf.p, c->denseMeta.p);
        hD = (cipc_triplet*)c->pinD.reserve((size_t)nDense * 144 * sizeof(cipc_triplet));
        hM = (uint2*)c->pinM.reserve((size_t)nDense * sizeof(uint2));
        CIPC_CUDA(cudaMemcpyAsync(hD, c->denseBuf.p, (size_t)nDense * 144 * sizeof(cipc_triplet), cudaMemcpyDeviceToHost, c->st));
        CIPC_CUDA(cudaMemcpyAsync(hM, c->denseMeta.p, (size_t)nDense * sizeof(uint2), cudaMemcpyDeviceToHost, c->st));
    }
    // expansion of the factored stencils overlaps the dense D2H
    const bool aligned = (reinterpret_cast<uintptr_t>(out) & 15) == 0;
    const double* y0 = hY; const double* y1 = y0 + (size_t)nk[0] * yd0; const double* y2 = y1 + (size_t)nk[1] * yd1;
    const YHdr* g0 = hH; const YHdr* g1 = g0 + nk[0]; const YHdr* g2 = g1 + nk[1];
    advise_huge(out, (size_t)c->nTrip * sizeof(cipc_triplet));
    const size_t CH = 2048; // stencils per work item
    const size_t items0 = (nk[0] + CH - 1) / CH, items1 = (nk[1] + CH - 1) / CH, items2 = (nk[2] + CH - 1) / CH;
    std::atomic<size_t> next(0);
    auto wait_for = [&](size_t needEnd) {
        while (landed.load(std::memory_order_acquire) < needEnd) std::this_thread::yield();
    };
    auto worker = [&]() {
        for (;;) {
            size_t it = next.fetch_add(1);
            if (it >= items0 + items1 + items2) break;
            if (it < items0) {
                const size_t a = it * CH, b = std::min<size_t>(a + CH, nk[0]);
                wait_for(b * yd0);
                if (fr) { for (size_t q = a; q < b; ++q) if (g0[q].off != 0xffffffffu) expand_one_host<4, 2>(y0 + q * yd0, g0[q], out, aligned); }
                else for (size_t q = a; q < b; ++q) if (g0[q].off != 0xffffffffu) expand_one_host<4, 3>(y0 + q * yd0, g0[q], out, aligned);
            }
            else if (it < items0 + items1) {
                const size_t a = (it - items0) * CH, b = std::min<size_t>(a + CH, nk[1]);
                wait_for((size_t)nk[0] * yd0 + b * yd1);
                for (size_t q = a; q < b; ++q) if (g1[q].off != 0xffffffffu) expand_one_host<3, 2>(y1 + q * yd1, g1[q], out, aligned);
            }
            else {
                const size_t a = (it - items0 - items1) * CH, b = std::min<size_t>(a + CH, nk[2]);
                wait_for((size_t)nk[0] * yd0 + (size_t)nk[1] * yd1 + b * yd2);
                if (fr) { for (size_t q = a; q < b; ++q) if (g2[q].off != 0xffffffffu) expand_one_host<2, 2>(y2 + q * yd2, g2[q], out, aligned); }
                else for (size_t q = a; q < b; ++q) if (g2[q].off != 0xffffffffu) expand_one_host<2, 1>(y2 + q * yd2, g2[q], out, aligned);
            }
        }
        _mm_sfence();
    };
    bool copyFailed = false;
    HostPool::get().run([&](int tid) {
        if (tid == 0) { // the calling thread publishes the pieces as they land, then joins the work
            for (size_t k = 0; k < evs.size(); ++k) {
                if (cudaEventSynchronize(evs[k]) != cudaSuccess) copyFailed = true;
                landed.store(copyFailed ? yD : evEnd[k], std::memory_order_release);
            }
            landed.store(yD, std::memory_order_release);
        }
        worker();
    });
    for (auto e : evs) cudaEventDestroy(e);
    if (copyFailed) throw CudaError("device-to-host copy of the Hessian factors failed");
    if (nDense) {
        CIPC_CUDA(cudaStreamSynchronize(c->st));
        for (u32 k = 0; k < nDense; ++k) memcpy(out + (size_t)hM[k].x * 9, hD + (size_t)k * 144, (size_t)hM[k].y * sizeof(cipc_triplet));
    }
}

// Delivers the MERGED triplets (cipc_barrier_hessian_merged / cipc_friction_hessian_merged): the unique upper blocks cross PCIe
// (80 bytes per block: row, col, 9 values) in pieces and the host threads write, for every block, its 9 triplets and --
// off-diagonal blocks -- the 9 triplets of the transposed block.  Layout of `out`: [0, 9 nU) the upper blocks in (row, col)
// order, [9 nU, 9 (nU + nOff)) the mirrored blocks in the same order.
void deliver_merged_host(cipc_ctx* c, cipc_triplet* out)
{
    const size_t nU = c->nUm;
    if (!nU) return;
    const auto tp0 = std::chrono::steady_clock::now();
    auto us_since = [&](std::chrono::steady_clock::time_point t) { return (int64_t)std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now() - t).count(); };
    u32* hR = (u32*)c->pinMK.reserve(nU * 8 + 64);
    u32* hC = hR + nU;
    double* hV = (double*)c->pinMV.reserve(nU * 72 + 64);
    CIPC_CUDA(cudaMemcpyAsync(hR, c->mRow.p, nU * 4, cudaMemcpyDeviceToHost, c->st));
    CIPC_CUDA(cudaMemcpyAsync(hC, c->mCol.p, nU * 4, cudaMemcpyDeviceToHost, c->st));
    cudaEvent_t keysEv;
    CIPC_CUDA(cudaEventCreateWithFlags(&keysEv, cudaEventDisableTiming));
    CIPC_CUDA(cudaEventRecord(keysEv, c->st));
    const size_t PIECE = (size_t)1 << 19; // blocks per piece (36 MiB of values)
    const size_t nPieces = (nU + PIECE - 1) / PIECE;
    std::vector<cudaEvent_t> evs(nPieces);
    for (size_t k = 0; k < nPieces; ++k) {
        const size_t a = k * PIECE, b = std::min(nU, a + PIECE);
        CIPC_CUDA(cudaMemcpyAsync(hV + a * 9, c->mVal.p + a * 9, (b - a) * 72, cudaMemcpyDeviceToHost, c->st));
        CIPC_CUDA(cudaEventCreateWithFlags(&evs[k], cudaEventDisableTiming));
        CIPC_CUDA(cudaEventRecord(evs[k], c->st));
    }
    advise_huge(out, (size_t)c->nTrip * sizeof(cipc_triplet));
    bool failed = cudaEventSynchronize(keysEv) != cudaSuccess;
    cudaEventDestroy(keysEv);
    c->ctr["deliver_us_keys"] = us_since(tp0);
    // off-diagonal blocks before each chunk of CH blocks (the mirrored region is indexed by the off-diagonal rank)
    const size_t CH = 4096, nChunks = (nU + CH - 1) / CH; // PIECE is a multiple of CH
    std::vector<size_t> offBefore(nChunks + 1, 0);
    HostPool::get().for_each(nChunks, [&](size_t k) {
        const size_t a = k * CH, b = std::min(nU, a + CH);
        size_t n = 0;
        for (size_t u = a; u < b; ++u) n += hR[u] != hC[u];
        offBefore[k + 1] = n;
    });
    for (size_t k = 0; k < nChunks; ++k) offBefore[k + 1] += offBefore[k];
    c->ctr["deliver_us_count"] = us_since(tp0);
    const bool aligned = (reinterpret_cast<uintptr_t>(out) & 15) == 0;
    std::atomic<size_t> landed(0), next(0);
    // entry e = 3 i + j of a block: index qword (col << 32 | row) = base + (j << 32 | i); two entries per 128-bit register.  The
    // mirrored block's entries are written in the same order with the halves of the index qword swapped (row <-> col).
    const __m128i o01 = _mm_set_epi64x((1LL << 32) | 0, (0LL << 32) | 0), o23 = _mm_set_epi64x((0LL << 32) | 1, (2LL << 32) | 0),
                  o45 = _mm_set_epi64x((2LL << 32) | 1, (1LL << 32) | 1), o67 = _mm_set_epi64x((1LL << 32) | 2, (0LL << 32) | 2),
                  o8 = _mm_set_epi64x(0, (2LL << 32) | 2);
    auto st = [&](__m128i* o, __m128i t) { if (aligned) _mm_stream_si128(o, t); else _mm_storeu_si128(o, t); };
    auto worker = [&]() {
        for (size_t k; (k = next.fetch_add(1)) < nChunks;) {
            const size_t a = k * CH, b = std::min(nU, a + CH);
            while (landed.load(std::memory_order_acquire) <= (b - 1) / PIECE) std::this_thread::yield();
            size_t off = offBefore[k];
            for (size_t u = a; u < b; ++u) {
                const unsigned r3 = 3u * hR[u], c3 = 3u * hC[u];
                const double* v = hV + u * 9;
                const __m128i base = _mm_set1_epi64x((long long)(((unsigned long long)c3 << 32) | r3));
                const __m128i i01 = _mm_add_epi64(base, o01), i23 = _mm_add_epi64(base, o23), i45 = _mm_add_epi64(base, o45),
                              i67 = _mm_add_epi64(base, o67), i8 = _mm_add_epi64(base, o8);
                const __m128i v01 = _mm_castpd_si128(_mm_loadu_pd(v)), v23 = _mm_castpd_si128(_mm_loadu_pd(v + 2)),
                              v45 = _mm_castpd_si128(_mm_loadu_pd(v + 4)), v67 = _mm_castpd_si128(_mm_loadu_pd(v + 6)),
                              v8 = _mm_castpd_si128(_mm_load_sd(v + 8));
                __m128i* o = reinterpret_cast<__m128i*>(out + u * 9);
                st(o + 0, _mm_unpacklo_epi64(i01, v01)); st(o + 1, _mm_unpackhi_epi64(i01, v01));
                st(o + 2, _mm_unpacklo_epi64(i23, v23)); st(o + 3, _mm_unpackhi_epi64(i23, v23));
                st(o + 4, _mm_unpacklo_epi64(i45, v45)); st(o + 5, _mm_unpackhi_epi64(i45, v45));
                st(o + 6, _mm_unpacklo_epi64(i67, v67)); st(o + 7, _mm_unpackhi_epi64(i67, v67));
                st(o + 8, _mm_unpacklo_epi64(i8, v8));
                if (r3 != c3) {
                    __m128i* m = reinterpret_cast<__m128i*>(out + (nU + off) * 9);
                    const __m128i m01 = _mm_shuffle_epi32(i01, _MM_SHUFFLE(2, 3, 0, 1)), m23 = _mm_shuffle_epi32(i23, _MM_SHUFFLE(2, 3, 0, 1)),
                                  m45 = _mm_shuffle_epi32(i45, _MM_SHUFFLE(2, 3, 0, 1)), m67 = _mm_shuffle_epi32(i67, _MM_SHUFFLE(2, 3, 0, 1)),
                                  m8 = _mm_shuffle_epi32(i8, _MM_SHUFFLE(2, 3, 0, 1));
                    st(m + 0, _mm_unpacklo_epi64(m01, v01)); st(m + 1, _mm_unpackhi_epi64(m01, v01));
                    st(m + 2, _mm_unpacklo_epi64(m23, v23)); st(m + 3, _mm_unpackhi_epi64(m23, v23));
                    st(m + 4, _mm_unpacklo_epi64(m45, v45)); st(m + 5, _mm_unpackhi_epi64(m45, v45));
                    st(m + 6, _mm_unpacklo_epi64(m67, v67)); st(m + 7, _mm_unpackhi_epi64(m67, v67));
                    st(m + 8, _mm_unpacklo_epi64(m8, v8));
                    ++off;
                }
            }
        }
        _mm_sfence();
    };
    HostPool::get().run([&](int tid) {
        if (tid == 0) { // the calling thread publishes the pieces as they land, then joins the work
            for (size_t k = 0; k < nPieces; ++k) {
                if (!failed && cudaEventSynchronize(evs[k]) != cudaSuccess) failed = true;
                landed.store(k + 1, std::memory_order_release);
            }
            c->ctr["deliver_us_copied"] = us_since(tp0);
        }
        worker();
    });
    c->ctr["deliver_us_total"] = us_since(tp0);
    for (auto e : evs) cudaEventDestroy(e);
    if (failed) throw CudaError("device-to-host copy of the merged Hessian blocks failed");
}

// ---- multi-device fan-out (defined behind the C ABI, multidev.h)
static int multi_set_topology(cipc_ctx* ctx, int nV, int nBN, const int32_t* BN, int nBE, const int32_t* BE, int be_stride, int nBT, const int32_t* BT,
    int bt_stride, int nRod, const int32_t codim[2], const uint8_t* dbc, int nNnx, const int32_t* nnxPairs, const double* BNArea, const double* BEArea,
    const double* BTArea);
static int multi_upload(cipc_ctx* ctx, int which, const double* src, int stride_bytes);
static int multi_constraint_set(cipc_ctx* ctx, int elastic, double dHat2, double thickness, int* nC_out);
static int multi_get_constraints(cipc_ctx* ctx, int32_t* cs, double* info, int info_stride_bytes);
static int multi_set_constraints(cipc_ctx* ctx, const int32_t* cs, const double* info, int info_stride_bytes, int nC);
static int multi_barrier_energy(cipc_ctx* ctx, int elastic, double dHat2, const double kappa[3], double thickness, double* E);
static int multi_barrier_gradient(cipc_ctx* ctx, int elastic, double dHat2, const double kappa[3], double thickness, double* g, int stride);
static int multi_barrier_hessian(cipc_ctx* ctx, int elastic, double dHat2, const double kappa[3], double thickness, int projectSPD, int64_t* nTrip, bool merged);
static int multi_get_triplets(cipc_ctx* ctx, cipc_triplet* out);
static int multi_step_size(cipc_ctx* ctx, int elastic, double thickness, double* step);
static int multi_min_dist2(cipc_ctx* ctx, double thickness, double* dist2, double* minDist2);
#define CIPC_MULTI_UNSUPPORTED(ctx) do { if ((ctx) && (ctx)->multi) { (ctx)->err = "not available on a multi-device context"; return CIPC_ERR_UNSUPPORTED; } } while (0)

// ========================================================================================= C ABI
extern "C" {

const char* cipc_version(void) { return "cipc_b200 0.2 (sm_100a)"; }
int cipc_host_parallel_for(size_t n, size_t grain, cipc_range_fn fn, void* user)
{
    if (!fn) return CIPC_ERR_ARG;
    if (grain == 0) grain = 1;
    const size_t items = (n + grain - 1) / grain;
    if (items <= 1) { if (n) fn(0, n, user); return CIPC_OK; }
    HostPool::get().for_each(items, [&](size_t k) { fn(k * grain, std::min(n, (k + 1) * grain), user); });
    return CIPC_OK;
}
int cipc_host_prefault_async(void* p, size_t bytes)
{
    try { Prefault::get().start(p, bytes); } catch (...) { return CIPC_ERR_ARG; }
    return CIPC_OK;
}
uint64_t cipc_hash_bytes(const void* p, size_t n)
{
    const HashSeg sg{(const unsigned char*)p, n};
    return hash_segments(&sg, 1, 0xcbf29ce484222325ULL);
}
int64_t cipc_kernel_launches(void) { return g_launches; }

int cipc_create(int device, int rank, int world, cipc_ctx** out)
{
    if (!out || world < 1 || rank < 0 || rank >= world) return CIPC_ERR_ARG;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0 || device >= ndev) {
        fprintf(stderr, "cipc_b200: no usable CUDA device (requested %d of %d); this library has no CPU path\n", device, ndev);
        return CIPC_ERR_CUDA;
    }
    std::unique_ptr<cipc_ctx> c(new cipc_ctx());
    c->dev = device; c->rank = rank; c->world = world;
    try {
        CIPC_CUDA(cudaSetDevice(device));
        CIPC_CUDA(cudaStreamCreateWithFlags(&c->ownSt, cudaStreamNonBlocking));
        c->st = c->ownSt;
        CIPC_CUDA(cudaStreamCreateWithFlags(&c->sideSt, cudaStreamNonBlocking));
        CIPC_CUDA(cudaEventCreateWithFlags(&c->evFork, cudaEventDisableTiming));
        CIPC_CUDA(cudaEventCreateWithFlags(&c->evJoin, cudaEventDisableTiming));
        c->partial.reserve(RED_GRID, c->st);
        c->scal.reserve(16, c->st);
        c->bbox.reserve(6, c->st);
        c->counters.reserve(16, c->st);
        c->errFlag.reserve(1, c->st);
        c->nTasksDev.reserve(1, c->st);
        CIPC_CUDA(cudaMemsetAsync(c->scal.p, 0, 16 * sizeof(double), c->st));
        CIPC_CUDA(cudaFuncSetAttribute(k_barrier_hessian, cudaFuncAttributeMaxDynamicSharedMemorySize, DENSE_SMEM));
        CIPC_CUDA(cudaFuncSetAttribute((k_hessian_fused<0, false>), cudaFuncAttributeMaxDynamicSharedMemorySize, FusedShape<0>::SMEM));
        CIPC_CUDA(cudaFuncSetAttribute((k_hessian_fused<1, false>), cudaFuncAttributeMaxDynamicSharedMemorySize, FusedShape<1>::SMEM));
        CIPC_CUDA(cudaFuncSetAttribute((k_hessian_fused<2, false>), cudaFuncAttributeMaxDynamicSharedMemorySize, FusedShape<2>::SMEM));
        CIPC_CUDA(cudaFuncSetAttribute((k_hessian_fused<0, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, FusedShape<0>::SMEM));
        CIPC_CUDA(cudaFuncSetAttribute((k_hessian_fused<1, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, FusedShape<1>::SMEM));
        CIPC_CUDA(cudaFuncSetAttribute((k_hessian_fused<2, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, FusedShape<2>::SMEM));
        CIPC_CUDA(cudaFuncSetAttribute((k_friction_fused<0, false>), cudaFuncAttributeMaxDynamicSharedMemorySize, FrFusedShape<0>::SMEM));
        CIPC_CUDA(cudaFuncSetAttribute((k_friction_fused<1, false>), cudaFuncAttributeMaxDynamicSharedMemorySize, FrFusedShape<1>::SMEM));
        CIPC_CUDA(cudaFuncSetAttribute((k_friction_fused<2, false>), cudaFuncAttributeMaxDynamicSharedMemorySize, FrFusedShape<2>::SMEM));
        CIPC_CUDA(cudaFuncSetAttribute((k_friction_fused<0, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, FrFusedShape<0>::SMEM));
        CIPC_CUDA(cudaFuncSetAttribute((k_friction_fused<1, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, FrFusedShape<1>::SMEM));
        CIPC_CUDA(cudaFuncSetAttribute((k_friction_fused<2, true>), cudaFuncAttributeMaxDynamicSharedMemorySize, FrFusedShape<2>::SMEM));
        CIPC_CUDA(cudaStreamSynchronize(c->st));
    }
    catch (const std::exception& e) {
        fprintf(stderr, "cipc_b200: %s\n", e.what());
        return CIPC_ERR_CUDA;
    }
    *out = c.release();
    return CIPC_OK;
}
void cipc_destroy(cipc_ctx* ctx)
{
    if (!ctx) return;
    if (ctx->multi) {
        ctx->multi->thr.clear(); // joins the helper threads
        for (cipc_ctx* s : ctx->multi->sub) cipc_destroy(s);
        cudaSetDevice(ctx->dev);
        ctx->multi.reset();
        delete ctx;
        return;
    }
    cudaSetDevice(ctx->dev);
    cudaStreamSynchronize(ctx->st);
    for (auto e : ctx->evPool) cudaEventDestroy(e);
    for (auto e : ctx->userEv) if (e) cudaEventDestroy(e);
    cudaStreamDestroy(ctx->ownSt);
    if (ctx->sideSt) { cudaStreamSynchronize(ctx->sideSt); cudaStreamDestroy(ctx->sideSt); }
    if (ctx->evFork) cudaEventDestroy(ctx->evFork);
    if (ctx->evJoin) cudaEventDestroy(ctx->evJoin);
    delete ctx;
}
const char* cipc_last_error(cipc_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }
int cipc_sync(cipc_ctx* ctx)
{
    if (ctx && ctx->multi) { int st = CIPC_OK; for (cipc_ctx* s : ctx->multi->sub) { const int r = cipc_sync(s); if (r) st = r; } return st; }
    return guarded(ctx, [&]() { CIPC_CUDA(cudaStreamSynchronize(ctx->st)); return (int)CIPC_OK; });
}

int cipc_set_stream(cipc_ctx* ctx, void* stream)
{
    CIPC_MULTI_UNSUPPORTED(ctx);
    return guarded(ctx, [&]() {
        CIPC_CUDA(cudaStreamSynchronize(ctx->st));
        ctx->st = stream ? (cudaStream_t)stream : ctx->ownSt;
        return (int)CIPC_OK;
    });
}
int cipc_set_timing(cipc_ctx* ctx, int on)
{
    if (!ctx) return CIPC_ERR_ARG;
    ctx->timing = on != 0;
    if (ctx->multi) for (cipc_ctx* s : ctx->multi->sub) s->timing = on != 0;
    return CIPC_OK;
}
int cipc_event_record(cipc_ctx* ctx, int slot)
{
    CIPC_MULTI_UNSUPPORTED(ctx);
    return guarded(ctx, [&]() {
        if (slot < 0 || slot >= 64) return (int)CIPC_ERR_ARG;
        if (!ctx->userEv[slot]) CIPC_CUDA(cudaEventCreate(&ctx->userEv[slot]));
        CIPC_CUDA(cudaEventRecord(ctx->userEv[slot], ctx->st));
        return (int)CIPC_OK;
    });
}
double cipc_event_elapsed_ms(cipc_ctx* ctx, int a, int b)
{
    if (!ctx || a < 0 || b < 0 || a >= 64 || b >= 64 || !ctx->userEv[a] || !ctx->userEv[b]) return -1;
    cudaSetDevice(ctx->dev);
    if (cudaEventSynchronize(ctx->userEv[b]) != cudaSuccess) return -1;
    float ms = 0;
    if (cudaEventElapsedTime(&ms, ctx->userEv[a], ctx->userEv[b]) != cudaSuccess) return -1;
    return ms;
}

int cipc_set_topology(cipc_ctx* ctx, int nV, int nBN, const int32_t* BN, int nBE, const int32_t* BE, int be_stride, int nBT,
    const int32_t* BT, int bt_stride, int nRod, const int32_t codim[2], const uint8_t* dbc, int nNnx, const int32_t* nnxPairs,
    const double* BNArea, const double* BEArea, const double* BTArea)
{
    if (ctx && ctx->multi) return multi_set_topology(ctx, nV, nBN, BN, nBE, BE, be_stride, nBT, BT, bt_stride, nRod, codim, dbc, nNnx, nnxPairs, BNArea, BEArea, BTArea);
    return guarded(ctx, [&]() {
        if (nV <= 0 || nBN < 0 || nBE < 0 || nBT < 0 || (be_stride != 2 && be_stride != 4) || (bt_stride != 3 && bt_stride != 4) || !dbc)
            return (int)CIPC_ERR_ARG;
        cipc_ctx* c = ctx;
        u64 h = 0xcbf29ce484222325ULL;
        const int hdr[8] = {nV, nBN, nBE, nBT, nRod, codim[0], codim[1], nNnx};
        h = fnv(hdr, sizeof(hdr), h);
        {
            const HashSeg segs[4] = {{(const unsigned char*)BN, (size_t)nBN * 4}, {(const unsigned char*)BE, (size_t)nBE * be_stride * 4},
                {(const unsigned char*)BT, (size_t)nBT * bt_stride * 4}, {(const unsigned char*)dbc, (size_t)nV}};
            h = hash_segments(segs, 4, h);
        }
        if (nNnx) h = fnv(nnxPairs, (size_t)nNnx * 8, h);
        if (BNArea) h = fnv(BNArea, (size_t)nBN * 8, h ^ 1);
        if (BEArea) h = fnv(BEArea, (size_t)nBE * 8, h ^ 2);
        if (BTArea) h = fnv(BTArea, (size_t)nBT * 8, h ^ 3);
        if (h == c->topoHash && c->T.nV == nV) return (int)CIPC_OK; // unchanged: keep the resident copy
        Topo& T = c->T;
        T.nV = nV; T.nBN = nBN; T.nBE = nBE; T.nBT = nBT; T.nRod = nRod; T.codim0 = codim[0]; T.codim1 = codim[1];
        c->BN.reserve(std::max(nBN, 1), c->st); c->BE.reserve(std::max(nBE, 1), c->st); c->BT.reserve(std::max(nBT, 1), c->st);
        c->flags.reserve(nV, c->st); c->v2sv.reserve(nV, c->st);
        CIPC_CUDA(cudaMemcpyAsync(c->BN.p, BN, (size_t)nBN * 4, cudaMemcpyHostToDevice, c->st));
        const size_t need_i = std::max((size_t)nBE * be_stride, (size_t)nBT * bt_stride);
        c->stageI.reserve(std::max<size_t>(need_i, 1), c->st);
        if (nBE) {
            CIPC_CUDA(cudaMemcpyAsync(c->stageI.p, BE, (size_t)nBE * be_stride * 4, cudaMemcpyHostToDevice, c->st));
            CIPC_LAUNCH(k_pack_edges, div_up(nBE, TB), TB, 0, c->st, c->stageI.p, be_stride, nBE, c->BE.p);
        }
        if (nBT) {
            CIPC_CUDA(cudaMemcpyAsync(c->stageI.p, BT, (size_t)nBT * bt_stride * 4, cudaMemcpyHostToDevice, c->st));
            CIPC_LAUNCH(k_pack_tris, div_up(nBT, TB), TB, 0, c->st, c->stageI.p, bt_stride, nBT, c->BT.p);
        }
        // flags and vertex -> boundary slot map are small host loops
        std::vector<uint8_t> fl(dbc, dbc + nV);
        for (int v = 0; v < nV; ++v) fl[v] = fl[v] ? 1 : 0;
        std::vector<u64> nn(nNnx);
        for (int i = 0; i < nNnx; ++i) {
            const int k = nnxPairs[2 * i], m = nnxPairs[2 * i + 1];
            if (k < 0 || k >= nV) return (int)CIPC_ERR_ARG;
            fl[k] |= 2;
            nn[i] = ((u64)(u32)k << 32) | (u32)m;
        }
        std::sort(nn.begin(), nn.end());
        std::vector<int> v2(nV, 0);
        for (int i = 0; i < nBN; ++i) {
            if (BN[i] < 0 || BN[i] >= nV) return (int)CIPC_ERR_ARG;
            v2[BN[i]] = i;
        }
        CIPC_CUDA(cudaMemcpyAsync(c->flags.p, fl.data(), nV, cudaMemcpyHostToDevice, c->st));
        CIPC_CUDA(cudaMemcpyAsync(c->v2sv.p, v2.data(), (size_t)nV * 4, cudaMemcpyHostToDevice, c->st));
        c->nnx.reserve(std::max(nNnx, 1), c->st);
        if (nNnx) CIPC_CUDA(cudaMemcpyAsync(c->nnx.p, nn.data(), (size_t)nNnx * 8, cudaMemcpyHostToDevice, c->st));
        if (BNArea && BEArea && BTArea) {
            c->BNArea.reserve(std::max(nBN, 1), c->st); c->BEArea.reserve(std::max(nBE, 1), c->st); c->BTArea.reserve(std::max(nBT, 1), c->st);
            CIPC_CUDA(cudaMemcpyAsync(c->BEArea.p, BEArea, (size_t)nBE * 8, cudaMemcpyHostToDevice, c->st));
            CIPC_CUDA(cudaMemcpyAsync(c->BTArea.p, BTArea, (size_t)nBT * 8, cudaMemcpyHostToDevice, c->st));
        }
        CIPC_CUDA(cudaStreamSynchronize(c->st)); // host vectors above go out of scope
        T.BN = c->BN.p; T.BE = c->BE.p; T.BT = c->BT.p; T.flags = c->flags.p; T.v2sv = c->v2sv.p; T.nnx = c->nnx.p; T.nNnx = nNnx;
        c->topoHash = h;
        c->haveX = c->haveX0 = c->haveP = c->haveXn = c->haveXprev = false;
        c->tagX = c->tagX0 = c->tagP = c->tagXn = 0;
        c->slabCache[0].valid = c->slabCache[1].valid = false;
        c->dedupLastRaw = c->dedupLastUnique = 0; // a new scene: no duplicate-ratio history for the merge table
        ++c->xVersion;
        c->nC = 0;
        c->nF = 0;
        return (int)CIPC_OK;
    });
}
int cipc_set_positions(cipc_ctx* ctx, const double* X, int stride_bytes)
{
    if (ctx && ctx->multi) return multi_upload(ctx, 0, X, stride_bytes);
    return guarded(ctx, [&]() {
        need(ctx->T.nV > 0, "topology not set");
        if (upload_vec3(ctx, ctx->X, X, stride_bytes, &ctx->tagX)) ++ctx->xVersion;
        ctx->haveX = true;
        return (int)CIPC_OK;
    });
}
int cipc_set_rest_positions(cipc_ctx* ctx, const double* X0, int stride_bytes)
{
    if (ctx && ctx->multi) return multi_upload(ctx, 1, X0, stride_bytes);
    return guarded(ctx, [&]() {
        need(ctx->T.nV > 0, "topology not set");
        if (upload_vec3(ctx, ctx->X0, X0, stride_bytes, &ctx->tagX0) || !ctx->haveX0) {
            ctx->restLen2.reserve(std::max(ctx->T.nBE, 1), ctx->st);
            if (ctx->T.nBE) CIPC_LAUNCH(k_rest_len2, div_up(ctx->T.nBE, TB), TB, 0, ctx->st, ctx->X0.p, ctx->BE.p, ctx->T.nBE, ctx->restLen2.p);
        }
        ctx->haveX0 = true;
        return (int)CIPC_OK;
    });
}
int cipc_set_search_dir(cipc_ctx* ctx, const double* p)
{
    if (ctx && ctx->multi) return multi_upload(ctx, 2, p, 24);
    return guarded(ctx, [&]() {
        need(ctx->T.nV > 0, "topology not set");
        upload_vec3(ctx, ctx->P, p, 24, &ctx->tagP);
        ctx->haveP = true;
        return (int)CIPC_OK;
    });
}

int cipc_constraint_set(cipc_ctx* ctx, int elastic, double dHat2, double thickness, int* nC_out)
{
    if (ctx && ctx->multi) return multi_constraint_set(ctx, elastic, dHat2, thickness, nC_out);
    return guarded(ctx, [&]() {
        cipc_ctx* c = ctx;
        need(c->haveX && c->haveX0, "positions / rest positions not set");
        if (elastic) return (int)CIPC_ERR_UNSUPPORTED; // every FEMShell scene runs initialize_OIPC (elasticIPC=false)
        c->begin_call();
        const Topo& T = c->T;
        const double dHat = std::sqrt(dHat2) + thickness; // IPC.h:53-54
        const double dHat2o = dHat * dHat;
        HashInfo H;
        {
            cipc_ctx::Scope sc(c, "ccs_hash_build");
            edge_len_begin(c); // rides on the bounding-box round trip
            double mn[3], mx[3];
            bbox_to_host(c, nullptr, 0.0, mn, mx);
            // Voxel size of the constraint-set grid (our own choice: any grid gives the same set).  The reference uses the
            // mean edge length, which puts hundreds of particles into one voxel of a granular pile whose only edges belong
            // to a coarse obstacle mesh.  Here: the mean extent over ALL primitives (points count as 0; edges and triangles
            // ~ one edge length) plus the activation distance every box is inflated by -- for a cloth mesh that is the
            // reference's choice within a few percent, for point clouds it drops to a few dHat.
            const double nPrim = (double)T.nBN + T.nBE + T.nBT;
            double voxelSize = (T.nBE ? edge_len_end(c) * ((double)T.nBE + T.nBT) / nPrim : 0.0) + dHat;
            voxelSize = std::max(voxelSize, 2.0 * dHat);
            double mag = 0;
            for (int d = 0; d < 3; ++d) mag = std::max(mag, std::max(std::fabs(mn[d]), std::fabs(mx[d])));
            // inflation radius: half the activation distance plus a rounding guard, so that any pair
            // that passes the reference's AABB gap test (gap <= dHat) shares a voxel
            const double r = 0.5 * dHat * (1.0 + 1e-9) + 8.0 * 2.220446049250313e-16 * mag;
            for (int d = 0; d < 3; ++d) { mn[d] -= r; mx[d] += r; }
            if (!size_grid(mn, mx, voxelSize, H.G, true)) return (int)CIPC_ERR_GRID;
            const int nP = T.nBN + T.nBE + T.nBT;
            c->boxLo.reserve(nP, c->st); c->boxHi.reserve(nP, c->st); c->cnt.reserve(nP, c->st); c->fine.reserve(nP, c->st);
            CIPC_LAUNCH(k_boxes_ccs, div_up(nP, TB), TB, 0, c->st, T, c->X.p, H.G, r, c->boxLo.p, c->boxHi.p, c->fine.p, c->cnt.p);
            build_cell_lists(c, H, 0);
        }
        u32 counts[4];
        {
            cipc_ctx::Scope sc(c, "ccs_pairs");
            run_pairs<false>(c, H, dHat, counts);
        }
        c->ctr["candidates_pt"] = counts[0]; c->ctr["candidates_ee"] = counts[1];
        c->ctr["candidates_pe"] = counts[2]; c->ctr["candidates_pp"] = counts[3];
        const size_t total = (size_t)counts[0] + counts[1] + counts[2] + counts[3];
        c->cs.reserve(total + 1, c->st);
        c->raw.reserve(total + 1, c->st);
        u32 hc[2];
        {
            cipc_ctx::Scope sc(c, "ccs_narrow");
            CIPC_CUDA(cudaMemsetAsync(c->counters.p + 8, 0, 4 * sizeof(u32), c->st));
            NarrowOut out{c->cs.p, c->raw.p, c->counters.p + 8};
            {
                cipc_ctx::Scope sp(c, "ccs_narrow_pt"); // point queries: PT + the codimensional PE / PP candidates (the reference's _PT scope)
                if (counts[0]) CIPC_LAUNCH(k_narrow_pt, div_up(counts[0], NARROW_BT), NARROW_BT, 0, c->st, T, c->X.p, c->cand[0].p, counts[0], dHat2o, out);
                if (counts[2]) CIPC_LAUNCH(k_narrow_pe, div_up(counts[2], NARROW_BT), NARROW_BT, 0, c->st, T, c->X.p, c->cand[2].p, counts[2], dHat2o, out);
                if (counts[3]) CIPC_LAUNCH(k_narrow_pp, div_up(counts[3], NARROW_BT), NARROW_BT, 0, c->st, T, c->X.p, c->cand[3].p, counts[3], dHat2o, out);
            }
            {
                cipc_ctx::Scope se(c, "ccs_narrow_ee");
                if (counts[1]) CIPC_LAUNCH(k_narrow_ee, div_up(counts[1], NARROW_BT), NARROW_BT, 0, c->st, T, c->X.p, c->restLen2.p, c->cand[1].p, counts[1], dHat2o, out);
            }
            CIPC_CUDA(cudaMemcpyAsync(hc, c->counters.p + 8, 2 * sizeof(u32), cudaMemcpyDeviceToHost, c->st));
            CIPC_CUDA(cudaStreamSynchronize(c->st));
        }
        u32 nC = hc[0];
        c->nPassLast = hc[0];
        {
            cipc_ctx::Scope sc(c, "ccs_merge");
            const u32 nRaw = hc[1];
            if (nRaw) {
                // Table size: at most ~55 % full for the number of UNIQUE stencils expected from the previous call's duplicate ratio
                // (a cloth stack has 2-3 raw records per unique stencil, a particle pile one), never smaller than the record
                // count.  A table that is larger than needed costs the insert pass its L2 locality (0.31 vs 0.22 ms at 1M
                // triangles), a fuller one long probe chains (particles: 0.12 vs 0.09 ms); the result does not depend on it.
                const double uniqRatio = c->dedupLastRaw ? std::min(1.0, (double)c->dedupLastUnique / (double)c->dedupLastRaw) : 1.0;
                const u64 want = std::max<u64>((u64)(1.82 * uniqRatio * nRaw), (u64)nRaw + nRaw / 16 + 1024);
                u32 nSlots = 1024;
                while (nSlots < want) nSlots <<= 1;
                c->slots.reserve(nSlots, c->st); c->slotCnt.reserve(nSlots, c->st);
                CIPC_CUDA(cudaMemsetAsync(c->slots.p, 0xff, (size_t)nSlots * 4, c->st));
                CIPC_CUDA(cudaMemsetAsync(c->slotCnt.p, 0, (size_t)nSlots * 4, c->st));
                c->dedupOwn.reserve(nRaw, c->st);
                CIPC_LAUNCH(k_dedup_insert, div_up(nRaw, TB), TB, 0, c->st, c->raw.p, nRaw, c->slots.p, c->slotCnt.p, nSlots - 1, c->dedupOwn.p);
                // the unique PP/PE stencils are appended behind the pass-through ones (counter keeps running)
                CIPC_LAUNCH(k_dedup_emit_records, div_up(nRaw, 256), 256, 0, c->st, c->raw.p, c->dedupOwn.p, c->slotCnt.p, nRaw, c->cs.p, c->counters.p + 8);
                CIPC_CUDA(cudaMemcpyAsync(&nC, c->counters.p + 8, sizeof(u32), cudaMemcpyDeviceToHost, c->st));
                CIPC_CUDA(cudaStreamSynchronize(c->st));
                c->dedupLastRaw = nRaw; c->dedupLastUnique = nC - hc[0];
            }
            c->info.reserve((size_t)nC + 1, c->st);
            if (nC) CIPC_LAUNCH(k_fill_info, div_up(nC, TB), TB, 0, c->st, c->info.p, nC, 1.0, dHat2o);
            c->infoUniform = true; c->infoW = 1.0; c->infoD = dHat2o;
        }
        c->nC = nC;
        c->ctr["constraints"] = nC;
        if (nC_out) *nC_out = (int)nC;
        return (int)CIPC_OK;
    });
}
int cipc_get_constraints_strided(cipc_ctx* ctx, int32_t* cs, double* info, int info_stride_bytes)
{
    if (ctx && ctx->multi) return multi_get_constraints(ctx, cs, info, info_stride_bytes);
    return guarded(ctx, [&]() {
        cipc_ctx* c = ctx;
        if (info && (info_stride_bytes < 16 || info_stride_bytes % 8)) return (int)CIPC_ERR_ARG;
        if (c->nC == 0) return (int)CIPC_OK;
        const size_t n = c->nC;
        if (cs) staged_d2h(c->ring, c->st, cs, c->cs.p, n * 16);
        if (info) {
            const size_t sd = (size_t)info_stride_bytes / 8;
            const size_t CH = 1 << 15;
            if (c->infoUniform) { // (w, dHat2) is the same for every constraint of a non-elastic set: nothing crosses PCIe
                const double w = c->infoW, d = c->infoD;
                HostPool::get().for_each((n + CH - 1) / CH, [&](size_t k) {
                    if (sd == 4 && ((uintptr_t)info & 15) == 0) { // the reference's VECTOR<T,2>: T data[4], the unused half zero like its constructor leaves it
                        const __m128d lo = _mm_set_pd(d, w), z = _mm_setzero_pd();
                        for (size_t i = k * CH, e = std::min(n, i + CH); i < e; ++i) { _mm_stream_pd(info + 4 * i, lo); _mm_stream_pd(info + 4 * i + 2, z); }
                        _mm_sfence();
                    }
                    else if (sd == 4)
                        for (size_t i = k * CH, e = std::min(n, i + CH); i < e; ++i) { info[4 * i] = w; info[4 * i + 1] = d; info[4 * i + 2] = 0; info[4 * i + 3] = 0; }
                    else
                        for (size_t i = k * CH, e = std::min(n, i + CH); i < e; ++i) { info[i * sd] = w; info[i * sd + 1] = d; }
                });
            }
            else if (sd == 2) staged_d2h(c->ring, c->st, info, c->info.p, n * 16);
            else {
                const double* h = (const double*)c->pin.reserve(n * 16);
                CIPC_CUDA(cudaMemcpyAsync((void*)h, c->info.p, n * 16, cudaMemcpyDeviceToHost, c->st));
                CIPC_CUDA(cudaStreamSynchronize(c->st));
                HostPool::get().for_each((n + CH - 1) / CH, [&](size_t k) {
                    for (size_t i = k * CH, e = std::min(n, i + CH); i < e; ++i) { info[i * sd] = h[2 * i]; info[i * sd + 1] = h[2 * i + 1]; }
                });
            }
        }
        return (int)CIPC_OK;
    });
}
int cipc_get_constraints(cipc_ctx* ctx, int32_t* cs, double* info) { return cipc_get_constraints_strided(ctx, cs, info, 16); }
int cipc_set_constraints_strided(cipc_ctx* ctx, const int32_t* cs, const double* info, int info_stride_bytes, int nC)
{
    if (ctx && ctx->multi) return multi_set_constraints(ctx, cs, info, info_stride_bytes, nC);
    return guarded(ctx, [&]() {
        cipc_ctx* c = ctx;
        if (nC < 0 || info_stride_bytes < 16 || info_stride_bytes % 8) return (int)CIPC_ERR_ARG;
        const size_t n = (size_t)nC;
        c->cs.reserve(n + 1, c->st); c->info.reserve(n + 1, c->st);
        c->infoUniform = false;
        if (n) {
            staged_h2d(c->ring, c->st, c->cs.p, cs, n * 16);
            if (info_stride_bytes == 16) staged_h2d(c->ring, c->st, c->info.p, info, n * 16);
            else {
                const size_t sd = (size_t)info_stride_bytes / 8, CH = 1 << 15;
                double* h = (double*)c->pin.reserve(n * 16);
                HostPool::get().for_each((n + CH - 1) / CH, [&](size_t k) {
                    for (size_t i = k * CH, e = std::min(n, i + CH); i < e; ++i) { h[2 * i] = info[i * sd]; h[2 * i + 1] = info[i * sd + 1]; }
                });
                CIPC_CUDA(cudaMemcpyAsync(c->info.p, h, n * 16, cudaMemcpyHostToDevice, c->st));
            }
            CIPC_CUDA(cudaStreamSynchronize(c->st)); // the caller's arrays and the staging buffers may be reused
        }
        c->nC = (u32)nC;
        return (int)CIPC_OK;
    });
}
int cipc_set_constraints(cipc_ctx* ctx, const int32_t* cs, const double* info, int nC) { return cipc_set_constraints_strided(ctx, cs, info, 16, nC); }

int cipc_barrier_energy_dev(cipc_ctx* ctx, int elastic, double dHat2, const double kappa[3], double thickness)
{
    CIPC_MULTI_UNSUPPORTED(ctx);
    return guarded(ctx, [&]() { ctx->begin_call(); return do_barrier_energy(ctx, elastic, dHat2, kappa, thickness); });
}
int cipc_barrier_energy(cipc_ctx* ctx, int elastic, double dHat2, const double kappa[3], double thickness, double* E)
{
    if (ctx && ctx->multi) return multi_barrier_energy(ctx, elastic, dHat2, kappa, thickness, E);
    return guarded(ctx, [&]() {
        ctx->begin_call();
        int r = do_barrier_energy(ctx, elastic, dHat2, kappa, thickness);
        if (r) return r;
        double v;
        CIPC_CUDA(cudaMemcpyAsync(&v, ctx->scal.p, sizeof(double), cudaMemcpyDeviceToHost, ctx->st));
        r = fetch_err(ctx);
        if (r) return r;
        *E += v;
        return (int)CIPC_OK;
    });
}
int cipc_barrier_gradient_dev(cipc_ctx* ctx, int elastic, double dHat2, const double kappa[3], double thickness)
{
    CIPC_MULTI_UNSUPPORTED(ctx);
    return guarded(ctx, [&]() { ctx->begin_call(); return do_barrier_gradient(ctx, elastic, dHat2, kappa, thickness); });
}
int cipc_barrier_gradient(cipc_ctx* ctx, int elastic, double dHat2, const double kappa[3], double thickness, double* g, int stride)
{
    if (ctx && ctx->multi) return multi_barrier_gradient(ctx, elastic, dHat2, kappa, thickness, g, stride);
    return guarded(ctx, [&]() {
        cipc_ctx* c = ctx;
        if (stride < 24 || stride % 8) return (int)CIPC_ERR_ARG;
        c->begin_call();
        int r = do_barrier_gradient(c, elastic, dHat2, kappa, thickness);
        if (r) return r;
        const size_t n = (size_t)c->T.nV;
        double* h = (double*)c->pin.reserve(n * 24);
        CIPC_CUDA(cudaMemcpyAsync(h, c->g.p, n * 24, cudaMemcpyDeviceToHost, c->st));
        CIPC_CUDA(cudaStreamSynchronize(c->st));
        const size_t sd = stride / 8, CH = 1 << 14;
        HostPool::get().for_each((n + CH - 1) / CH, [&](size_t k) { // nodeAttr.g += (IPC.h:1034-1042)
            for (size_t v = k * CH, e = std::min(n, v + CH); v < e; ++v) { g[v * sd] += h[3 * v]; g[v * sd + 1] += h[3 * v + 1]; g[v * sd + 2] += h[3 * v + 2]; }
        });
        return (int)CIPC_OK;
    });
}
enum HessOut { H_HOST = 0, H_DEV = 1, H_BLK = 2 };
// H_HOST: factors for the host-side expansion (cipc_get_triplets ships them over PCIe); H_DEV: the (row, col, value) stream in
// HBM (fused factor + expansion); H_BLK: upper 3x3 blocks + device-side merge into the unique blocks (merge.cuh)
static int barrier_hessian_impl(cipc_ctx* ctx, int elastic, double dHat2, const double kappa[3], double thickness, int projectSPD,
    int64_t* nTrip, HessOut mode, bool withGradient = false)
{
    return guarded(ctx, [&]() {
        cipc_ctx* c = ctx;
        need(c->haveX, "positions not set");
        c->begin_call();
        const BarrierParams bp = make_bp(elastic, dHat2, kappa, thickness);
        const bool blk = mode == H_BLK;
        c->nTrip = 0;
        c->factorValid = false;
        c->mergedValid = false;
        c->hessStencils = c->cs.p; c->hessN = c->nC;
        if (c->nC) {
            cipc_ctx::Scope sc(c, "barrier_H");
            c->tripOff.reserve(c->nC, c->st);
            if (blk) CIPC_LAUNCH(k_blk_sizes, div_up(c->nC, TB), TB, 0, c->st, c->cs.p, c->nC, c->tripOff.p);
            else CIPC_LAUNCH(k_block_sizes, div_up(c->nC, TB), TB, 0, c->st, c->cs.p, c->nC, c->tripOff.p);
            device_excl_scan(c->tripOff.p, c->tripOff.p, c->nC, c->scanwk, c->st);
            u32 tot, nk[4];
            CIPC_CUDA(cudaMemcpyAsync(&tot, c->scanwk.total.p, sizeof(u32), cudaMemcpyDeviceToHost, c->st));
            for (int k = 0; k < 4; ++k) c->clsIdx[k].reserve(c->nC, c->st);
            CIPC_CUDA(cudaMemsetAsync(c->counters.p + 12, 0, 4 * sizeof(u32), c->st));
            CIPC_LAUNCH(k_classify, div_up(c->nC, TB), TB, 0, c->st, c->cs.p, c->nC, c->clsIdx[0].p, c->clsIdx[1].p, c->clsIdx[2].p,
                c->clsIdx[3].p, c->counters.p + 12);
            CIPC_CUDA(cudaMemcpyAsync(nk, c->counters.p + 12, 4 * sizeof(u32), cudaMemcpyDeviceToHost, c->st));
            CIPC_CUDA(cudaStreamSynchronize(c->st)); // one host round trip for the triplet / block total and the class counts
            void* outp;
            if (blk) {
                c->blkVal.reserve((size_t)tot * 9 + 16, c->st);
                c->nBlkLast = tot;
                outp = c->blkVal.p;
            }
            else {
                c->nTrip = (int64_t)tot * 9;
                c->trip.reserve((size_t)c->nTrip, c->st);
                outp = c->trip.p;
            }
            cipc_triplet* outT = reinterpret_cast<cipc_triplet*>(outp);
            const int bm = blk ? 1 : 0;
            const char* dense = getenv("CIPC_HESSIAN_DENSE"); // cross-check switch: force the dense eigen path for every stencil
            if (dense && dense[0] == '1') {
                cipc_ctx::Scope sk(c, "k_barrier_hessian");
                CIPC_LAUNCH(k_barrier_hessian, div_up(c->nC, DENSE_BD), DENSE_BD, DENSE_SMEM, c->st, c->X.p, c->X0.p, c->cs.p, c->info.p, c->tripOff.p,
                    (const u32*)nullptr, c->nC, (const u32*)nullptr, bp, projectSPD, outT, bm);
            }
            else {
                if (projectSPD && mode != H_HOST) {
                    // device-resident output: fused factor + expansion, the factors never leave the SM
                    u32* dl = c->clsIdx[3].p; u32* dn = c->counters.p + 15;
                    double* gOut = nullptr;
                    if (withGradient) { // the barrier gradient of every stencil rides on the same pass
                        c->g.reserve((size_t)3 * c->T.nV + 8, c->st);
                        CIPC_CUDA(cudaMemsetAsync(c->g.p, 0, (size_t)3 * c->T.nV * sizeof(double), c->st));
                        // g[3 nV] = the energy of the last cipc_barrier_energy_dev: one all-reduce(sum) covers gradient and energy
                        CIPC_CUDA(cudaMemcpyAsync(c->g.p + (size_t)3 * c->T.nV, c->scal.p, sizeof(double), cudaMemcpyDeviceToDevice, c->st));
                        gOut = c->g.p;
                        if (nk[3]) CIPC_LAUNCH(k_barrier_gradient, div_up(nk[3], 128), 128, 0, c->st, c->X.p, c->X0.p, c->cs.p, c->info.p, nk[3], bp, c->g.p,
                            (const u32*)c->clsIdx[3].p); // mollified stencils (the first nk[3] entries of the dense list)
                    }
#define CIPC_FUSED(CLS, B) CIPC_LAUNCH((k_hessian_fused<CLS, B>), div_up(nk[CLS], FUSED_BD), FUSED_BD, FusedShape<CLS>::SMEM, c->st, c->X.p, c->cs.p, \
    c->info.p, c->tripOff.p, c->clsIdx[CLS].p, nk[CLS], bp, outp, dl, dn, gOut)
                    if (nk[3]) {
                        // mollified stencils (the first nk[3] entries of the dense list, known now): a thread needs ~0.1 ms for one
                        // dense block, so even a handful would be a serial tail behind the fused kernels -- they run beside them
                        CIPC_CUDA(cudaEventRecord(c->evFork, c->st));
                        CIPC_CUDA(cudaStreamWaitEvent(c->sideSt, c->evFork, 0));
                        CIPC_LAUNCH(k_barrier_hessian, std::min(1184u, div_up(nk[3], DENSE_BD)), DENSE_BD, DENSE_SMEM, c->sideSt, c->X.p, c->X0.p, c->cs.p,
                            c->info.p, c->tripOff.p, c->clsIdx[3].p, nk[3], (const u32*)nullptr, bp, projectSPD, outT, bm, 0u);
                        CIPC_CUDA(cudaEventRecord(c->evJoin, c->sideSt));
                    }
                    {
                        cipc_ctx::Scope sk(c, "k_hessian_fused0"); // the longest launch of the stage: PT/EE blocks
                        if (nk[0]) { if (blk) CIPC_FUSED(0, true); else CIPC_FUSED(0, false); }
                    }
                    {
                        cipc_ctx::Scope sk(c, "k_hessian_fused12");
                        if (nk[1]) { if (blk) CIPC_FUSED(1, true); else CIPC_FUSED(1, false); }
                        if (nk[2]) { if (blk) CIPC_FUSED(2, true); else CIPC_FUSED(2, false); }
                    }
#undef CIPC_FUSED
                    for (int k = 0; k < 4; ++k) c->nk[k] = nk[k];
                    // the stencils the fused kernels rejected (appended behind the mollified ones, list length read on the device)
                    CIPC_LAUNCH(k_barrier_hessian, 148, DENSE_BD, DENSE_SMEM, c->st, c->X.p, c->X0.p, c->cs.p, c->info.p, c->tripOff.p,
                        c->clsIdx[3].p, 0u, (const u32*)dn, bp, projectSPD, outT, bm, nk[3]);
                    if (nk[3]) CIPC_CUDA(cudaStreamWaitEvent(c->st, c->evJoin, 0));
                }
                else if (projectSPD) {
                    // (A) factor, (B) expand; stencils the factor kernels reject are appended to the dense list
                    const size_t ydoubles = (size_t)nk[0] * 36 + (size_t)nk[1] * 18 + (size_t)nk[2] * 6;
                    c->Y.reserve(ydoubles + 4, c->st);
                    c->yhdr.reserve((size_t)nk[0] + nk[1] + nk[2] + 1, c->st);
                    double* Y0 = c->Y.p; double* Y1 = Y0 + (size_t)nk[0] * 36; double* Y2 = Y1 + (size_t)nk[1] * 18;
                    YHdr* h0 = (YHdr*)c->yhdr.p; YHdr* h1 = h0 + nk[0]; YHdr* h2 = h1 + nk[1];
                    u32* dl = c->clsIdx[3].p; u32* dn = c->counters.p + 15; // counters[15] = length of the dense list (starts at nk[3])
                    {
                        cipc_ctx::Scope sk(c, "k_hessian_factor");
                        if (nk[0]) CIPC_LAUNCH(k_hessian_factor<0>, div_up(nk[0], 128), 128, 0, c->st, c->X.p, c->cs.p, c->info.p, c->tripOff.p,
                            c->clsIdx[0].p, nk[0], bp, Y0, h0, dl, dn);
                        if (nk[1]) CIPC_LAUNCH(k_hessian_factor<1>, div_up(nk[1], 128), 128, 0, c->st, c->X.p, c->cs.p, c->info.p, c->tripOff.p,
                            c->clsIdx[1].p, nk[1], bp, Y1, h1, dl, dn);
                        if (nk[2]) CIPC_LAUNCH(k_hessian_factor<2>, div_up(nk[2], 128), 128, 0, c->st, c->X.p, c->cs.p, c->info.p, c->tripOff.p,
                            c->clsIdx[2].p, nk[2], bp, Y2, h2, dl, dn);
                    }
                    c->factorValid = true;
                    c->expanded = false;
                    c->Y0 = Y0; c->Y1 = Y1; c->Y2 = Y2; c->h0 = h0; c->h1 = h1; c->h2 = h2;
                    for (int k = 0; k < 4; ++k) c->nk[k] = nk[k];
                    c->ny[0] = 3; c->ny[1] = 2; c->ny[2] = 1;
                    // mollified stencils + rejected ones: dense eigen path, list length read on the device
                    CIPC_LAUNCH(k_barrier_hessian, 1184, DENSE_BD, DENSE_SMEM, c->st, c->X.p, c->X0.p, c->cs.p, c->info.p, c->tripOff.p, c->clsIdx[3].p, 0u,
                        (const u32*)dn, bp, projectSPD, outT, 0);
                }
                else {
                    cipc_ctx::Scope sk(c, "k_barrier_hessian");
                    if (nk[0]) CIPC_LAUNCH(k_hessian_lowrank<0>, div_up(nk[0], 128), 128, 0, c->st, c->X.p, c->cs.p, c->info.p, c->tripOff.p,
                        c->clsIdx[0].p, nk[0], bp, projectSPD, outT, bm);
                    if (nk[1]) CIPC_LAUNCH(k_hessian_lowrank<1>, div_up(nk[1], 128), 128, 0, c->st, c->X.p, c->cs.p, c->info.p, c->tripOff.p,
                        c->clsIdx[1].p, nk[1], bp, projectSPD, outT, bm);
                    if (nk[2]) CIPC_LAUNCH(k_hessian_lowrank<2>, div_up(nk[2], 128), 128, 0, c->st, c->X.p, c->cs.p, c->info.p, c->tripOff.p,
                        c->clsIdx[2].p, nk[2], bp, projectSPD, outT, bm);
                    if (nk[3]) CIPC_LAUNCH(k_barrier_hessian, div_up(nk[3], DENSE_BD), DENSE_BD, DENSE_SMEM, c->st, c->X.p, c->X0.p, c->cs.p, c->info.p, c->tripOff.p,
                        c->clsIdx[3].p, nk[3], (const u32*)nullptr, bp, projectSPD, outT, bm);
                }
                c->ctr["hessian_4pt"] = nk[0]; c->ctr["hessian_pe"] = nk[1]; c->ctr["hessian_pp"] = nk[2]; c->ctr["hessian_mollified"] = nk[3];
            }
            if (blk) merge_blocks(c, c->cs.p, c->nC, tot);
        }
        if (nTrip) *nTrip = c->nTrip;
        return (int)CIPC_OK;
    });
}
int cipc_barrier_hessian(cipc_ctx* ctx, int elastic, double dHat2, const double kappa[3], double thickness, int projectSPD,
    int64_t* nTrip)
{
    if (ctx && ctx->multi) return multi_barrier_hessian(ctx, elastic, dHat2, kappa, thickness, projectSPD, nTrip, false);
    return barrier_hessian_impl(ctx, elastic, dHat2, kappa, thickness, projectSPD, nTrip, H_HOST);
}
int cipc_barrier_hessian_merged(cipc_ctx* ctx, int elastic, double dHat2, const double kappa[3], double thickness, int projectSPD,
    int64_t* nTrip)
{
    if (ctx && ctx->multi) return multi_barrier_hessian(ctx, elastic, dHat2, kappa, thickness, projectSPD, nTrip, true);
    return barrier_hessian_impl(ctx, elastic, dHat2, kappa, thickness, projectSPD, nTrip, H_BLK);
}
int cipc_barrier_hessian_dev(cipc_ctx* ctx, int elastic, double dHat2, const double kappa[3], double thickness, int projectSPD,
    int64_t* nTrip)
{
    CIPC_MULTI_UNSUPPORTED(ctx);
    return barrier_hessian_impl(ctx, elastic, dHat2, kappa, thickness, projectSPD, nTrip, H_DEV);
}
int cipc_barrier_gradient_hessian_dev(cipc_ctx* ctx, int elastic, double dHat2, const double kappa[3], double thickness, int64_t* nTrip)
{
    CIPC_MULTI_UNSUPPORTED(ctx);
    return barrier_hessian_impl(ctx, elastic, dHat2, kappa, thickness, 1, nTrip, H_DEV, true);
}
int cipc_get_triplets(cipc_ctx* ctx, cipc_triplet* out)
{
    Prefault::get().wait(); // a background first touch of the destination (cipc_host_prefault_async) must have finished
    if (ctx && ctx->multi) return multi_get_triplets(ctx, out);
    return guarded(ctx, [&]() {
        if (!ctx->nTrip) return (int)CIPC_OK;
        if (ctx->mergedValid) { deliver_merged_host(ctx, out); return (int)CIPC_OK; }
        const char* dma = getenv("CIPC_TRIPLETS_DMA"); // =1: always copy the expanded stream over PCIe
        if (ctx->factorValid && !(dma && dma[0] == '1')) deliver_triplets_host(ctx, out);
        else {
            expand_on_device(ctx);
            CIPC_CUDA(cudaMemcpyAsync(out, ctx->trip.p, (size_t)ctx->nTrip * sizeof(cipc_triplet), cudaMemcpyDeviceToHost, ctx->st));
            CIPC_CUDA(cudaStreamSynchronize(ctx->st));
        }
        return (int)CIPC_OK;
    });
}
cipc_triplet* cipc_dev_triplets(cipc_ctx* ctx)
{
    if (!ctx || ctx->multi) return nullptr;
    int r = guarded(ctx, [&]() {
        ctx->begin_call();
        ctx->trip.reserve(1, ctx->st); // an empty stream (no constraints) still has a valid address
        expand_on_device(ctx);
        return (int)CIPC_OK;
    });
    return r == CIPC_OK ? ctx->trip.p : nullptr;
}
int cipc_step_size_dev(cipc_ctx* ctx, int elastic, double thickness, double stepIn)
{
    CIPC_MULTI_UNSUPPORTED(ctx);
    return guarded(ctx, [&]() { ctx->begin_call(); return do_step_size(ctx, elastic, thickness, stepIn); });
}
int cipc_step_size(cipc_ctx* ctx, int elastic, double thickness, double* step)
{
    if (ctx && ctx->multi) return multi_step_size(ctx, elastic, thickness, step);
    return guarded(ctx, [&]() {
        ctx->begin_call();
        int r = do_step_size(ctx, elastic, thickness, *step);
        if (r) return r;
        double v;
        CIPC_CUDA(cudaMemcpyAsync(&v, ctx->scal.p + 1, sizeof(double), cudaMemcpyDeviceToHost, ctx->st));
        r = fetch_err(ctx);
        if (r) return r;
        *step = v;
        return (int)CIPC_OK;
    });
}
int cipc_min_dist2_dev(cipc_ctx* ctx, double thickness)
{
    CIPC_MULTI_UNSUPPORTED(ctx);
    (void)thickness;
    return guarded(ctx, [&]() { ctx->begin_call(); return do_min_dist(ctx, false); });
}
int cipc_min_dist2(cipc_ctx* ctx, double thickness, double* dist2, double* minDist2)
{
    if (ctx && ctx->multi) return multi_min_dist2(ctx, thickness, dist2, minDist2);
    return guarded(ctx, [&]() {
        cipc_ctx* c = ctx;
        c->begin_call();
        if (c->nC == 0) return (int)CIPC_OK; // the reference returns early on an empty set and leaves dist2 / minDist2 untouched (IPC.h:2253)
        int r = do_min_dist(c, dist2 != nullptr);
        if (r) return r;
        long long bits;
        CIPC_CUDA(cudaMemcpyAsync(&bits, c->scal.p + 2, 8, cudaMemcpyDeviceToHost, c->st));
        CIPC_CUDA(cudaStreamSynchronize(c->st));
        if (dist2) { advise_huge(dist2, (size_t)c->nC * 8); staged_d2h(c->ring, c->st, dist2, c->dist2.p, (size_t)c->nC * 8); }
        bits = bits >= 0 ? bits : (bits ^ 0x7fffffffffffffffLL);
        double m;
        memcpy(&m, &bits, 8);
        *minDist2 = m - thickness * thickness;
        return (int)CIPC_OK;
    });
}

// ---- friction (FEM/FRICTION.h)
int cipc_set_prev_positions(cipc_ctx* ctx, const double* Xn, int stride_bytes)
{
    CIPC_MULTI_UNSUPPORTED(ctx);
    return guarded(ctx, [&]() {
        need(ctx->T.nV > 0, "topology not set");
        upload_vec3(ctx, ctx->Xn, Xn, stride_bytes, &ctx->tagXn);
        ctx->haveXn = true;
        return (int)CIPC_OK;
    });
}
int cipc_friction_basis(cipc_ctx* ctx, int elastic, double dHat2, const double kappa[3], double thickness, int* nF_out)
{
    CIPC_MULTI_UNSUPPORTED(ctx);
    return guarded(ctx, [&]() {
        ctx->begin_call();
        int r = do_friction_basis(ctx, elastic, dHat2, kappa, thickness);
        if (nF_out) *nF_out = (int)ctx->nF;
        return r;
    });
}
int cipc_get_friction_basis(cipc_ctx* ctx, int32_t* fcs, double* closestPoint, double* tanBasis, double* normalForce)
{
    CIPC_MULTI_UNSUPPORTED(ctx);
    return guarded(ctx, [&]() {
        cipc_ctx* c = ctx;
        const size_t n = c->nF;
        if (!n) return (int)CIPC_OK;
        if (fcs) staged_d2h(c->ring, c->st, fcs, c->fcs.p, n * 16);
        if (closestPoint) staged_d2h(c->ring, c->st, closestPoint, c->fcp.p, n * 16);
        if (tanBasis) staged_d2h(c->ring, c->st, tanBasis, c->fB.p, n * 48);
        if (normalForce) staged_d2h(c->ring, c->st, normalForce, c->fnf.p, n * 8);
        CIPC_CUDA(cudaStreamSynchronize(c->st));
        return (int)CIPC_OK;
    });
}
int cipc_set_friction_basis(cipc_ctx* ctx, const int32_t* fcs, const double* closestPoint, const double* tanBasis, const double* normalForce,
    int nF)
{
    CIPC_MULTI_UNSUPPORTED(ctx);
    return guarded(ctx, [&]() {
        cipc_ctx* c = ctx;
        if (nF < 0 || (nF && (!fcs || !normalForce))) return (int)CIPC_ERR_ARG;
        const size_t n = (size_t)nF;
        c->fcs.reserve(n + 1, c->st); c->fcp.reserve(n + 1, c->st); c->fB.reserve(n * 6 + 2, c->st); c->fnf.reserve(n + 1, c->st);
        if (n) {
            staged_h2d(c->ring, c->st, c->fcs.p, fcs, n * 16);
            if (closestPoint) staged_h2d(c->ring, c->st, c->fcp.p, closestPoint, n * 16);
            if (tanBasis) staged_h2d(c->ring, c->st, c->fB.p, tanBasis, n * 48);
            staged_h2d(c->ring, c->st, c->fnf.p, normalForce, n * 8);
            CIPC_CUDA(cudaStreamSynchronize(c->st));
        }
        c->nF = (u32)nF;
        return (int)CIPC_OK;
    });
}
int cipc_friction_coef(cipc_ctx* ctx, int nComp, const int32_t* compNodeRange, const double* muComp, double* mu_out)
{
    CIPC_MULTI_UNSUPPORTED(ctx);
    return guarded(ctx, [&]() {
        cipc_ctx* c = ctx;
        if (nComp <= 0 || !compNodeRange || !muComp) return (int)CIPC_ERR_ARG;
        c->begin_call();
        c->compRange.reserve(nComp, c->st); c->muComp.reserve((size_t)nComp * nComp, c->st);
        CIPC_CUDA(cudaMemcpyAsync(c->compRange.p, compNodeRange, (size_t)nComp * 4, cudaMemcpyHostToDevice, c->st));
        CIPC_CUDA(cudaMemcpyAsync(c->muComp.p, muComp, (size_t)nComp * nComp * 8, cudaMemcpyHostToDevice, c->st));
        CIPC_CUDA(cudaMemsetAsync(c->errFlag.p, 0, sizeof(int), c->st));
        if (c->nF) CIPC_LAUNCH(k_friction_coef, div_up(c->nF, TB), TB, 0, c->st, c->fcs.p, c->nF, c->compRange.p, nComp, c->muComp.p, c->fnf.p,
            c->errFlag.p);
        const int r = fetch_err(c);
        if (r) return r;
        if (mu_out) *mu_out = 1.0; // FRICTION.h:132
        return (int)CIPC_OK;
    });
}
int cipc_friction_energy_dev(cipc_ctx* ctx, double epsvh2, double mu)
{
    CIPC_MULTI_UNSUPPORTED(ctx);
    return guarded(ctx, [&]() { ctx->begin_call(); return do_friction_energy(ctx, epsvh2, mu); });
}
int cipc_friction_energy(cipc_ctx* ctx, double epsvh2, double mu, double* E)
{
    CIPC_MULTI_UNSUPPORTED(ctx);
    return guarded(ctx, [&]() {
        ctx->begin_call();
        int r = do_friction_energy(ctx, epsvh2, mu);
        if (r) return r;
        double v;
        CIPC_CUDA(cudaMemcpyAsync(&v, ctx->scal.p + 4, sizeof(double), cudaMemcpyDeviceToHost, ctx->st));
        CIPC_CUDA(cudaStreamSynchronize(ctx->st));
        *E += v;
        return (int)CIPC_OK;
    });
}
int cipc_friction_gradient_dev(cipc_ctx* ctx, double epsvh2, double mu, int accumulate)
{
    CIPC_MULTI_UNSUPPORTED(ctx);
    return guarded(ctx, [&]() { ctx->begin_call(); return do_friction_gradient(ctx, epsvh2, mu, accumulate != 0); });
}
int cipc_friction_gradient(cipc_ctx* ctx, double epsvh2, double mu, double* g, int stride)
{
    CIPC_MULTI_UNSUPPORTED(ctx);
    return guarded(ctx, [&]() {
        cipc_ctx* c = ctx;
        if (stride < 24 || stride % 8) return (int)CIPC_ERR_ARG;
        c->begin_call();
        int r = do_friction_gradient(c, epsvh2, mu, false);
        if (r) return r;
        const size_t n = (size_t)c->T.nV;
        double* h = (double*)c->pin.reserve(n * 24);
        CIPC_CUDA(cudaMemcpyAsync(h, c->g.p, n * 24, cudaMemcpyDeviceToHost, c->st));
        CIPC_CUDA(cudaStreamSynchronize(c->st));
        const size_t sd = stride / 8, CH = 1 << 14;
        HostPool::get().for_each((n + CH - 1) / CH, [&](size_t k) { // nodeAttr.g += (FRICTION.h:294-297)
            for (size_t v = k * CH, e = std::min(n, v + CH); v < e; ++v) { g[v * sd] += h[3 * v]; g[v * sd + 1] += h[3 * v + 1]; g[v * sd + 2] += h[3 * v + 2]; }
        });
        return (int)CIPC_OK;
    });
}
int cipc_friction_hessian(cipc_ctx* ctx, double epsvh2, double mu, int projectSPD, int64_t* nTrip)
{
    CIPC_MULTI_UNSUPPORTED(ctx);
    (void)projectSPD; // the inner 2x2 matrix is positive semi-definite in closed form: makePD is the identity (friction.cuh)
    return guarded(ctx, [&]() {
        ctx->begin_call();
        int r = do_friction_hessian(ctx, epsvh2, mu, HF_HOST);
        if (nTrip) *nTrip = ctx->nTrip;
        return r;
    });
}
int cipc_friction_hessian_dev(cipc_ctx* ctx, double epsvh2, double mu, int projectSPD, int64_t* nTrip)
{
    CIPC_MULTI_UNSUPPORTED(ctx);
    (void)projectSPD;
    return guarded(ctx, [&]() {
        ctx->begin_call();
        int r = do_friction_hessian(ctx, epsvh2, mu, HF_DEV);
        if (nTrip) *nTrip = ctx->nTrip;
        return r;
    });
}

int cipc_friction_hessian_merged(cipc_ctx* ctx, double epsvh2, double mu, int projectSPD, int64_t* nTrip)
{
    CIPC_MULTI_UNSUPPORTED(ctx);
    (void)projectSPD;
    return guarded(ctx, [&]() {
        ctx->begin_call();
        int r = do_friction_hessian(ctx, epsvh2, mu, HF_BLK);
        if (nTrip) *nTrip = ctx->nTrip;
        return r;
    });
}

// ---- device-resident line search: Shell/IMPLICIT_EULER.h:102-131 without a position upload per trial
int cipc_save_positions(cipc_ctx* ctx)
{
    CIPC_MULTI_UNSUPPORTED(ctx);
    return guarded(ctx, [&]() {
        cipc_ctx* c = ctx;
        need(c->haveX, "positions not set");
        c->Xprev.reserve(c->T.nV, c->st);
        CIPC_CUDA(cudaMemcpyAsync(c->Xprev.p, c->X.p, (size_t)c->T.nV * 32, cudaMemcpyDeviceToDevice, c->st));
        c->haveXprev = true;
        return (int)CIPC_OK;
    });
}
int cipc_step_positions(cipc_ctx* ctx, double alpha)
{
    CIPC_MULTI_UNSUPPORTED(ctx);
    return guarded(ctx, [&]() {
        cipc_ctx* c = ctx;
        need(c->haveXprev && c->haveP, "saved positions / search direction not set");
        CIPC_LAUNCH(k_step_positions, div_up(c->T.nV, TB), TB, 0, c->st, c->Xprev.p, c->P.p, alpha, c->T.nV, c->X.p);
        c->tagX = 0; // the resident positions no longer mirror any host array
        c->haveX = true;
        ++c->xVersion;
        return (int)CIPC_OK;
    });
}
int cipc_get_positions(cipc_ctx* ctx, double* X, int stride_bytes)
{
    CIPC_MULTI_UNSUPPORTED(ctx);
    return guarded(ctx, [&]() {
        cipc_ctx* c = ctx;
        need(c->haveX, "positions not set");
        if (stride_bytes != 32 && stride_bytes != 24) return (int)CIPC_ERR_ARG;
        const size_t n = (size_t)c->T.nV;
        if (stride_bytes == 32) {
            CIPC_CUDA(cudaMemcpyAsync(X, c->X.p, n * 32, cudaMemcpyDeviceToHost, c->st));
            CIPC_CUDA(cudaStreamSynchronize(c->st));
            return (int)CIPC_OK;
        }
        double* h = (double*)c->pin.reserve(n * 32);
        CIPC_CUDA(cudaMemcpyAsync(h, c->X.p, n * 32, cudaMemcpyDeviceToHost, c->st));
        CIPC_CUDA(cudaStreamSynchronize(c->st));
        for (size_t v = 0; v < n; ++v) { X[3 * v] = h[4 * v]; X[3 * v + 1] = h[4 * v + 1]; X[3 * v + 2] = h[4 * v + 2]; }
        return (int)CIPC_OK;
    });
}

// ---- triplets -> CSR (Math/CSR_MATRIX.h:49-56)
int cipc_csr_begin(cipc_ctx* ctx)
{
    CIPC_MULTI_UNSUPPORTED(ctx);
    return guarded(ctx, [&]() { ctx->nBlk = 0; ctx->nU = 0; ctx->csrValid = false; return (int)CIPC_OK; });
}
int cipc_csr_add(cipc_ctx* ctx)
{
    CIPC_MULTI_UNSUPPORTED(ctx);
    return guarded(ctx, [&]() {
        cipc_ctx* c = ctx;
        c->begin_call();
        if (!c->nTrip || !c->hessN) return (int)CIPC_OK;
        expand_on_device(c); // the blocks are read from the device-resident triplet stream
        cipc_ctx::Scope sc(c, "csr_add");
        const size_t add = (size_t)(c->nTrip / 9), tot = c->nBlk + add;
        if (tot >= 0xfffffff0ull) throw std::runtime_error("CSR assembly: more than 2^32 blocks");
        c->blkVal.reserve(tot * 9, c->st, true); c->keyI.reserve(tot, c->st, true); c->keyJ.reserve(tot, c->st, true);
        CIPC_LAUNCH(k_trip_to_blocks, div_up((size_t)c->hessN * 32, 256), 256, 0, c->st, c->trip.p, c->hessStencils, c->tripOff.p, c->hessN,
            c->nBlk, c->blkVal.p, c->keyI.p, c->keyJ.p);
        c->nBlk = tot;
        c->csrValid = false;
        return (int)CIPC_OK;
    });
}
int cipc_csr_finish(cipc_ctx* ctx, int64_t* nnz_out)
{
    CIPC_MULTI_UNSUPPORTED(ctx);
    return guarded(ctx, [&]() {
        cipc_ctx* c = ctx;
        need(c->T.nV > 0, "topology not set");
        c->begin_call();
        const size_t n = c->nBlk;
        const int nV = c->T.nV;
        c->csrRowPtr.reserve((size_t)3 * nV + 1, c->st);
        c->nU = 0;
        if (n == 0) {
            CIPC_CUDA(cudaMemsetAsync(c->csrRowPtr.p, 0, ((size_t)3 * nV + 1) * sizeof(int), c->st));
            c->csrValid = true;
            if (nnz_out) *nnz_out = 0;
            return (int)CIPC_OK;
        }
        int bits = 1;
        while ((1ll << bits) < (long long)nV) ++bits;
        u32 nU = 0;
        {
            cipc_ctx::Scope sc(c, "csr_sort");
            c->csrKey.reserve(n, c->st); c->csrIds.reserve(n, c->st);
            CIPC_LAUNCH(k_iota_copy, div_up(n, TB), TB, 0, c->st, c->keyJ.p, c->csrKey.p, c->csrIds.p, n);
            device_radix_sort(c->csrKey.p, c->csrIds.p, n, bits, c->sortwk, c->st);                      // by column vertex
            CIPC_LAUNCH(k_gather_u32, div_up(n, TB), TB, 0, c->st, c->keyI.p, c->csrIds.p, c->csrKey.p, n);
            device_radix_sort(c->csrKey.p, c->csrIds.p, n, bits, c->sortwk, c->st);                      // then (stable) by row vertex
        }
        {
            cipc_ctx::Scope sc(c, "csr_pattern");
            c->csrColS.reserve(n, c->st); c->csrHeads.reserve(n, c->st); c->csrScan.reserve(n, c->st);
            CIPC_LAUNCH(k_csr_heads, div_up(n, TB), TB, 0, c->st, c->csrKey.p, c->keyJ.p, c->csrIds.p, n, c->csrColS.p, c->csrHeads.p);
            device_excl_scan(c->csrHeads.p, c->csrScan.p, n, c->scanwk, c->st);
            CIPC_CUDA(cudaMemcpyAsync(&nU, c->scanwk.total.p, sizeof(u32), cudaMemcpyDeviceToHost, c->st));
            CIPC_CUDA(cudaStreamSynchronize(c->st));
            if ((uint64_t)nU * 9 > 0x7fffffffull) throw std::runtime_error("CSR assembly: nnz exceeds the reference's 32-bit index type");
            c->urow.reserve(nU, c->st); c->ucol.reserve(nU, c->st); c->ustart.reserve(nU, c->st); c->browCnt.reserve((size_t)nV + 1, c->st);
            CIPC_CUDA(cudaMemsetAsync(c->browCnt.p, 0, ((size_t)nV + 1) * 4, c->st));
            CIPC_LAUNCH(k_csr_unique, div_up(n, TB), TB, 0, c->st, c->csrKey.p, c->csrColS.p, c->csrHeads.p, c->csrScan.p, n, c->urow.p, c->ucol.p,
                c->ustart.p, c->browCnt.p);
            device_excl_scan(c->browCnt.p, c->browCnt.p, (size_t)nV + 1, c->scanwk, c->st);
        }
        {
            cipc_ctx::Scope sc(c, "csr_emit");
            c->csrColIdx.reserve((size_t)nU * 9, c->st); c->csrVal.reserve((size_t)nU * 9, c->st);
            CIPC_LAUNCH(k_csr_rowptr, div_up((size_t)nV + 1, TB), TB, 0, c->st, c->browCnt.p, nV, nU, c->csrRowPtr.p);
            CIPC_LAUNCH(k_csr_emit, div_up(nU, 32), 288, 0, c->st, c->blkVal.p, c->csrIds.p, c->urow.p, c->ucol.p, c->ustart.p, nU, n, c->browCnt.p,
                c->csrColIdx.p, c->csrVal.p);
        }
        c->nU = nU;
        c->csrValid = true;
        c->ctr["csr_blocks_in"] = (int64_t)n; c->ctr["csr_blocks_unique"] = nU;
        if (nnz_out) *nnz_out = (int64_t)nU * 9;
        return (int)CIPC_OK;
    });
}
int cipc_get_csr(cipc_ctx* ctx, int32_t* rowPtr, int32_t* colIdx, double* val)
{
    CIPC_MULTI_UNSUPPORTED(ctx);
    return guarded(ctx, [&]() {
        cipc_ctx* c = ctx;
        need(c->csrValid, "no assembled CSR matrix (call cipc_csr_finish)");
        const size_t nnz = (size_t)c->nU * 9;
        if (rowPtr) CIPC_CUDA(cudaMemcpyAsync(rowPtr, c->csrRowPtr.p, ((size_t)3 * c->T.nV + 1) * sizeof(int), cudaMemcpyDeviceToHost, c->st));
        if (colIdx && nnz) CIPC_CUDA(cudaMemcpyAsync(colIdx, c->csrColIdx.p, nnz * sizeof(int), cudaMemcpyDeviceToHost, c->st));
        if (val && nnz) CIPC_CUDA(cudaMemcpyAsync(val, c->csrVal.p, nnz * sizeof(double), cudaMemcpyDeviceToHost, c->st));
        CIPC_CUDA(cudaStreamSynchronize(c->st));
        return (int)CIPC_OK;
    });
}

// ---- boundary-primitive construction: Utils/MESHIO.h:768-834 + Shell/IMPLICIT_EULER.h:245-276
int cipc_build_boundary(cipc_ctx* ctx, int nV, const double* X, int x_stride_bytes, int nTri, const int32_t* tri, int tri_stride, int nSeg,
    const int32_t* seg, int seg_stride, int nRod, const int32_t* rod, int rod_stride, const double* rodRadius, int nParticle,
    const int32_t* particle, int32_t counts_out[6])
{
    CIPC_MULTI_UNSUPPORTED(ctx);
    return guarded(ctx, [&]() {
        cipc_ctx* c = ctx;
        if (nV <= 0 || nTri < 0 || nSeg < 0 || nRod < 0 || nParticle < 0 || !X || (x_stride_bytes != 24 && x_stride_bytes != 32) ||
            (nTri && (tri_stride != 3 && tri_stride != 4)) || (nSeg && seg_stride != 2 && seg_stride != 4) || (nRod && rod_stride != 2 && rod_stride != 4) ||
            (nRod && !rodRadius))
            return (int)CIPC_ERR_ARG;
        c->begin_call();
        // content hash of every input: an unchanged call returns the lists of the previous build
        u64 h = 0x9AD0C0DEULL;
        const int hdr[8] = {nV, nTri, nSeg, nRod, nParticle, tri_stride, seg_stride, rod_stride};
        h = fnv(hdr, sizeof(hdr), h);
        const HashSeg segs[6] = {{(const unsigned char*)tri, (size_t)nTri * tri_stride * 4}, {(const unsigned char*)seg, (size_t)nSeg * seg_stride * 4},
            {(const unsigned char*)rod, (size_t)nRod * rod_stride * 4}, {(const unsigned char*)rodRadius, (size_t)nRod * 8},
            {(const unsigned char*)particle, (size_t)nParticle * 4}, {(const unsigned char*)X, (size_t)nV * x_stride_bytes}};
        h = hash_segments(segs, 6, h) | 1ULL;
        auto counts = [&]() {
            if (!counts_out) return;
            counts_out[0] = (int)c->hBN.size(); counts_out[1] = (int)c->hBE.size(); counts_out[2] = (int)(c->hTri.size() / 3);
            counts_out[3] = c->bdCodim[0]; counts_out[4] = c->bdCodim[1]; counts_out[5] = (int)c->hBNArea.size();
        };
        if (h == c->bdHash) { counts(); return (int)CIPC_OK; }
        cipc_ctx::Scope sc(c, "build_boundary");
        // node positions: a private upload (the context's resident X belongs to the contact calls and needs a topology)
        DevBuf<double4> Xd;
        Xd.reserve(nV, c->st);
        {
            const int keepNV = c->T.nV;
            c->T.nV = nV; // upload_vec3 sizes by T.nV
            try { upload_vec3(c, Xd, X, x_stride_bytes, nullptr); } catch (...) { c->T.nV = keepNV; throw; }
            c->T.nV = keepNV;
        }
        int bits = 1;
        while ((1ll << bits) < (long long)nV) ++bits;
        u32 nUE = 0, nNode = 0, nKeep = 0;
        c->hTri.resize((size_t)3 * nTri);
        if (nTri) {
            for (int t = 0; t < nTri; ++t) for (int k = 0; k < 3; ++k) {
                const int v = tri[(size_t)t * tri_stride + k];
                if (v < 0 || v >= nV) return (int)CIPC_ERR_ARG;
                c->hTri[(size_t)3 * t + k] = v;
            }
            const size_t nE = (size_t)3 * nTri;
            if (nE >= 0xfffffff0ull) return (int)CIPC_ERR_ARG;
            c->bdTri.reserve(nTri, c->st); c->bdThird.reserve(nTri, c->st); c->bdBTArea.reserve(nTri, c->st);
            c->stageI.reserve((size_t)nTri * tri_stride, c->st);
            staged_h2d(c->ring, c->st, c->stageI.p, tri, (size_t)nTri * tri_stride * 4);
            CIPC_LAUNCH(k_pack_tris, div_up(nTri, TB), TB, 0, c->st, c->stageI.p, tri_stride, nTri, c->bdTri.p);
            CIPC_LAUNCH(k_bd_tri_area, div_up(nTri, TB), TB, 0, c->st, Xd.p, c->bdTri.p, nTri, c->bdBTArea.p, c->bdThird.p);
            c->bdLo.reserve(nE, c->st); c->bdHi.reserve(nE, c->st); c->bdV.reserve(nE, c->st); c->bdIds.reserve(nE, c->st); c->bdKey.reserve(nE, c->st);
            c->bdHeads.reserve(nE, c->st); c->bdScan.reserve(nE, c->st);
            CIPC_LAUNCH(k_bd_events, div_up(nE, TB), TB, 0, c->st, c->bdTri.p, nTri, c->bdLo.p, c->bdHi.p, c->bdV.p);
            // ---- edges: events ordered by (min vertex, max vertex, visiting order)
            CIPC_LAUNCH(k_iota_copy, div_up(nE, TB), TB, 0, c->st, c->bdHi.p, c->bdKey.p, c->bdIds.p, nE);
            device_radix_sort(c->bdKey.p, c->bdIds.p, nE, bits, c->sortwk, c->st);
            CIPC_LAUNCH(k_gather_u32, div_up(nE, TB), TB, 0, c->st, c->bdLo.p, c->bdIds.p, c->bdKey.p, nE);
            device_radix_sort(c->bdKey.p, c->bdIds.p, nE, bits, c->sortwk, c->st);
            CIPC_LAUNCH(k_bd_heads, div_up(nE, TB), TB, 0, c->st, c->bdLo.p, c->bdHi.p, c->bdIds.p, (u32)nE, c->bdHeads.p);
            device_excl_scan(c->bdHeads.p, c->bdScan.p, nE, c->scanwk, c->st);
            CIPC_CUDA(cudaMemcpyAsync(&nUE, c->scanwk.total.p, 4, cudaMemcpyDeviceToHost, c->st));
            CIPC_CUDA(cudaStreamSynchronize(c->st));
            c->bdUeA.reserve((size_t)nUE + 1, c->st); c->bdUeB.reserve((size_t)nUE + 1, c->st); c->bdUeVal.reserve((size_t)nUE + 1, c->st);
            c->bdIds2.reserve((size_t)nUE + 1, c->st); c->bdBE.reserve((size_t)nUE + 1, c->st); c->bdBEArea.reserve((size_t)nUE + 1, c->st);
            CIPC_LAUNCH(k_bd_edge_fold, div_up(nE, TB), TB, 0, c->st, c->bdTri.p, c->bdThird.p, c->bdLo.p, c->bdHi.p, c->bdIds.p, c->bdHeads.p, c->bdScan.p,
                (u32)nE, c->bdUeA.p, c->bdUeB.p, c->bdUeVal.p);
            // oriented edges in lexicographic (a, b) order -- the iteration order of the reference's std::map
            CIPC_LAUNCH(k_iota_copy, div_up(nUE, TB), TB, 0, c->st, c->bdUeB.p, c->bdKey.p, c->bdIds2.p, (size_t)nUE);
            device_radix_sort(c->bdKey.p, c->bdIds2.p, nUE, bits, c->sortwk, c->st);
            CIPC_LAUNCH(k_gather_u32, div_up(nUE, TB), TB, 0, c->st, c->bdUeA.p, c->bdIds2.p, c->bdKey.p, (size_t)nUE);
            device_radix_sort(c->bdKey.p, c->bdIds2.p, nUE, bits, c->sortwk, c->st);
            CIPC_LAUNCH(k_bd_edge_emit, div_up(nUE, TB), TB, 0, c->st, c->bdUeA.p, c->bdUeB.p, c->bdUeVal.p, c->bdIds2.p, nUE, c->bdBE.p, c->bdBEArea.p);
            // ---- nodes: events ordered by (vertex, visiting order)
            CIPC_LAUNCH(k_iota_copy, div_up(nE, TB), TB, 0, c->st, c->bdV.p, c->bdKey.p, c->bdIds.p, nE);
            device_radix_sort(c->bdKey.p, c->bdIds.p, nE, bits, c->sortwk, c->st);
            CIPC_LAUNCH(k_bd_heads, div_up(nE, TB), TB, 0, c->st, c->bdV.p, (const u32*)nullptr, c->bdIds.p, (u32)nE, c->bdHeads.p);
            device_excl_scan(c->bdHeads.p, c->bdScan.p, nE, c->scanwk, c->st);
            CIPC_CUDA(cudaMemcpyAsync(&nNode, c->scanwk.total.p, 4, cudaMemcpyDeviceToHost, c->st));
            CIPC_CUDA(cudaStreamSynchronize(c->st));
            c->bdNodeV.reserve((size_t)nNode + 1, c->st); c->bdNodeSum.reserve((size_t)nNode + 1, c->st); c->bdKeep.reserve((size_t)nNode + 1, c->st);
            c->bdKeepScan.reserve((size_t)nNode + 1, c->st); c->bdBN.reserve((size_t)nNode + 1, c->st); c->bdBNArea.reserve((size_t)nNode + 1, c->st);
            CIPC_LAUNCH(k_bd_node_fold, div_up(nE, TB), TB, 0, c->st, c->bdThird.p, c->bdV.p, c->bdIds.p, c->bdHeads.p, c->bdScan.p, (u32)nE, c->bdNodeV.p,
                c->bdNodeSum.p, c->bdKeep.p);
            device_excl_scan(c->bdKeep.p, c->bdKeepScan.p, nNode, c->scanwk, c->st);
            CIPC_CUDA(cudaMemcpyAsync(&nKeep, c->scanwk.total.p, 4, cudaMemcpyDeviceToHost, c->st));
            CIPC_LAUNCH(k_bd_node_emit, div_up(nNode, TB), TB, 0, c->st, c->bdNodeV.p, c->bdNodeSum.p, c->bdKeep.p, c->bdKeepScan.p, nNode, c->bdBN.p, c->bdBNArea.p);
            CIPC_CUDA(cudaStreamSynchronize(c->st));
        }
        // ---- host lists: surface primitives from the device, then the appends of Shell/IMPLICIT_EULER.h:245-276
        c->hBN.resize(nKeep); c->hBNArea.resize(nKeep); c->hBE.resize(nUE); c->hBEArea.resize(nUE); c->hBTArea.resize(nTri);
        if (nKeep) { staged_d2h(c->ring, c->st, c->hBN.data(), c->bdBN.p, (size_t)nKeep * 4); staged_d2h(c->ring, c->st, c->hBNArea.data(), c->bdBNArea.p, (size_t)nKeep * 8); }
        if (nUE) { staged_d2h(c->ring, c->st, c->hBE.data(), c->bdBE.p, (size_t)nUE * 8); staged_d2h(c->ring, c->st, c->hBEArea.data(), c->bdBEArea.p, (size_t)nUE * 8); }
        if (nTri) staged_d2h(c->ring, c->st, c->hBTArea.data(), c->bdBTArea.p, (size_t)nTri * 8);
        auto xyz = [&](int v, double* o) { const double* p = (const double*)((const char*)X + (size_t)v * x_stride_bytes); o[0] = p[0]; o[1] = p[1]; o[2] = p[2]; };
        for (int i = 0; i < nSeg; ++i) { // :245-250 -- edges appended, both end nodes appended (duplicates kept, no BNArea entry)
            const int a = seg[(size_t)i * seg_stride], b = seg[(size_t)i * seg_stride + 1];
            if (a < 0 || a >= nV || b < 0 || b >= nV) return (int)CIPC_ERR_ARG;
            c->hBE.push_back(make_int2(a, b));
        }
        for (int i = 0; i < nSeg; ++i) { c->hBN.push_back(seg[(size_t)i * seg_stride]); c->hBN.push_back(seg[(size_t)i * seg_stride + 1]); }
        std::vector<std::pair<int, double>> rodNode; // (node, half edge area) in rod order; folded per node in that order like std::map<int,T>::operator[] +=
        for (int i = 0; i < nRod; ++i) { // :252-266
            const int a = rod[(size_t)i * rod_stride], b = rod[(size_t)i * rod_stride + 1];
            if (a < 0 || a >= nV || b < 0 || b >= nV) return (int)CIPC_ERR_ARG;
            c->hBE.push_back(make_int2(a, b));
            double p0[3], p1[3];
            xyz(a, p0); xyz(b, p1);
            const double dx = p0[0] - p1[0], dy = p0[1] - p1[1], dz = p0[2] - p1[2];
            const double len = std::sqrt((dx * dx + dy * dy) + dz * dz);
            const double area = len * M_PI * rodRadius[i] / 6; // 1/6 of the cylinder surface participates in one contact
            rodNode.emplace_back(a, area / 2); rodNode.emplace_back(b, area / 2);
            c->hBEArea.push_back(area / 2);
        }
        c->bdCodim[0] = (int)c->hBN.size();
        std::stable_sort(rodNode.begin(), rodNode.end(), [](const std::pair<int, double>& x, const std::pair<int, double>& y) { return x.first < y.first; });
        for (size_t i = 0; i < rodNode.size();) { // :267-273 -- ascending node ids, areas summed in rod order
            size_t j = i;
            double s = 0.0;
            for (; j < rodNode.size() && rodNode[j].first == rodNode[i].first; ++j) s += rodNode[j].second;
            c->hBN.push_back(rodNode[i].first); c->hBNArea.push_back(s);
            i = j;
        }
        c->bdCodim[1] = (int)c->hBN.size();
        for (int i = 0; i < nParticle; ++i) { // :275-277
            if (particle[i] < 0 || particle[i] >= nV) return (int)CIPC_ERR_ARG;
            c->hBN.push_back(particle[i]);
        }
        c->bdHash = h;
        counts();
        return (int)CIPC_OK;
    });
}
int cipc_get_boundary(cipc_ctx* ctx, int32_t* BN, int32_t* BE, int be_stride, int32_t* BT, int bt_stride, double* BNArea, double* BEArea, double* BTArea)
{
    CIPC_MULTI_UNSUPPORTED(ctx);
    return guarded(ctx, [&]() {
        cipc_ctx* c = ctx;
        if (!c->bdHash) return (int)CIPC_ERR_ARG;
        if ((BE && be_stride != 2 && be_stride != 4) || (BT && bt_stride != 3 && bt_stride != 4)) return (int)CIPC_ERR_ARG;
        if (BN && !c->hBN.empty()) memcpy(BN, c->hBN.data(), c->hBN.size() * 4);
        if (BE) for (size_t i = 0; i < c->hBE.size(); ++i) {
            BE[i * be_stride] = c->hBE[i].x; BE[i * be_stride + 1] = c->hBE[i].y;
            for (int k = 2; k < be_stride; ++k) BE[i * be_stride + k] = 0;
        }
        if (BT) for (size_t t = 0; t < c->hTri.size() / 3; ++t) {
            for (int k = 0; k < 3; ++k) BT[t * bt_stride + k] = c->hTri[3 * t + k];
            for (int k = 3; k < bt_stride; ++k) BT[t * bt_stride + k] = 0;
        }
        if (BNArea && !c->hBNArea.empty()) memcpy(BNArea, c->hBNArea.data(), c->hBNArea.size() * 8);
        if (BEArea && !c->hBEArea.empty()) memcpy(BEArea, c->hBEArea.data(), c->hBEArea.size() * 8);
        if (BTArea && !c->hBTArea.empty()) memcpy(BTArea, c->hBTArea.data(), c->hBTArea.size() * 8);
        return (int)CIPC_OK;
    });
}

double* cipc_dev_positions(cipc_ctx* ctx) { return ctx && !ctx->multi ? (double*)ctx->X.p : nullptr; }
double* cipc_dev_gradient(cipc_ctx* ctx) { return ctx && !ctx->multi ? ctx->g.p : nullptr; }
double* cipc_dev_scalars(cipc_ctx* ctx) { return ctx && !ctx->multi ? ctx->scal.p : nullptr; }

double cipc_stage_ms(cipc_ctx* ctx, const char* stage)
{
    if (!ctx) return -1;
    if (ctx->multi) { // the slowest rank
        double m = -1;
        for (cipc_ctx* s : ctx->multi->sub) m = std::max(m, cipc_stage_ms(s, stage));
        return m;
    }
    cudaSetDevice(ctx->dev);
    cudaStreamSynchronize(ctx->st);
    double tot = -1;
    for (auto& s : ctx->stages)
        if (s.name == stage) {
            float ms = 0;
            if (cudaEventElapsedTime(&ms, s.a, s.b) == cudaSuccess) tot = (tot < 0 ? 0 : tot) + ms;
        }
    return tot;
}
int64_t cipc_counter(cipc_ctx* ctx, const char* name)
{
    if (!ctx) return -1;
    if (ctx->multi) { // summed over the ranks
        int64_t t = -1;
        for (cipc_ctx* s : ctx->multi->sub) { const int64_t v = cipc_counter(s, name); if (v >= 0) t = (t < 0 ? 0 : t) + v; }
        return t;
    }
    auto it = ctx->ctr.find(name);
    return it == ctx->ctr.end() ? -1 : it->second;
}

// ---- test hooks
int cipc_test_scan(cipc_ctx* ctx, const uint32_t* in, uint32_t* out, int64_t n, uint32_t* total)
{
    return guarded(ctx, [&]() {
        DevBuf<u32> d;
        d.reserve(n + 1, ctx->st);
        CIPC_CUDA(cudaMemcpyAsync(d.p, in, n * 4, cudaMemcpyHostToDevice, ctx->st));
        device_excl_scan(d.p, d.p, n, ctx->scanwk, ctx->st);
        CIPC_CUDA(cudaMemcpyAsync(out, d.p, n * 4, cudaMemcpyDeviceToHost, ctx->st));
        CIPC_CUDA(cudaMemcpyAsync(total, ctx->scanwk.total.p, 4, cudaMemcpyDeviceToHost, ctx->st));
        CIPC_CUDA(cudaStreamSynchronize(ctx->st));
        return (int)CIPC_OK;
    });
}
int cipc_test_sort(cipc_ctx* ctx, uint32_t* keys, uint32_t* vals, int64_t n, int bits)
{
    return guarded(ctx, [&]() {
        DevBuf<u32> k, v;
        k.reserve(n + 1, ctx->st); v.reserve(n + 1, ctx->st);
        CIPC_CUDA(cudaMemcpyAsync(k.p, keys, n * 4, cudaMemcpyHostToDevice, ctx->st));
        CIPC_CUDA(cudaMemcpyAsync(v.p, vals, n * 4, cudaMemcpyHostToDevice, ctx->st));
        device_radix_sort(k.p, v.p, n, bits, ctx->sortwk, ctx->st);
        CIPC_CUDA(cudaMemcpyAsync(keys, k.p, n * 4, cudaMemcpyDeviceToHost, ctx->st));
        CIPC_CUDA(cudaMemcpyAsync(vals, v.p, n * 4, cudaMemcpyDeviceToHost, ctx->st));
        CIPC_CUDA(cudaStreamSynchronize(ctx->st));
        return (int)CIPC_OK;
    });
}
int cipc_test_make_pd(cipc_ctx* ctx, double* H, int n, int count)
{
    return guarded(ctx, [&]() {
        if (n != 6 && n != 9 && n != 12) return (int)CIPC_ERR_ARG;
        DevBuf<double> d;
        const size_t tot = (size_t)count * n * n;
        d.reserve(tot + 1, ctx->st);
        CIPC_CUDA(cudaMemcpyAsync(d.p, H, tot * 8, cudaMemcpyHostToDevice, ctx->st));
        if (n == 12) CIPC_LAUNCH(k_test_make_pd<12>, div_up(count, 64), 64, 0, ctx->st, d.p, count);
        else if (n == 9) CIPC_LAUNCH(k_test_make_pd<9>, div_up(count, 64), 64, 0, ctx->st, d.p, count);
        else CIPC_LAUNCH(k_test_make_pd<6>, div_up(count, 64), 64, 0, ctx->st, d.p, count);
        CIPC_CUDA(cudaMemcpyAsync(H, d.p, tot * 8, cudaMemcpyDeviceToHost, ctx->st));
        CIPC_CUDA(cudaStreamSynchronize(ctx->st));
        return (int)CIPC_OK;
    });
}

} // extern "C"

// ========================================================================================= multi-device context (multidev.h)
namespace {
// copies n int4 records between devices on the destination's stream
void peer_copy(cipc_ctx* dst, int4* d, cipc_ctx* src, const int4* sp, size_t n)
{
    if (!n) return;
    CIPC_CUDA(cudaSetDevice(dst->dev));
    if (dst->dev == src->dev) CIPC_CUDA(cudaMemcpyAsync(d, sp, n * 16, cudaMemcpyDeviceToDevice, dst->st));
    else CIPC_CUDA(cudaMemcpyPeerAsync(d, dst->dev, sp, src->dev, n * 16, dst->st));
}
template <class F>
int multi_guard(cipc_ctx* ctx, F f)
{
    try { return f(); }
    catch (const std::exception& e) { ctx->err = e.what(); return CIPC_ERR_CUDA; }
}
int sub_err(cipc_ctx* ctx, cipc_ctx* s, int st)
{
    if (st) ctx->err = s->err;
    return st;
}
} // namespace

extern "C" int cipc_create_multi(int ndev, const int* devices, cipc_ctx** out)
{
    if (!out || ndev < 1 || ndev > MAX_RANKS || !devices) return CIPC_ERR_ARG;
    *out = nullptr;
    std::unique_ptr<cipc_ctx> c(new cipc_ctx());
    c->dev = devices[0];
    c->multi.reset(new cipc_multi());
    cipc_multi& m = *c->multi;
    m.n = ndev;
    for (int r = 0; r < ndev; ++r) {
        cipc_ctx* sub = nullptr;
        const int st = cipc_create(devices[r], r, ndev, &sub);
        if (st != CIPC_OK) { for (cipc_ctx* q : m.sub) cipc_destroy(q); return st; }
        m.sub.push_back(sub);
    }
    for (int a = 0; a < ndev; ++a) // peer access for the NVLink copies / loads between the ranks
        for (int b = 0; b < ndev; ++b) {
            if (devices[a] == devices[b]) continue;
            int can = 0;
            cudaDeviceCanAccessPeer(&can, devices[a], devices[b]);
            if (!can) { for (cipc_ctx* q : m.sub) cipc_destroy(q); fprintf(stderr, "cipc_b200: devices %d and %d have no peer access\n", devices[a], devices[b]); return CIPC_ERR_CUDA; }
            cudaSetDevice(devices[a]);
            const cudaError_t e = cudaDeviceEnablePeerAccess(devices[b], 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { for (cipc_ctx* q : m.sub) cipc_destroy(q); return CIPC_ERR_CUDA; }
            cudaGetLastError();
        }
    for (int r = 1; r < ndev; ++r) m.thr.emplace_back(new RankThread());
    m.chunk.assign(ndev + 1, 0);
    m.tripCount.assign(ndev, 0);
    cudaSetDevice(devices[0]);
    *out = c.release();
    return CIPC_OK;
}

static int multi_set_topology(cipc_ctx* ctx, int nV, int nBN, const int32_t* BN, int nBE, const int32_t* BE, int be_stride, int nBT, const int32_t* BT,
    int bt_stride, int nRod, const int32_t codim[2], const uint8_t* dbc, int nNnx, const int32_t* nnxPairs, const double* BNArea, const double* BEArea,
    const double* BTArea)
{
    cipc_multi& m = *ctx->multi;
    // the content check is the same on every rank: ask rank 0 first, fan the upload out only when something changed
    const u64 before = m.sub[0]->topoHash;
    int st = cipc_set_topology(m.sub[0], nV, nBN, BN, nBE, BE, be_stride, nBT, BT, bt_stride, nRod, codim, dbc, nNnx, nnxPairs, BNArea, BEArea, BTArea);
    if (st) return sub_err(ctx, m.sub[0], st);
    bool same = m.sub[0]->topoHash == before;
    for (int r = 1; r < m.n && same; ++r) same = m.sub[r]->topoHash == before && m.sub[r]->T.nV == nV;
    if (same) return CIPC_OK;
    for (int r = 1; r < m.n; ++r) {
        st = cipc_set_topology(m.sub[r], nV, nBN, BN, nBE, BE, be_stride, nBT, BT, bt_stride, nRod, codim, dbc, nNnx, nnxPairs, BNArea, BEArea, BTArea);
        if (st) return sub_err(ctx, m.sub[r], st);
    }
    ctx->T.nV = nV;
    return CIPC_OK;
}
// which: 0 positions, 1 rest positions, 2 search direction.  Rank 0 takes the host array (content-tagged: an unchanged array is
// not sent), the other ranks receive it over NVLink.
static int multi_upload(cipc_ctx* ctx, int which, const double* src, int stride_bytes)
{
    cipc_multi& m = *ctx->multi;
    return multi_guard(ctx, [&]() {
        cipc_ctx* c0 = m.sub[0];
        const u64 v0 = c0->xVersion, t0 = which == 0 ? c0->tagX : (which == 1 ? c0->tagX0 : c0->tagP);
        int st = which == 0 ? cipc_set_positions(c0, src, stride_bytes) : (which == 1 ? cipc_set_rest_positions(c0, src, stride_bytes) : cipc_set_search_dir(c0, src));
        if (st) return sub_err(ctx, c0, st);
        const u64 t1 = which == 0 ? c0->tagX : (which == 1 ? c0->tagX0 : c0->tagP);
        bool changed = t1 != t0 || t1 == 0 || c0->xVersion != v0;
        for (int r = 1; r < m.n && !changed; ++r) {
            cipc_ctx* c = m.sub[r];
            changed = (which == 0 ? c->tagX : (which == 1 ? c->tagX0 : c->tagP)) != t1;
        }
        if (!changed) return (int)CIPC_OK;
        CIPC_CUDA(cudaSetDevice(c0->dev));
        CIPC_CUDA(cudaStreamSynchronize(c0->st));
        const size_t bytes = (size_t)c0->T.nV * 32;
        for (int r = 1; r < m.n; ++r) {
            cipc_ctx* c = m.sub[r];
            need(c->T.nV == c0->T.nV && c->T.nV > 0, "topology not set");
            CIPC_CUDA(cudaSetDevice(c->dev));
            DevBuf<double4>& d = which == 0 ? c->X : (which == 1 ? c->X0 : c->P);
            const DevBuf<double4>& s0 = which == 0 ? c0->X : (which == 1 ? c0->X0 : c0->P);
            d.reserve(c->T.nV, c->st);
            CIPC_CUDA(cudaMemcpyPeerAsync(d.p, c->dev, s0.p, c0->dev, bytes, c->st));
            if (which == 0) { c->haveX = true; c->tagX = t1; ++c->xVersion; }
            else if (which == 1) {
                c->restLen2.reserve(std::max(c->T.nBE, 1), c->st);
                if (c->T.nBE) CIPC_LAUNCH(k_rest_len2, div_up(c->T.nBE, TB), TB, 0, c->st, c->X0.p, c->BE.p, c->T.nBE, c->restLen2.p);
                c->haveX0 = true; c->tagX0 = t1;
            }
            else { c->haveP = true; c->tagP = t1; }
        }
        CIPC_CUDA(cudaSetDevice(c0->dev));
        return (int)CIPC_OK;
    });
}
static int multi_constraint_set(cipc_ctx* ctx, int elastic, double dHat2, double thickness, int* nC_out)
{
    cipc_multi& m = *ctx->multi;
    return multi_guard(ctx, [&]() {
        const int N = m.n;
        std::vector<int> nr(N, 0);
        int st = m.run_all([&](int r) {
            const int e = cipc_constraint_set(m.sub[r], elastic, dHat2, thickness, &nr[r]);
            if (e == CIPC_OK) { cudaSetDevice(m.sub[r]->dev); cudaStreamSynchronize(m.sub[r]->st); }
            return e;
        });
        if (st) { for (cipc_ctx* s : m.sub) if (!s->err.empty()) ctx->err = s->err; return st; }
        cipc_ctx* c0 = m.sub[0];
        // ---- PP / PE stencils of all ranks -> device 0 -> merged with their multiplicities
        size_t nPP = 0, nPass = 0;
        for (int r = 0; r < N; ++r) { nPP += (size_t)nr[r] - m.sub[r]->nPassLast; nPass += m.sub[r]->nPassLast; }
        u32 nM = 0;
        std::vector<u32> ownCnt(N, 0), ownBase(N + 1, 0);
        CIPC_CUDA(cudaSetDevice(c0->dev));
        if (nPP) {
            c0->raw.reserve(nPP + 1, c0->st);
            RankOffsets ro;
            ro.n = N;
            size_t o = 0;
            for (int r = 0; r < N; ++r) {
                const size_t k = (size_t)nr[r] - m.sub[r]->nPassLast;
                ro.off[r] = (u32)o;
                peer_copy(c0, c0->raw.p + o, m.sub[r], m.sub[r]->cs.p + m.sub[r]->nPassLast, k);
                o += k;
            }
            ro.off[N] = (u32)o;
            CIPC_CUDA(cudaSetDevice(c0->dev));
            u32 nSlots = 1024;
            while (nSlots < 2 * nPP) nSlots <<= 1;
            c0->slots.reserve(nSlots, c0->st); c0->slotCnt.reserve(nSlots, c0->st);
            m.mergeTmp.reserve(nPP + 1, c0->st);
            m.ownerDev.reserve(2 * MAX_RANKS + 2, c0->st);
            CIPC_CUDA(cudaMemsetAsync(c0->slots.p, 0xff, (size_t)nSlots * 4, c0->st));
            CIPC_CUDA(cudaMemsetAsync(c0->slotCnt.p, 0, (size_t)nSlots * 4, c0->st));
            CIPC_CUDA(cudaMemsetAsync(m.ownerDev.p, 0, (2 * MAX_RANKS + 2) * 4, c0->st));
            CIPC_LAUNCH(k_dedup_insert_w, div_up(nPP, TB), TB, 0, c0->st, c0->raw.p, (u32)nPP, c0->slots.p, c0->slotCnt.p, nSlots - 1);
            // merged stencils grouped by owner rank: count, then emit at the owners' bases
            CIPC_LAUNCH(k_dedup_emit_owned<0>, div_up(nSlots, TB), TB, 0, c0->st, c0->raw.p, c0->slots.p, c0->slotCnt.p, nSlots, ro, (const u32*)nullptr,
                m.ownerDev.p, (int4*)nullptr);
            CIPC_CUDA(cudaMemcpyAsync(ownCnt.data(), m.ownerDev.p, (size_t)N * 4, cudaMemcpyDeviceToHost, c0->st));
            CIPC_CUDA(cudaStreamSynchronize(c0->st));
            for (int r = 0; r < N; ++r) ownBase[r + 1] = ownBase[r] + ownCnt[r];
            nM = ownBase[N];
            CIPC_CUDA(cudaMemcpyAsync(m.ownerDev.p + MAX_RANKS + 1, ownBase.data(), (size_t)(N + 1) * 4, cudaMemcpyHostToDevice, c0->st));
            CIPC_CUDA(cudaMemsetAsync(m.ownerDev.p, 0, (size_t)N * 4, c0->st));
            CIPC_LAUNCH(k_dedup_emit_owned<1>, div_up(nSlots, TB), TB, 0, c0->st, c0->raw.p, c0->slots.p, c0->slotCnt.p, nSlots, ro,
                (const u32*)(m.ownerDev.p + MAX_RANKS + 1), m.ownerDev.p, m.mergeTmp.p);
            CIPC_CUDA(cudaStreamSynchronize(c0->st));
        }
        // ---- global list G = [pass_0 | merged_0 | pass_1 | merged_1 | ...] (merged_r: the merged PP/PE stencils rank r owns),
        // re-cut into N contiguous chunks
        const size_t nG = nPass + nM;
        if (nG > 0x7fffffffull) throw std::runtime_error("constraint set exceeds the reference's 32-bit container sizes");
        struct Seg { cipc_ctx* c; const int4* p; size_t n, off; };
        std::vector<Seg> segs;
        size_t off = 0;
        for (int r = 0; r < N; ++r) {
            segs.push_back({m.sub[r], m.sub[r]->cs.p, m.sub[r]->nPassLast, off}); off += m.sub[r]->nPassLast;
            segs.push_back({c0, m.mergeTmp.p + ownBase[r], ownCnt[r], off}); off += ownCnt[r];
        }
        for (int r = 0; r <= N; ++r) m.chunk[r] = nG * (size_t)r / (size_t)N;
        for (int r = 0; r < N; ++r) {
            cipc_ctx* c = m.sub[r];
            const size_t a = m.chunk[r], b = m.chunk[r + 1];
            CIPC_CUDA(cudaSetDevice(c->dev));
            c->raw.reserve(b - a + 1, c->st); // the chunk is assembled in `raw` and swapped in below
            for (const Seg& sg : segs) {
                const size_t lo = std::max(a, sg.off), hi = std::min(b, sg.off + sg.n);
                if (hi > lo) peer_copy(c, c->raw.p + (lo - a), sg.c, sg.p + (lo - sg.off), hi - lo);
            }
        }
        for (int r = 0; r < N; ++r) { CIPC_CUDA(cudaSetDevice(m.sub[r]->dev)); CIPC_CUDA(cudaStreamSynchronize(m.sub[r]->st)); }
        const double dHat = std::sqrt(dHat2) + thickness, dHat2o = dHat * dHat;
        for (int r = 0; r < N; ++r) {
            cipc_ctx* c = m.sub[r];
            const u32 n = (u32)(m.chunk[r + 1] - m.chunk[r]);
            CIPC_CUDA(cudaSetDevice(c->dev));
            std::swap(c->cs.p, c->raw.p); std::swap(c->cs.cap, c->raw.cap);
            c->nC = n;
            c->info.reserve((size_t)n + 1, c->st);
            if (n) CIPC_LAUNCH(k_fill_info, div_up(n, TB), TB, 0, c->st, c->info.p, n, 1.0, dHat2o);
            c->infoUniform = true; c->infoW = 1.0; c->infoD = dHat2o;
            c->ctr["constraints"] = n;
        }
        CIPC_CUDA(cudaSetDevice(c0->dev));
        if (nC_out) *nC_out = (int)nG;
        return (int)CIPC_OK;
    });
}
static int multi_get_constraints(cipc_ctx* ctx, int32_t* cs, double* info, int info_stride_bytes)
{
    cipc_multi& m = *ctx->multi;
    return m.run_all([&](int r) {
        const size_t a = m.chunk[r];
        return sub_err(ctx, m.sub[r], cipc_get_constraints_strided(m.sub[r], cs ? cs + 4 * a : nullptr,
            info ? (double*)((char*)info + a * (size_t)info_stride_bytes) : nullptr, info_stride_bytes));
    });
}
static int multi_set_constraints(cipc_ctx* ctx, const int32_t* cs, const double* info, int info_stride_bytes, int nC)
{
    cipc_multi& m = *ctx->multi;
    if (nC < 0) return CIPC_ERR_ARG;
    for (int r = 0; r <= m.n; ++r) m.chunk[r] = (size_t)nC * (size_t)r / (size_t)m.n;
    return m.run_all([&](int r) {
        const size_t a = m.chunk[r], n = m.chunk[r + 1] - a;
        return sub_err(ctx, m.sub[r], cipc_set_constraints_strided(m.sub[r], cs ? cs + 4 * a : nullptr,
            info ? (const double*)((const char*)info + a * (size_t)info_stride_bytes) : nullptr, info_stride_bytes, (int)n));
    });
}
static int multi_barrier_energy(cipc_ctx* ctx, int elastic, double dHat2, const double kappa[3], double thickness, double* E)
{
    cipc_multi& m = *ctx->multi;
    std::vector<double> e(m.n, 0.0);
    const int st = m.run_all([&](int r) { return sub_err(ctx, m.sub[r], cipc_barrier_energy(m.sub[r], elastic, dHat2, kappa, thickness, &e[r])); });
    if (st) return st;
    double s = 0;
    for (int r = 0; r < m.n; ++r) s += e[r]; // rank order: reproducible
    *E += s;
    return CIPC_OK;
}
static int multi_barrier_gradient(cipc_ctx* ctx, int elastic, double dHat2, const double kappa[3], double thickness, double* g, int stride)
{
    cipc_multi& m = *ctx->multi;
    if (stride < 24 || stride % 8) return CIPC_ERR_ARG;
    return multi_guard(ctx, [&]() {
        int st = m.run_all([&](int r) {
            const int e = cipc_barrier_gradient_dev(m.sub[r], elastic, dHat2, kappa, thickness);
            if (e == CIPC_OK) { cudaSetDevice(m.sub[r]->dev); cudaStreamSynchronize(m.sub[r]->st); }
            return sub_err(ctx, m.sub[r], e);
        });
        if (st) return st;
        cipc_ctx* c = m.sub[0];
        CIPC_CUDA(cudaSetDevice(c->dev));
        const size_t n = (size_t)c->T.nV;
        if (m.n > 1) { // device 0 sums the ranks' vectors with peer loads over NVLink
            PeerPtrs pp;
            pp.n = m.n - 1;
            for (int r = 1; r < m.n; ++r) pp.p[r - 1] = m.sub[r]->g.p;
            CIPC_LAUNCH(k_sum_peers, 592, 256, 0, c->st, c->g.p, pp, 3 * n);
        }
        double* h = (double*)c->pin.reserve(n * 24);
        CIPC_CUDA(cudaMemcpyAsync(h, c->g.p, n * 24, cudaMemcpyDeviceToHost, c->st));
        CIPC_CUDA(cudaStreamSynchronize(c->st));
        const size_t sd = stride / 8, CH = 1 << 14;
        HostPool::get().for_each((n + CH - 1) / CH, [&](size_t k) { // nodeAttr.g += (IPC.h:1034-1042)
            for (size_t v = k * CH, e = std::min(n, v + CH); v < e; ++v) { g[v * sd] += h[3 * v]; g[v * sd + 1] += h[3 * v + 1]; g[v * sd + 2] += h[3 * v + 2]; }
        });
        return (int)CIPC_OK;
    });
}
static int multi_barrier_hessian(cipc_ctx* ctx, int elastic, double dHat2, const double kappa[3], double thickness, int projectSPD, int64_t* nTrip, bool merged)
{
    cipc_multi& m = *ctx->multi;
    const int st = m.run_all([&](int r) {
        int64_t n = 0;
        const int e = merged ? cipc_barrier_hessian_merged(m.sub[r], elastic, dHat2, kappa, thickness, projectSPD, &n)
                             : cipc_barrier_hessian(m.sub[r], elastic, dHat2, kappa, thickness, projectSPD, &n);
        m.tripCount[r] = n;
        return sub_err(ctx, m.sub[r], e);
    });
    if (st) return st;
    int64_t tot = 0;
    for (int r = 0; r < m.n; ++r) tot += m.tripCount[r];
    if (nTrip) *nTrip = tot;
    return CIPC_OK;
}
static int multi_get_triplets(cipc_ctx* ctx, cipc_triplet* out)
{
    cipc_multi& m = *ctx->multi;
    std::vector<int64_t> off(m.n + 1, 0);
    for (int r = 0; r < m.n; ++r) off[r + 1] = off[r] + m.tripCount[r];
    return m.run_all([&](int r) { return m.tripCount[r] ? sub_err(ctx, m.sub[r], cipc_get_triplets(m.sub[r], out + off[r])) : (int)CIPC_OK; });
}
static int multi_step_size(cipc_ctx* ctx, int elastic, double thickness, double* step)
{
    cipc_multi& m = *ctx->multi;
    std::vector<double> a(m.n, *step);
    const int st = m.run_all([&](int r) { return sub_err(ctx, m.sub[r], cipc_step_size(m.sub[r], elastic, thickness, &a[r])); });
    if (st) return st;
    double v = a[0];
    for (int r = 1; r < m.n; ++r) v = std::min(v, a[r]);
    *step = v;
    return CIPC_OK;
}
static int multi_min_dist2(cipc_ctx* ctx, double thickness, double* dist2, double* minDist2)
{
    cipc_multi& m = *ctx->multi;
    if (m.chunk[m.n] == 0) return CIPC_OK; // empty set: outputs untouched (IPC.h:2253)
    std::vector<double> v(m.n, 1e300);
    const int st = m.run_all([&](int r) {
        if (m.chunk[r + 1] == m.chunk[r]) return (int)CIPC_OK;
        return sub_err(ctx, m.sub[r], cipc_min_dist2(m.sub[r], thickness, dist2 ? dist2 + m.chunk[r] : nullptr, &v[r]));
    });
    if (st) return st;
    double mn = v[0];
    for (int r = 1; r < m.n; ++r) mn = std::min(mn, v[r]);
    *minDist2 = mn;
    return CIPC_OK;
}
