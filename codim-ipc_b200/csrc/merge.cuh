// merge.cuh -- device-side merge of the per-stencil Hessian blocks into the unique 3x3 blocks of the (symmetric) contact
// matrix, upper block triangle only: what Eigen's setFromTriplets (Math/CSR_MATRIX.h:49-56, fed by
// Shell/INC_POTENTIAL.h:374-382) does with the 144/81/36 triplets of every stencil -- sum the duplicates -- done in HBM
// before anything crosses PCIe.  At 1M triangles 7.2M stencils emit 61M upper blocks that collapse to 7.3M unique ones:
// the host receives 0.6 GB instead of 14.5 GB and the solver's assembly sorts 119M triplets instead of 909M.
//
// Block stream (written by the fused Hessian kernels in block mode, cipc_b200.cu): stencil i with vertices v_0..v_{nb-1}
// owns blocks [off[i], off[i] + nb(nb+1)/2), one per local pair I <= J in row-major order, 9 doubles each (row-major 3x3),
// stored for the ORIENTED vertex pair (min(v_I, v_J), max(v_I, v_J)) -- the block is transposed when v_I > v_J.
//
// Merge (no global sort; keys come from the stencil list, the block values are only gathered once at the end):
//   1. k_blk_count    per stencil: rowCnt[min vertex] += 1 per pair                       (L2-resident atomics)
//   2. scan           rowCnt -> rowStart
//   3. k_blk_scatter  per stencil: ent[rowCur[row]++] = (col << 32 | block id)            (8 B per block)
//   4. k_row_sort     one CTA per 16 rows: every warp sorts whole rows by (col, block id) in shared memory (a CTA-wide sort
//                     when a row is long; in global memory when the group exceeds the buffer); unique (row, col) runs per group
//   5. scan           unique runs per group -> first unique index of the group (nV / 16 entries)
//   6. k_blk_uniq     second pass over the sorted groups: (row, col, first entry) of every run
//   7. k_blk_sum      nine threads per unique block: sum of the run in (block id) order -- a reproducible sum
// Included by cipc_b200.cu after decode() / Stencil are defined.
#pragma once

namespace cipc {

__device__ __forceinline__ int stencil_nb(const int4 c) { return (c.x >= 0 || c.w >= 0) ? 4 : (c.z >= 0 ? 3 : 2); }
__host__ __device__ __forceinline__ int upper_pairs(int nb) { return nb * (nb + 1) / 2; }

// blocks per stencil (10 / 6 / 3): scanned into the block offsets
__global__ void k_blk_sizes(const int4* __restrict__ cs, u32 n, u32* sz)
{
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) sz[i] = (u32)upper_pairs(stencil_nb(cs[i]));
}
template <bool SCATTER>
__global__ void __launch_bounds__(256) k_blk_rows(const int4* __restrict__ cs, const u32* __restrict__ off, u32 n, u32* __restrict__ rowCnt,
    u64* __restrict__ ent)
{
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const Stencil s = decode(cs[i]);
    const int nb = (s.kind == K_PE) ? 3 : (s.kind == K_PP ? 2 : 4);
    u32 b = SCATTER ? off[i] : 0u;
    for (int I = 0; I < nb; ++I)
        for (int J = I; J < nb; ++J, ++b) {
            const u32 r = (u32)min(s.v[I], s.v[J]), c = (u32)max(s.v[I], s.v[J]);
            const u32 pos = atomicAdd(&rowCnt[r], 1u);
            if (SCATTER) ent[pos] = ((u64)c << 32) | b;
        }
}
// Sorts the entries of MRS_ROWS consecutive rows (contiguous in `ent`) by (col, block id) inside every row and counts the
// group's unique (row, col) runs.  The comparator network is the all-ascending bitonic variant (the first step of every
// merge mirrors the partner index), so an arbitrary length works with virtual +inf padding: comparators that reach past n
// are skipped.  Fast path (every row of the group has at most MRS_WCAP entries -- a cloth vertex has ~130): each warp sorts
// whole rows on its own in shared memory, __syncwarp only.  Otherwise the CTA sorts the whole group at once with the row in
// the key's top bits -- in shared memory up to MRS_CAP entries, in place in global memory beyond.
constexpr int MRS_ROWS = 16, MRS_CAP = 4096, MRS_BT = 256, MRS_WCAP = MRS_CAP / (MRS_BT / 32);
template <class Sync>
__device__ __forceinline__ void bitonic_asc(u64* buf, u32 n, u32 tid, u32 nthr, Sync sync)
{
    auto cx = [&](u32 i, u32 p) {
        if (p < n) {
            const u64 a = buf[i], b = buf[p];
            if (a > b) { buf[i] = b; buf[p] = a; }
        }
    };
    u32 np2 = 1;
    while (np2 < n) np2 <<= 1;
    const u32 nCmp = np2 >> 1; // comparators per step
    for (u32 k = 2; k <= np2; k <<= 1) {
        const u32 hk = k >> 1;
        for (u32 t = tid; t < nCmp; t += nthr) { // mirror step: i in the lower half of its k-block, partner mirrored
            const u32 blk = t / hk, pos = t - blk * hk, i = blk * k + pos;
            if (i < n) cx(i, blk * k + (k - 1 - pos));
        }
        sync();
        for (u32 j = k >> 2; j > 0; j >>= 1) {
            for (u32 t = tid; t < nCmp; t += nthr) {
                const u32 i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                if (i < n) cx(i, i | j);
            }
            sync();
        }
    }
}
__global__ void __launch_bounds__(MRS_BT) k_row_sort(const u32* __restrict__ rowStart, int nV, u64* __restrict__ ent, u32* __restrict__ grpUniq,
    u32* __restrict__ nRowsUsed)
{
    __shared__ u64 sh[MRS_CAP];
    __shared__ u32 rs[MRS_ROWS + 1];
    __shared__ u32 wcnt[MRS_BT / 32];
    const int r0 = blockIdx.x * MRS_ROWS, nr = min(MRS_ROWS, nV - r0);
    const u32 lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    if ((int)threadIdx.x <= nr) rs[threadIdx.x] = rowStart[r0 + threadIdx.x];
    __syncthreads();
    const u32 s0 = rs[0], n = rs[nr] - s0;
    if (n == 0) { if (threadIdx.x == 0) grpUniq[blockIdx.x] = 0; return; }
    u32 maxLen = 0, used = 0;
    for (int k = 0; k < nr; ++k) { maxLen = max(maxLen, rs[k + 1] - rs[k]); used += rs[k + 1] > rs[k]; }
    if (threadIdx.x == 0) atomicAdd(nRowsUsed, used);
    u32 uniq = 0; // per thread, summed below
    if (maxLen <= (u32)MRS_WCAP) {
        u64* buf = sh + warp * MRS_WCAP;
        for (int k = (int)warp; k < nr; k += MRS_BT / 32) {
            const u32 a = rs[k], m = rs[k + 1] - a;
            for (u32 i = lane; i < m; i += 32) buf[i] = ent[a + i];
            __syncwarp();
            bitonic_asc(buf, m, lane, 32u, [] { __syncwarp(); });
            for (u32 i = lane; i < m; i += 32) {
                const u64 e = buf[i];
                uniq += (i == 0 || (e >> 32) != (buf[i - 1] >> 32)) ? 1u : 0u;
                ent[a + i] = e;
            }
            __syncwarp();
        }
    }
    else {
        u64* buf = n <= (u32)MRS_CAP ? sh : ent + s0;
        for (u32 i = threadIdx.x; i < n; i += MRS_BT) {
            int lr = 0;
            while (s0 + i >= rs[lr + 1]) ++lr;
            buf[i] = ent[s0 + i] | ((u64)lr << 60);
        }
        __syncthreads();
        bitonic_asc(buf, n, threadIdx.x, (u32)MRS_BT, [] { __syncthreads(); });
        for (u32 i = threadIdx.x; i < n; i += MRS_BT) uniq += (i == 0 || (buf[i] >> 32) != (buf[i - 1] >> 32)) ? 1u : 0u; // (row, col) changes
        __syncthreads(); // buf may alias ent: strip the row tags only after every comparison with the neighbour is done
        for (u32 i = threadIdx.x; i < n; i += MRS_BT) ent[s0 + i] = buf[i] & 0x0fffffffffffffffULL;
    }
    for (int o = 16; o > 0; o >>= 1) uniq += __shfl_down_sync(0xffffffffu, uniq, o);
    if (lane == 0) wcnt[warp] = uniq;
    __syncthreads();
    if (threadIdx.x == 0) {
        u32 t = 0;
        for (int w = 0; w < MRS_BT / 32; ++w) t += wcnt[w];
        grpUniq[blockIdx.x] = t;
    }
}
// second pass over the sorted groups: (row, col, first entry) of every unique run, at grpBase[group] + rank inside the group
__global__ void __launch_bounds__(MRS_BT) k_blk_uniq(const u32* __restrict__ rowStart, int nV, const u64* __restrict__ ent, const u32* __restrict__ grpBase,
    u32* __restrict__ urow, u32* __restrict__ ucol, u32* __restrict__ ustart)
{
    __shared__ u32 rs[MRS_ROWS + 1];
    const int r0 = blockIdx.x * MRS_ROWS, nr = min(MRS_ROWS, nV - r0);
    if ((int)threadIdx.x <= nr) rs[threadIdx.x] = rowStart[r0 + threadIdx.x];
    __syncthreads();
    const u32 s0 = rs[0], n = rs[nr] - s0;
    u32 base = grpBase[blockIdx.x];
    for (u32 c0 = 0; c0 < n; c0 += MRS_BT) { // uniform trip count: block_excl_scan synchronises
        const u32 i = c0 + threadIdx.x;
        u32 head = 0, col = 0;
        int lr = 0;
        if (i < n) {
            while (s0 + i >= rs[lr + 1]) ++lr;
            col = (u32)(ent[s0 + i] >> 32);
            head = (s0 + i == rs[lr] || col != (u32)(ent[s0 + i - 1] >> 32)) ? 1u : 0u;
        }
        u32 tot;
        const u32 rank = block_excl_scan<MRS_BT>(head, tot);
        if (head) {
            const u32 u = base + rank;
            urow[u] = (u32)(r0 + lr); ucol[u] = col; ustart[u] = s0 + i;
        }
        base += tot;
    }
}
__global__ void __launch_bounds__(288) k_blk_sum(const double* __restrict__ blkVal, const u64* __restrict__ ent, const u32* __restrict__ ustart, u32 nU,
    u32 nEnt, double* __restrict__ uval)
{
    const u32 u = blockIdx.x * 32u + threadIdx.x / 9u, k = threadIdx.x % 9u;
    if (u >= nU) return;
    const u32 s0 = ustart[u], s1 = (u + 1 < nU) ? ustart[u + 1] : nEnt;
    double acc = 0.0;
    for (u32 i = s0; i < s1; ++i) acc += blkVal[(size_t)(u32)ent[i] * 9 + k];
    uval[(size_t)u * 9 + k] = acc;
}

// ---- block-mode output of the Hessian kernels
// pair index b (row-major upper triangle of an NB x NB grid) -> (I, J)
template <int NB>
__device__ __forceinline__ void upper_pair_ij(int b, int& I, int& J)
{
    I = 0;
    int rowLen = NB;
    while (b >= rowLen) { b -= rowLen; --rowLen; ++I; }
    J = I + b;
}
// Block-mode twin of warp_expand_stencils: lane l owns the doubles e = l + 32 j of every stencil's block run
// (9 nb(nb+1)/2 doubles, contiguous), so a warp writes 256 contiguous bytes per store instruction.
// bit b of the mask: pair b = (I, J) of the upper triangle has v_I > v_J, i.e. its block is stored transposed (for the
// (min vertex, max vertex) orientation).  Computed once per stencil by the factoring thread, kept in header word 6.
template <int NB>
__device__ __forceinline__ int swap_mask(const int* v)
{
    int m = 0, b = 0;
#pragma unroll
    for (int I = 0; I < NB; ++I)
#pragma unroll
        for (int J = I; J < NB; ++J, ++b) m |= (v[I] > v[J]) ? (1 << b) : 0;
    return m;
}
template <int NB, int NY, int YS, bool SIGNED>
__device__ __forceinline__ void warp_expand_blocks(const double* sY, const int* sH, u32 wq0, u32 g, u32 lane, double* __restrict__ blk)
{
    constexpr int NN = 3 * NB, PER = 9 * (NB * (NB + 1) / 2), NJ = (PER + 31) / 32;
    int bI[NJ], r0[NJ], c0[NJ], r1[NJ], c1[NJ]; // pair index; (row, col) inside the stencil, plain and transposed
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
        const int e = (int)lane + 32 * j;
        const int b = (e < PER ? e : 0) / 9, k = (e < PER ? e : 0) - 9 * b;
        int I, J;
        upper_pair_ij<NB>(b, I, J);
        const int ra = k / 3, rc = k - 3 * ra;
        bI[j] = b;
        r0[j] = 3 * I + ra; c0[j] = 3 * J + rc;
        r1[j] = 3 * J + ra; c1[j] = 3 * I + rc;
    }
    const u32 wn = (wq0 < g) ? min(32u, g - wq0) : 0u;
    for (u32 qq = 0; qq < wn; ++qq) {
        const int* h = sH + (wq0 + qq) * 8;
        const u32 o = (u32)h[0];
        if (o == 0xffffffffu) continue;
        const double* y = sY + (wq0 + qq) * YS;
        const bool neg = SIGNED && h[5] != 0;
        const int mask = h[6];
        double* dst = blk + (size_t)o * 9;
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            const int e = (int)lane + 32 * j;
            if (e < PER) {
                const bool sw = (mask >> bI[j]) & 1;
                const int r = sw ? r1[j] : r0[j], c = sw ? c1[j] : c0[j];
                double v = 0.0;
#pragma unroll
                for (int k = 0; k < NY; ++k) v += y[k * NN + r] * y[k * NN + c];
                __stcs(dst + e, (SIGNED && neg) ? -v : v);
            }
        }
    }
}
// dense n x n block H (row-major, n = 3 nb) of one stencil -> its upper block run (dense path, unprojected path)
__device__ __forceinline__ void store_blocks_dense(const double* H, int nb, const int* vids, double* __restrict__ dst)
{
    const int nn = 3 * nb;
    int b = 0;
    for (int I = 0; I < nb; ++I)
        for (int J = I; J < nb; ++J, ++b) {
            const bool sw = vids[I] > vids[J];
            const int R = sw ? J : I, C = sw ? I : J;
            for (int a = 0; a < 3; ++a)
                for (int c = 0; c < 3; ++c) dst[b * 9 + a * 3 + c] = H[(3 * R + a) * nn + 3 * C + c];
        }
}

} // namespace cipc
