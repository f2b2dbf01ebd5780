// csr.cuh -- device-side assembly of the Hessian triplet stream into CSR.
// Reference: Math/CSR_MATRIX.h:49-56 Construct_From_Triplet = Eigen setFromTriplets (duplicates summed, inner indices
// sorted), fed by Shell/INC_POTENTIAL.h:382-394.  The contact / friction triplets are whole 3x3 blocks on (vertex,
// vertex) pairs, so the assembly works at block granularity -- 9x fewer keys than scalar entries:
//   1. k_trip_to_blocks : triplet stream of one Hessian -> 3x3 blocks (9 doubles) + (row vertex, col vertex) keys; several
//                         Hessians (barrier, friction) can be appended before finishing
//   2. two stable LSD radix sorts (by column vertex, then by row vertex) of (key, block id)
//   3. head flags + scan -> unique blocks; blocks per block-row -> scan -> block row pointer
//   4. k_csr_emit       : nine threads per unique block sum its run in sorted (= deterministic) order and write the
//                         three scalar rows' (col, value) entries; k_csr_rowptr writes the scalar row pointer
// Included by cipc_b200.cu (block_dim_of / cipc_triplet / u32 live there).
#pragma once

namespace cipc {

// one warp per stencil: the stencil's n x n triplets (row-major, contiguous) -> nb x nb blocks at block offset off[i]
__global__ void __launch_bounds__(256) k_trip_to_blocks(const cipc_triplet* __restrict__ trip, const int4* __restrict__ cs,
    const u32* __restrict__ off, u32 nSt, size_t blkBase, double* __restrict__ blkVal, u32* __restrict__ keyI, u32* __restrict__ keyJ)
{
    const u32 i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31u;
    if (i >= nSt) return;
    const int nn = block_dim_of(cs[i]), nb = nn / 3, per = nn * nn;
    const size_t o = off[i];
    const int4* src = reinterpret_cast<const int4*>(trip + o * 9);
    for (int t = (int)lane; t < per; t += 32) {
        const int4 q = src[t];
        const int r = t / nn, c = t - r * nn;
        const int I = r / 3, a = r - 3 * I, J = c / 3, b = c - 3 * J;
        const size_t e = blkBase + o + (size_t)(I * nb + J);
        blkVal[e * 9 + (size_t)(a * 3 + b)] = __longlong_as_double(((long long)(u32)q.w << 32) | (long long)(u32)q.z);
        if (a == 0 && b == 0) { keyI[e] = (u32)q.x / 3u; keyJ[e] = (u32)q.y / 3u; }
    }
}
__global__ void k_iota_copy(const u32* __restrict__ src, u32* __restrict__ key, u32* __restrict__ ids, size_t n)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { key[i] = src[i]; ids[i] = (u32)i; }
}
__global__ void k_gather_u32(const u32* __restrict__ src, const u32* __restrict__ ids, u32* __restrict__ out, size_t n)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = src[ids[i]];
}
// sorted entry i: row = rowS[i], col = colOrig[ids[i]]
__global__ void k_csr_heads(const u32* __restrict__ rowS, const u32* __restrict__ colOrig, const u32* __restrict__ ids, size_t n,
    u32* __restrict__ colS, u32* __restrict__ heads)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const u32 c = colOrig[ids[i]];
    colS[i] = c;
    heads[i] = (i == 0 || rowS[i] != rowS[i - 1] || c != colOrig[ids[i - 1]]) ? 1u : 0u;
}
__global__ void k_csr_unique(const u32* __restrict__ rowS, const u32* __restrict__ colS, const u32* __restrict__ heads,
    const u32* __restrict__ headScan, size_t n, u32* __restrict__ urow, u32* __restrict__ ucol, u32* __restrict__ ustart, u32* __restrict__ browCnt)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || !heads[i]) return;
    const u32 u = headScan[i];
    urow[u] = rowS[i]; ucol[u] = colS[i]; ustart[u] = (u32)i;
    atomicAdd(&browCnt[rowS[i]], 1u);
}
__global__ void k_csr_rowptr(const u32* __restrict__ browPtr, int nV, u32 nU, int* __restrict__ rowPtr)
{
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v > nV) return;
    if (v == nV) { rowPtr[3 * nV] = (int)(9u * nU); return; }
    const u32 b0 = browPtr[v], m = browPtr[v + 1] - b0;
    for (int a = 0; a < 3; ++a) rowPtr[3 * v + a] = (int)(9u * b0 + (u32)a * 3u * m);
}
// nine threads per unique block (one per entry of the 3x3): the 72-byte blocks of a run are read with contiguous
// 8-byte loads across the nine threads, the run is walked in sorted order (reproducible sum)
__global__ void __launch_bounds__(288) k_csr_emit(const double* __restrict__ blkVal, const u32* __restrict__ ids, const u32* __restrict__ urow,
    const u32* __restrict__ ucol, const u32* __restrict__ ustart, u32 nU, size_t nBlk, const u32* __restrict__ browPtr, int* __restrict__ colIdx,
    double* __restrict__ val)
{
    const u32 u = blockIdx.x * 32u + threadIdx.x / 9u, k = threadIdx.x % 9u;
    if (u >= nU) return;
    const size_t s0 = ustart[u], s1 = (u + 1 < nU) ? (size_t)ustart[u + 1] : nBlk;
    double acc = 0.0;
    for (size_t i = s0; i < s1; ++i) acc += blkVal[(size_t)ids[i] * 9 + k];
    const u32 vi = urow[u], vj = ucol[u];
    const u32 b0 = browPtr[vi], m = browPtr[vi + 1] - b0, t = u - b0;
    const u32 a = k / 3u, b = k - 3u * a;
    const size_t o = (size_t)9 * b0 + (size_t)a * 3 * m + (size_t)3 * t + b;
    colIdx[o] = (int)(3u * vj + b);
    val[o] = acc;
}

} // namespace cipc
