// boundary.cuh -- boundary-primitive construction on the device.
// Reference: Find_Surface_Primitives_And_Compute_Area (Utils/MESHIO.h:768-834) -- a std::map<VECTOR<int,2>, T> over the
// 3T directed triangle edges, rebuilt every time step (Shell/IMPLICIT_EULER.h:224-243: "TODO: only once"), ~2 s at 1M
// triangles -- followed by the seg / rod / particle appends of Shell/IMPLICIT_EULER.h:245-276.
//
// Semantics reproduced exactly (index lists bit-identical, areas with the reference's operation order):
//   boundaryTri  = the triangles in order;  BTArea = A / 2,  A = 0.5 |(v1 - v0) x (v2 - v0)|
//   boundaryEdge = one entry per undirected edge {a, b}, oriented like the FIRST directed edge that mentions it (visiting
//                  order: triangle by triangle, (v0,v1), (v1,v2), (v2,v0)), listed in lexicographic order of the oriented
//                  pair (std::map order).  Its area is folded over the edge's events in visiting order: an event with the
//                  stored orientation ASSIGNS A/3 (map[(a,b)] = ..., MESHIO.h:789,796,803), an event with the opposite
//                  orientation ADDS A/3;  BEArea = that / 2
//   boundaryNode = ascending vertices whose accumulated A/3 (added triangle by triangle) is non-zero;  BNArea = that sum
// Device algorithm: the 3T events are ordered by stable LSD radix sorts (own device_radix_sort: by the larger vertex, then by
// the smaller one; the event id breaks ties = visiting order), one thread per run folds its events sequentially, a second
// pair of sorts puts the oriented edges in lexicographic order; node sums are folded the same way over events sorted by vertex.
// Included by cipc_b200.cu.
#pragma once

namespace cipc {

__global__ void k_bd_tri_area(const double4* __restrict__ X, const int4* __restrict__ tri, int nT, double* __restrict__ BTArea, double* __restrict__ third)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nT) return;
    const int4 q = tri[t];
    const xv3 v0 = ldx(X, q.x), v1 = ldx(X, q.y), v2 = ldx(X, q.z);
    const xd A = xd(0.5) * xd(sqrt(norm2(cross(v1 - v0, v2 - v0)).v)); // MESHIO.h:786
    third[t] = (A / xd(3.0)).v;
    BTArea[t] = (A / xd(2.0)).v; // MESHIO.h:811
}
// event e = 3 t + k: directed edge (tri[t][k], tri[t][(k+1)%3]) and vertex tri[t][k]
__device__ __forceinline__ int tri_v(const int4 q, int k) { return k == 0 ? q.x : (k == 1 ? q.y : q.z); }
__global__ void k_bd_events(const int4* __restrict__ tri, int nT, u32* __restrict__ evLo, u32* __restrict__ evHi, u32* __restrict__ evV)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= 3 * nT) return;
    const int t = e / 3, k = e - 3 * t;
    const int4 q = tri[t];
    const int a = tri_v(q, k), b = tri_v(q, (k + 1) % 3);
    evLo[e] = (u32)min(a, b); evHi[e] = (u32)max(a, b); evV[e] = (u32)a;
}
// heads of the runs of equal (k1[ids], k2[ids]) in sorted order (k2 may be null)
__global__ void k_bd_heads(const u32* __restrict__ k1, const u32* __restrict__ k2, const u32* __restrict__ ids, u32 n, u32* __restrict__ heads)
{
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    bool h = i == 0;
    if (!h) {
        const u32 a = ids[i], b = ids[i - 1];
        h = k1[a] != k1[b] || (k2 && k2[a] != k2[b]);
    }
    heads[i] = h ? 1u : 0u;
}
__global__ void k_bd_edge_fold(const int4* __restrict__ tri, const double* __restrict__ third, const u32* __restrict__ evLo, const u32* __restrict__ evHi,
    const u32* __restrict__ ids, const u32* __restrict__ heads, const u32* __restrict__ headScan, u32 n, u32* __restrict__ ueA, u32* __restrict__ ueB,
    double* __restrict__ ueVal)
{
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || !heads[i]) return;
    const u32 e0 = ids[i];
    const int t0 = (int)(e0 / 3u), k0 = (int)(e0 - 3u * (u32)t0);
    const int4 q0 = tri[t0];
    const int a0 = tri_v(q0, k0), b0 = tri_v(q0, (k0 + 1) % 3);
    double val = third[t0];
    const u32 lo = evLo[e0], hi = evHi[e0];
    for (u32 j = i + 1; j < n; ++j) {
        const u32 e = ids[j];
        if (evLo[e] != lo || evHi[e] != hi) break;
        const int t = (int)(e / 3u), k = (int)(e - 3u * (u32)t);
        const int a = tri_v(tri[t], k);
        if (a == a0) val = third[t];                 // same orientation: the reference's map[(a,b)] = A/3 overwrites
        else val = __dadd_rn(val, third[t]);         // opposite orientation: finder->second += A/3
    }
    const u32 u = headScan[i];
    ueA[u] = (u32)a0; ueB[u] = (u32)b0; ueVal[u] = val;
}
__global__ void k_bd_edge_emit(const u32* __restrict__ ueA, const u32* __restrict__ ueB, const double* __restrict__ ueVal, const u32* __restrict__ ids, u32 n,
    int2* __restrict__ BE, double* __restrict__ BEArea)
{
    const u32 j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const u32 u = ids[j];
    BE[j] = make_int2((int)ueA[u], (int)ueB[u]);
    BEArea[j] = ueVal[u] / 2.0; // MESHIO.h:818
}
__global__ void k_bd_node_fold(const double* __restrict__ third, const u32* __restrict__ evV, const u32* __restrict__ ids, const u32* __restrict__ heads,
    const u32* __restrict__ headScan, u32 n, u32* __restrict__ nodeV, double* __restrict__ nodeSum, u32* __restrict__ nodeKeep)
{
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || !heads[i]) return;
    const u32 v = evV[ids[i]];
    double s = 0.0; // std::vector<T> isBoundaryNode(X.size, 0) ... += A/3 per incident triangle, in triangle order (MESHIO.h:806-808)
    for (u32 j = i; j < n; ++j) {
        const u32 e = ids[j];
        if (evV[e] != v) break;
        s = __dadd_rn(s, third[e / 3u]);
    }
    const u32 u = headScan[i];
    nodeV[u] = v; nodeSum[u] = s; nodeKeep[u] = s != 0.0 ? 1u : 0u; // MESHIO.h:822: if (isBoundaryNode[vI])
}
__global__ void k_bd_node_emit(const u32* __restrict__ nodeV, const double* __restrict__ nodeSum, const u32* __restrict__ keep, const u32* __restrict__ keepScan,
    u32 n, int* __restrict__ BN, double* __restrict__ BNArea)
{
    const u32 u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= n || !keep[u]) return;
    const u32 o = keepScan[u];
    BN[o] = (int)nodeV[u]; BNArea[o] = nodeSum[u];
}

} // namespace cipc
