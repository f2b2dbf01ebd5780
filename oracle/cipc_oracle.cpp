// TEST INFRASTRUCTURE -- CPU oracle for the C-IPC contact hot path ("port" of the reference).
//
// Restates, Eigen/Kokkos/Cabana-free, what the reference's 3-D, T=double contact code computes:
//   Library/Grid/SPATIAL_HASH.h   (static build :28-213, queries :215-381, swept build :432-622,
//                                  id queries :624-690, voxel index :693-708)
//   Library/FEM/IPC.h             (Compute_Constraint_Set :19-740, Compute_Barrier :742-941,
//                                  Compute_Barrier_Gradient :943-1256, Compute_Barrier_Hessian
//                                  :1258-1731, Compute_Intersection_Free_StepSize :1879-2244,
//                                  Compute_Min_Dist2 :2246-2388)
//   Library/Math/Distance/CCD.h   (ACCD kernels :279-483)
// It keeps the reference's parallel structure so that it can double as the CPU baseline:
// Par_Each loops are `omp parallel for`; hash-map insertion, the PP/PE merge, the energy, the
// gradient and min-dist loops are serial exactly as in the reference.
//
// PARITY STATUS: the reference ships no tests and cannot be built here (needs Eigen, Kokkos,
// Cabana, Boost, ...).  The geometric kernels (distances, classifiers, derivatives, barrier, ACCD
// per pair) are pinned against the reference's OWN headers compiled with a stub Eigen
// (oracle/_ref, see oracle/ref_build/); the hash / constraint-set / step-size drivers are pinned
// only by self-consistency (hash == brute force, FD checks): "parity unpinned" for those.
//
// Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline, --impl reference) may use this.
#include "geom.h"
#include "derivs.h"
#include "eig.h"

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <vector>
#include <array>
#include <map>
#include <set>
#include <unordered_map>
#include <unordered_set>
#include <numeric>
#include <chrono>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace cipc_oracle {

typedef std::array<int, 4> I4;
typedef std::array<double, 2> D2;

struct Scene {
    int nV = 0;
    const double* X = nullptr;   // 3*nV
    const double* X0 = nullptr;  // 3*nV rest positions (nodeAttr.x0)
    int nBN = 0, nBE = 0, nBT = 0, nRod = 0;
    const int* BN = nullptr;     // nBN
    const int* BE = nullptr;     // 2*nBE
    const int* BT = nullptr;     // 3*nBT
    int codim0 = 0, codim1 = 0;  // codimBNStartInd
    const uint8_t* DBC = nullptr; // nV
    std::map<int, std::set<int>> NNX; // NNExclusion
    const double *BNArea = nullptr, *BEArea = nullptr, *BTArea = nullptr;
    V3 x(int v) const { return V3(X + 3 * v); }
    V3 x0(int v) const { return V3(X0 + 3 * v); }
};

// ======================================================================= SPATIAL_HASH
struct SpatialHash {
    V3 lbc, rtc;
    double one_div_voxelSize = 0;
    int vc[3] = {1, 1, 1};
    int vc01 = 1;
    int edgeStart = 0, triStart = 0;
    std::unordered_map<int, std::vector<int>> voxel;
    std::vector<std::vector<int>> occupancy; // pointAndEdgeOccupancy (CCD)

    // SPATIAL_HASH.h:702-708
    void axis_index(const V3& pos, int* out) const
    {
        out[0] = (int)std::floor((pos.x - lbc.x) * one_div_voxelSize);
        out[1] = (int)std::floor((pos.y - lbc.y) * one_div_voxelSize);
        out[2] = (int)std::floor((pos.z - lbc.z) * one_div_voxelSize);
    }
    int lin(const int* a) const { return a[0] + a[1] * vc[0] + a[2] * vc01; }

    // SPATIAL_HASH.h:71-86 / :504-519 (shared by both builds)
    void size_grid(double voxelSize, const char* tag, bool quiet)
    {
        const double range[3] = {rtc.x - lbc.x, rtc.y - lbc.y, rtc.z - lbc.z};
        one_div_voxelSize = 1.0 / voxelSize;
        long voxelAmt = 1;
        for (int d = 0; d < 3; ++d) voxelAmt *= std::max(1L, (long)std::ceil(range[d] * one_div_voxelSize));
        if (voxelAmt > 1e9) {
            voxelSize *= std::pow(voxelAmt / 1.0e9, 1.0 / 3);
            one_div_voxelSize = 1.0 / voxelSize;
        }
        for (int d = 0; d < 3; ++d) vc[d] = std::max(1, (int)std::ceil(range[d] * one_div_voxelSize));
        if (!quiet) printf("%s SH voxel count %d\n", tag, vc[0] * vc[1] * vc[2]);
        if (std::min(vc[0], std::min(vc[1], vc[2])) <= 0) {
            one_div_voxelSize = 1.0 / (std::max(range[0], std::max(range[1], range[2])) * 1.01);
            vc[0] = vc[1] = vc[2] = 1;
        }
        vc01 = vc[0] * vc[1];
    }

    static double mean_edge_len(const Scene& s)
    {
        // SPATIAL_HASH.h:60-67 / :455-464: eLen.mean().  Eigen's vectorised redux order is not
        // restated; a plain sequential sum is used (differs by O(n eps), shifts voxel borders only).
        std::vector<double> eLen(s.nBE);
#pragma omp parallel for schedule(static)
        for (int e = 0; e < s.nBE; ++e) eLen[e] = std::sqrt(norm2(s.x(s.BE[2 * e]) - s.x(s.BE[2 * e + 1])));
        double sum = 0;
        for (int e = 0; e < s.nBE; ++e) sum += eLen[e];
        return sum / s.nBE;
    }

    // SPATIAL_HASH.h:28-213
    void build_static(const Scene& s, double voxelSize, bool quiet)
    {
        if (s.nBE) voxelSize *= mean_edge_len(s);
        lbc = V3(1e300, 1e300, 1e300); rtc = V3(-1e300, -1e300, -1e300);
        for (int v = 0; v < s.nV; ++v) { lbc = vmin(lbc, s.x(v)); rtc = vmax(rtc, s.x(v)); }
        size_grid(voxelSize, "CCS", quiet);
        edgeStart = s.nBN; triStart = edgeStart + s.nBE;

        std::vector<std::array<int, 3>> svVAI(s.nBN);
        std::vector<int> vI2SVI(s.nV, 0);
#pragma omp parallel for schedule(static)
        for (int svI = 0; svI < s.nBN; ++svI) axis_index(s.x(s.BN[svI]), svVAI[svI].data());
        for (int svI = 0; svI < s.nBN; ++svI) vI2SVI[s.BN[svI]] = svI;

        voxel.clear();
        for (int svI = 0; svI < s.nBN; ++svI) voxel[lin(svVAI[svI].data())].emplace_back(svI);

        auto fill = [&](const int* mins, const int* maxs, std::vector<int>& out) {
            for (int iz = mins[2]; iz <= maxs[2]; ++iz)
                for (int iy = mins[1]; iy <= maxs[1]; ++iy)
                    for (int ix = mins[0]; ix <= maxs[0]; ++ix) out.emplace_back(ix + iy * vc[0] + iz * vc01);
        };
        std::vector<std::vector<int>> locE(s.nBE), locT(s.nBT);
#pragma omp parallel for schedule(dynamic, 256)
        for (int e = 0; e < s.nBE; ++e) {
            const auto& a = svVAI[vI2SVI[s.BE[2 * e]]];
            const auto& b = svVAI[vI2SVI[s.BE[2 * e + 1]]];
            int mins[3], maxs[3];
            for (int d = 0; d < 3; ++d) { mins[d] = std::min(a[d], b[d]); maxs[d] = std::max(a[d], b[d]); }
            fill(mins, maxs, locE[e]);
        }
#pragma omp parallel for schedule(dynamic, 256)
        for (int t = 0; t < s.nBT; ++t) {
            const auto& a = svVAI[vI2SVI[s.BT[3 * t]]];
            const auto& b = svVAI[vI2SVI[s.BT[3 * t + 1]]];
            const auto& c = svVAI[vI2SVI[s.BT[3 * t + 2]]];
            int mins[3], maxs[3];
            for (int d = 0; d < 3; ++d) {
                mins[d] = std::min(std::min(a[d], b[d]), c[d]);
                maxs[d] = std::max(std::max(a[d], b[d]), c[d]);
            }
            fill(mins, maxs, locT[t]);
        }
        for (int e = 0; e < s.nBE; ++e) for (int c : locE[e]) voxel[c].emplace_back(e + edgeStart);
        for (int t = 0; t < s.nBT; ++t) for (int c : locT[t]) voxel[c].emplace_back(t + triStart);
    }

    template <class F>
    void for_range(const V3& lo, const V3& hi, F f) const
    {
        int mins[3], maxs[3];
        axis_index(lo, mins); axis_index(hi, maxs);
        for (int d = 0; d < 3; ++d) { mins[d] = std::max(mins[d], 0); maxs[d] = std::min(maxs[d], vc[d] - 1); }
        for (int iz = mins[2]; iz <= maxs[2]; ++iz)
            for (int iy = mins[1]; iy <= maxs[1]; ++iy)
                for (int ix = mins[0]; ix <= maxs[0]; ++ix) {
                    auto it = voxel.find(ix + iy * vc[0] + iz * vc01);
                    if (it != voxel.end()) for (int ind : it->second) f(ind);
                }
    }
    // SPATIAL_HASH.h:215-241
    void query_point_for_triangles(const V3& p, double r, std::unordered_set<int>& out) const
    {
        out.clear();
        for_range(V3(p.x - r, p.y - r, p.z - r), V3(p.x + r, p.y + r, p.z + r),
            [&](int ind) { if (ind >= triStart) out.insert(ind - triStart); });
    }
    // SPATIAL_HASH.h:293-336
    void query_point_for_edges(const V3& p, double r, std::unordered_set<int>& out) const
    {
        out.clear();
        for_range(V3(p.x - r, p.y - r, p.z - r), V3(p.x + r, p.y + r, p.z + r),
            [&](int ind) { if (ind >= edgeStart && ind < triStart) out.insert(ind - edgeStart); });
    }
    // SPATIAL_HASH.h:338-381
    void query_point_for_points(const V3& p, double r, std::unordered_set<int>& out) const
    {
        out.clear();
        for_range(V3(p.x - r, p.y - r, p.z - r), V3(p.x + r, p.y + r, p.z + r),
            [&](int ind) { if (ind < edgeStart) out.insert(ind); });
    }
    // SPATIAL_HASH.h:243-291
    void query_edge_for_edges(const V3& a, const V3& b, double r, std::vector<int>& out, int eIq) const
    {
        out.resize(0);
        const V3 lo = vmin(a, b), hi = vmax(a, b);
        for_range(V3(lo.x - r, lo.y - r, lo.z - r), V3(hi.x + r, hi.y + r, hi.z + r),
            [&](int ind) { if (ind >= edgeStart && ind < triStart && ind - edgeStart > eIq) out.emplace_back(ind - edgeStart); });
        std::sort(out.begin(), out.end());
        out.erase(std::unique(out.begin(), out.end()), out.end());
    }

    // SPATIAL_HASH.h:432-622 (swept build; may shrink curMaxStepSize)
    void build_swept(const Scene& s, const double* searchDir, double& curMaxStepSize, double voxelSize,
        double thickness, bool quiet)
    {
        if (s.nBE) voxelSize *= mean_edge_len(s);
        double pSize = 0;
        for (int svI = 0; svI < s.nBN; ++svI) {
            const int vI = s.BN[svI];
            pSize += std::fabs(searchDir[vI * 3]);
            pSize += std::fabs(searchDir[vI * 3 + 1]);
            pSize += std::fabs(searchDir[vI * 3 + 2]);
        }
        pSize /= s.nBN * 3;
        const double spanSize = curMaxStepSize * pSize / voxelSize;
        if (!quiet) printf("span size = %g\n", spanSize);
        if (spanSize > 1) {
            curMaxStepSize /= spanSize;
            if (!quiet) printf("curMaxStepSize reduced\n");
        }
        std::vector<V3> SV(s.nBN), SVt(s.nBN);
        std::unordered_map<int, int> vI2SVI;
        for (int svI = 0; svI < s.nBN; ++svI) {
            const int vI = s.BN[svI];
            vI2SVI[vI] = svI;
            SV[svI] = s.x(vI);
            SVt[svI] = V3(SV[svI].x + curMaxStepSize * searchDir[vI * 3], SV[svI].y + curMaxStepSize * searchDir[vI * 3 + 1],
                SV[svI].z + curMaxStepSize * searchDir[vI * 3 + 2]);
        }
        V3 mnS(1e300, 1e300, 1e300), mxS(-1e300, -1e300, -1e300), mnT = mnS, mxT = mxS;
        for (int i = 0; i < s.nBN; ++i) {
            mnS = vmin(mnS, SV[i]); mxS = vmax(mxS, SV[i]);
            mnT = vmin(mnT, SVt[i]); mxT = vmax(mxT, SVt[i]);
        }
        const V3 mn = vmin(mnS, mnT), mx = vmax(mxS, mxT);
        lbc = V3(mn.x - thickness / 2, mn.y - thickness / 2, mn.z - thickness / 2);
        rtc = V3(mx.x + thickness / 2, mx.y + thickness / 2, mx.z + thickness / 2);
        size_grid(voxelSize, "CCD", quiet);
        edgeStart = s.nBN; triStart = edgeStart + s.nBE;

        std::vector<std::array<int, 3>> svMin(s.nBN), svMax(s.nBN);
#pragma omp parallel for schedule(static)
        for (int svI = 0; svI < s.nBN; ++svI) {
            const V3 lo = vmin(SV[svI], SVt[svI]), hi = vmax(SV[svI], SVt[svI]);
            axis_index(V3(lo.x - thickness / 2, lo.y - thickness / 2, lo.z - thickness / 2), svMin[svI].data());
            axis_index(V3(hi.x + thickness / 2, hi.y + thickness / 2, hi.z + thickness / 2), svMax[svI].data());
        }
        voxel.clear();
        occupancy.assign(triStart, std::vector<int>());
        auto fill = [&](const int* mins, const int* maxs, std::vector<int>& out) {
            for (int iz = mins[2]; iz <= maxs[2]; ++iz)
                for (int iy = mins[1]; iy <= maxs[1]; ++iy)
                    for (int ix = mins[0]; ix <= maxs[0]; ++ix) out.emplace_back(ix + iy * vc[0] + iz * vc01);
        };
#pragma omp parallel for schedule(dynamic, 256)
        for (int svI = 0; svI < s.nBN; ++svI) fill(svMin[svI].data(), svMax[svI].data(), occupancy[svI]);
#pragma omp parallel for schedule(dynamic, 256)
        for (int e = 0; e < s.nBE; ++e) {
            const int a = vI2SVI.at(s.BE[2 * e]), b = vI2SVI.at(s.BE[2 * e + 1]);
            int mins[3], maxs[3];
            for (int d = 0; d < 3; ++d) { mins[d] = std::min(svMin[a][d], svMin[b][d]); maxs[d] = std::max(svMax[a][d], svMax[b][d]); }
            fill(mins, maxs, occupancy[e + edgeStart]);
        }
        std::vector<std::vector<int>> locT(s.nBT);
#pragma omp parallel for schedule(dynamic, 256)
        for (int t = 0; t < s.nBT; ++t) {
            const int a = vI2SVI.at(s.BT[3 * t]), b = vI2SVI.at(s.BT[3 * t + 1]), c = vI2SVI.at(s.BT[3 * t + 2]);
            int mins[3], maxs[3];
            for (int d = 0; d < 3; ++d) {
                mins[d] = std::min(std::min(svMin[a][d], svMin[b][d]), svMin[c][d]);
                maxs[d] = std::max(std::max(svMax[a][d], svMax[b][d]), svMax[c][d]);
            }
            fill(mins, maxs, locT[t]);
        }
        for (int i = 0; i < (int)occupancy.size(); ++i) for (int c : occupancy[i]) voxel[c].emplace_back(i);
        for (int t = 0; t < s.nBT; ++t) for (int c : locT[t]) voxel[c].emplace_back(t + triStart);
    }
    // SPATIAL_HASH.h:624-646
    void query_point_for_primitives(int svI, std::unordered_set<int>& pts, std::unordered_set<int>& edges,
        std::unordered_set<int>& tris) const
    {
        pts.clear(); edges.clear(); tris.clear();
        for (int c : occupancy[svI]) {
            auto it = voxel.find(c);
            for (int ind : it->second) {
                if (ind >= triStart) tris.insert(ind - triStart);
                else if (ind >= edgeStart) edges.insert(ind - edgeStart);
                else pts.insert(ind);
            }
        }
    }
    // SPATIAL_HASH.h:648-661
    void query_edge_for_edges(int seI, std::unordered_set<int>& edges) const
    {
        edges.clear();
        for (int c : occupancy[seI + edgeStart]) {
            auto it = voxel.find(c);
            for (int ind : it->second)
                if (ind >= edgeStart && ind < triStart && ind - edgeStart > seI) edges.insert(ind - edgeStart);
        }
    }
};

// ======================================================================= pair filters
static inline bool nnx_hit(const Scene& s, int v, int a, int b, int c)
{
    auto f = s.NNX.find(v);
    if (f == s.NNX.end()) return false;
    return f->second.count(a) || f->second.count(b) || (c >= 0 && f->second.count(c));
}
// IPC.h:171-181
static inline bool pt_pair_ok(const Scene& s, int vI, const int* t)
{
    if (vI == t[0] || vI == t[1] || vI == t[2]) return false;
    if (s.DBC[vI] && s.DBC[t[0]] && s.DBC[t[1]] && s.DBC[t[2]]) return false;
    if (nnx_hit(s, vI, t[0], t[1], t[2])) return false;
    return true;
}
// IPC.h:384-397
static inline bool ee_pair_ok(const Scene& s, int eI, int eJ, const int* a, const int* b)
{
    if (a[0] == b[0] || a[0] == b[1] || a[1] == b[0] || a[1] == b[1] || eI > eJ) return false;
    if (s.DBC[a[0]] && s.DBC[a[1]] && s.DBC[b[0]] && s.DBC[b[1]]) return false;
    if (nnx_hit(s, a[0], b[0], b[1], -1) || nnx_hit(s, a[1], b[0], b[1], -1)) return false;
    return true;
}

// ======================================================================= Compute_Constraint_Set
struct Timers { double t[8] = {0}; };
static inline double now_s()
{
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// IPC.h:143-661 (3-D branch).  use_hash=false takes the reference's `#else` brute-force loops.
static void compute_constraint_set(const Scene& s, bool elastic, double dHat2, double thickness, bool use_hash,
    std::vector<I4>& constraintSet, std::vector<D2>& stencilInfo, Timers* tm, bool quiet)
{
    SpatialHash sh;
    double t0 = now_s();
    if (use_hash) sh.build_static(s, 1.0, quiet);
    double t1 = now_s();
    if (elastic) thickness = 0;
    const double dHat = std::sqrt(dHat2) + thickness;
    dHat2 = dHat * dHat;

    // ---- point-triangle pass (IPC.h:145-355)
    std::vector<std::vector<I4>> csPT(s.nBN);
    std::vector<std::vector<D2>> infoPT(s.nBN);
#pragma omp parallel for schedule(dynamic, 64)
    for (int svI = 0; svI < s.nBN; ++svI) {
        const int vI = s.BN[svI];
        const V3 p = s.x(vI);
        std::unordered_set<int> triInds;
        auto visit_tri = [&](int sfI) {
            const int* t = s.BT + 3 * sfI;
            if (!pt_pair_ok(s, vI, t)) return;
            const V3 t0 = s.x(t[0]), t1 = s.x(t[1]), t2 = s.x(t[2]);
            if (!pt_cd_broadphase(p, t0, t1, t2, dHat)) return;
            double d = 1e300;
            switch (pt_type(p, t0, t1, t2)) {
            case 0: d = pp_dist2(p, t0); if (d < dHat2) csPT[svI].push_back({-vI - 1, t[0], -1, -1}); break;
            case 1: d = pp_dist2(p, t1); if (d < dHat2) csPT[svI].push_back({-vI - 1, t[1], -1, -1}); break;
            case 2: d = pp_dist2(p, t2); if (d < dHat2) csPT[svI].push_back({-vI - 1, t[2], -1, -1}); break;
            case 3: d = pe_dist2(p, t0, t1); if (d < dHat2) csPT[svI].push_back({-vI - 1, t[0], t[1], -1}); break;
            case 4: d = pe_dist2(p, t1, t2); if (d < dHat2) csPT[svI].push_back({-vI - 1, t[1], t[2], -1}); break;
            case 5: d = pe_dist2(p, t2, t0); if (d < dHat2) csPT[svI].push_back({-vI - 1, t[2], t[0], -1}); break;
            case 6: d = pt_dist2(p, t0, t1, t2); if (d < dHat2) csPT[svI].push_back({-vI - 1, t[0], t[1], t[2]}); break;
            default: break;
            }
            if (d < dHat2) {
                double w = elastic ? s.BNArea[svI] * s.BTArea[sfI] : 1.0; // BNArea is misaligned for codim nodes; not read when !elastic
                if (elastic && svI < s.codim0) w /= 2;
                infoPT[svI].push_back({w, dHat2});
            }
        };
        if (use_hash) {
            sh.query_point_for_triangles(p, dHat, triInds);
            for (int sfI : triInds) visit_tri(sfI);
        }
        else for (int sfI = 0; sfI < s.nBT; ++sfI) visit_tri(sfI);

        if (svI >= s.codim0) { // rod or particle points vs rod edges (IPC.h:271-326)
            auto visit_edge = [&](int eI) {
                if (eI < s.nBE - s.nRod) return;
                const int* e = s.BE + 2 * eI;
                if (vI == e[0] || vI == e[1] || (s.DBC[vI] && s.DBC[e[0]] && s.DBC[e[1]])) return;
                const V3 e0 = s.x(e[0]), e1 = s.x(e[1]);
                if (!pe_cd_broadphase(p, e0, e1, dHat)) return;
                double d = 1e300, ratio;
                switch (pe_type(p, e0, e1, ratio)) {
                case 0: d = pp_dist2(p, e0); if (d < dHat2) csPT[svI].push_back({-vI - 1, e[0], -1, -1}); break;
                case 1: d = pp_dist2(p, e1); if (d < dHat2) csPT[svI].push_back({-vI - 1, e[1], -1, -1}); break;
                case 2: d = pe_dist2(p, e0, e1); if (d < dHat2) csPT[svI].push_back({-vI - 1, e[0], e[1], -1}); break;
                }
                if (d < dHat2) infoPT[svI].push_back({1.0, dHat2});
            };
            if (use_hash) {
                std::unordered_set<int> edgeInds;
                sh.query_point_for_edges(p, dHat, edgeInds);
                for (int eI : edgeInds) visit_edge(eI);
            }
            else for (int eI = 0; eI < s.nBE; ++eI) visit_edge(eI);

            if (svI >= s.codim1) { // particle vs later points (IPC.h:328-352)
                auto visit_point = [&](int svJ) {
                    const int vJ = s.BN[svJ];
                    if (svJ > svI && !(s.DBC[vI] && s.DBC[vJ])) {
                        const double d = pp_dist2(p, s.x(vJ));
                        if (d < dHat2) {
                            csPT[svI].push_back({-vI - 1, vJ, -1, -1});
                            infoPT[svI].push_back({1.0, dHat2});
                        }
                    }
                };
                if (use_hash) {
                    std::unordered_set<int> pointInds;
                    sh.query_point_for_points(p, dHat, pointInds);
                    for (int svJ : pointInds) visit_point(svJ);
                }
                else for (int svJ = 0; svJ < s.nBN; ++svJ) visit_point(svJ);
            }
        }
    }
    double t2 = now_s();

    // ---- edge-edge pass (IPC.h:358-568)
    std::vector<std::vector<I4>> csEE(s.nBE);
    std::vector<std::vector<D2>> infoEE(s.nBE);
#pragma omp parallel for schedule(dynamic, 64)
    for (int eI = 0; eI < s.nBE; ++eI) {
        const int* a = s.BE + 2 * eI;
        const V3 ea0 = s.x(a[0]), ea1 = s.x(a[1]);
        auto visit = [&](int eJ) {
            const int* b = s.BE + 2 * eJ;
            if (!ee_pair_ok(s, eI, eJ, a, b)) return;
            const V3 eb0 = s.x(b[0]), eb1 = s.x(b[1]);
            if (!ee_cd_broadphase(ea0, ea1, eb0, eb1, dHat)) return;
            const double cn2 = ee_cross_norm2(ea0, ea1, eb0, eb1);
            const double eps_x = ee_mollifier_threshold(s.x0(a[0]), s.x0(a[1]), s.x0(b[0]), s.x0(b[1]));
            const bool mol = (cn2 < eps_x);
            double d = 1e300;
            auto& out = csEE[eI];
            switch (ee_type(ea0, ea1, eb0, eb1)) {
            case 0: d = pp_dist2(ea0, eb0); if (d < dHat2) out.push_back(mol ? I4{a[0], b[0], -a[1] - 1, -b[1] - 1} : I4{-a[0] - 1, b[0], -1, -1}); break;
            case 1: d = pp_dist2(ea0, eb1); if (d < dHat2) out.push_back(mol ? I4{a[0], b[1], -a[1] - 1, -b[0] - 1} : I4{-a[0] - 1, b[1], -1, -1}); break;
            case 2: d = pe_dist2(ea0, eb0, eb1); if (d < dHat2) out.push_back(mol ? I4{a[0], b[0], b[1], -a[1] - 1} : I4{-a[0] - 1, b[0], b[1], -1}); break;
            case 3: d = pp_dist2(ea1, eb0); if (d < dHat2) out.push_back(mol ? I4{a[1], b[0], -a[0] - 1, -b[1] - 1} : I4{-a[1] - 1, b[0], -1, -1}); break;
            case 4: d = pp_dist2(ea1, eb1); if (d < dHat2) out.push_back(mol ? I4{a[1], b[1], -a[0] - 1, -b[0] - 1} : I4{-a[1] - 1, b[1], -1, -1}); break;
            case 5: d = pe_dist2(ea1, eb0, eb1); if (d < dHat2) out.push_back(mol ? I4{a[1], b[0], b[1], -a[0] - 1} : I4{-a[1] - 1, b[0], b[1], -1}); break;
            case 6: d = pe_dist2(eb0, ea0, ea1); if (d < dHat2) out.push_back(mol ? I4{b[0], a[0], a[1], -b[1] - 1} : I4{-b[0] - 1, a[0], a[1], -1}); break;
            case 7: d = pe_dist2(eb1, ea0, ea1); if (d < dHat2) out.push_back(mol ? I4{b[1], a[0], a[1], -b[0] - 1} : I4{-b[1] - 1, a[0], a[1], -1}); break;
            case 8: d = ee_dist2(ea0, ea1, eb0, eb1); if (d < dHat2) out.push_back(mol ? I4{a[0], a[1], -b[0] - 1, b[1]} : I4{a[0], a[1], b[0], b[1]}); break;
            default: break;
            }
            if (d < dHat2) {
                double w = elastic ? s.BEArea[eI] * s.BEArea[eJ] : 1.0;
                if (elastic && (eI >= s.nBE - s.nRod) && (eJ >= s.nBE - s.nRod)) w *= 2;
                infoEE[eI].push_back({w, dHat2});
            }
        };
        if (use_hash) {
            std::vector<int> edgeInds;
            sh.query_edge_for_edges(ea0, ea1, dHat, edgeInds, eI);
            for (int eJ : edgeInds) visit(eJ);
        }
        else for (int eJ = eI + 1; eJ < s.nBE; ++eJ) visit(eJ);
    }
    double t3 = now_s();

    // ---- merge (IPC.h:571-661): PT/EE/mollified pass through; PP/PE de-duplicated by raw key
    constraintSet.resize(0);
    stencilInfo.resize(0);
    std::map<I4, int> counter;
    std::map<I4, D2> areaCounter;
    for (int i = 0; i < s.nBN; ++i)
        for (size_t k = 0; k < csPT[i].size(); ++k) {
            const I4& c = csPT[i][k];
            if (c[3] < 0) {
                ++counter[c];
                auto f = areaCounter.find(c);
                if (f == areaCounter.end()) areaCounter[c] = infoPT[i][k];
                else f->second[0] += infoPT[i][k][0];
            }
            else { constraintSet.push_back(c); stencilInfo.push_back(infoPT[i][k]); }
        }
    for (int i = 0; i < s.nBE; ++i)
        for (size_t k = 0; k < csEE[i].size(); ++k) {
            const I4& c = csEE[i][k];
            if (c[0] < 0) {
                ++counter[c];
                auto f = areaCounter.find(c);
                if (f == areaCounter.end()) areaCounter[c] = infoEE[i][k];
                else f->second[0] += infoEE[i][k][0];
            }
            else { constraintSet.push_back(c); stencilInfo.push_back(infoEE[i][k]); }
        }
    for (const auto& cc : counter) {
        constraintSet.push_back({cc.first[0], cc.first[1], cc.first[2], -cc.second});
        const D2& a = areaCounter[cc.first];
        stencilInfo.push_back({a[0] / cc.second, a[1]});
    }
    if (!elastic) for (auto& i : stencilInfo) i[0] = 1;
    double t4 = now_s();
    if (tm) { tm->t[0] = t1 - t0; tm->t[1] = t2 - t1; tm->t[2] = t3 - t2; tm->t[3] = t4 - t3; }
}

// ======================================================================= stencil decoding
// Appendix-A dispatch shared by energy/gradient/Hessian/min-dist (IPC.h:801-938 etc.)
enum Kind { K_PT, K_PE, K_PP, K_EE, K_EE_M, K_PE_M, K_PP_M };
struct Stencil {
    Kind kind;
    int v[4];    // vertex ids in block order
    int mult;    // multiplicity m (1 unless de-duplicated PP/PE)
};
static inline Stencil decode(const I4& c)
{
    Stencil st;
    st.mult = 1;
    if (c[0] >= 0) {
        if (c[3] >= 0 && c[2] >= 0) { st.kind = K_EE; st.v[0] = c[0]; st.v[1] = c[1]; st.v[2] = c[2]; st.v[3] = c[3]; }
        else if (c[3] >= 0) { st.kind = K_EE_M; st.v[0] = c[0]; st.v[1] = c[1]; st.v[2] = -c[2] - 1; st.v[3] = c[3]; }
        else if (c[2] >= 0) { st.kind = K_PE_M; st.v[0] = c[0]; st.v[1] = -c[3] - 1; st.v[2] = c[1]; st.v[3] = c[2]; }
        else { st.kind = K_PP_M; st.v[0] = c[0]; st.v[1] = -c[2] - 1; st.v[2] = c[1]; st.v[3] = -c[3] - 1; }
    }
    else {
        st.v[0] = -c[0] - 1; st.v[1] = c[1]; st.v[2] = c[2]; st.v[3] = c[3];
        if (c[3] >= 0) st.kind = K_PT;
        else if (c[2] >= 0) { st.kind = K_PE; st.mult = -c[3]; }
        else { st.kind = K_PP; st.mult = -c[3]; }
    }
    return st;
}
static inline double stencil_dist2(const Scene& s, const Stencil& st)
{
    switch (st.kind) {
    case K_PT: return pt_dist2(s.x(st.v[0]), s.x(st.v[1]), s.x(st.v[2]), s.x(st.v[3]));
    case K_PE: return pe_dist2(s.x(st.v[0]), s.x(st.v[1]), s.x(st.v[2]));
    case K_PP: return pp_dist2(s.x(st.v[0]), s.x(st.v[1]));
    case K_EE: case K_EE_M: return ee_dist2(s.x(st.v[0]), s.x(st.v[1]), s.x(st.v[2]), s.x(st.v[3]));
    case K_PE_M: return pe_dist2(s.x(st.v[0]), s.x(st.v[2]), s.x(st.v[3]));
    default: return pp_dist2(s.x(st.v[0]), s.x(st.v[2]));
    }
}
static inline bool mollified(Kind k) { return k == K_EE_M || k == K_PE_M || k == K_PP_M; }

// ======================================================================= Compute_Barrier (IPC.h:742-941)
// returns 0, or 1 if a non-positive distance was met (reference: printf + exit(-1))
static int compute_barrier(const Scene& s, bool elastic, const std::vector<I4>& cs, const std::vector<D2>& info,
    double dHat2, const double* kappa, double thickness, double& E)
{
    if (elastic) thickness = 0;
    const double thickness2 = thickness * thickness;
    dHat2 += 2 * std::sqrt(dHat2) * thickness;
    std::vector<double> barrierV(cs.size());
    for (size_t cI = 0; cI < cs.size(); ++cI) {
        const Stencil st = decode(cs[cI]);
        double dist2 = stencil_dist2(s, st);
        dist2 -= thickness2;
        if (dist2 <= 0) return 1;
        double b = barrier(elastic, dist2, dHat2, kappa);
        if (mollified(st.kind)) {
            const double eps_x = ee_mollifier_threshold(s.x0(st.v[0]), s.x0(st.v[1]), s.x0(st.v[2]), s.x0(st.v[3]));
            b *= ee_mollifier(s.x(st.v[0]), s.x(st.v[1]), s.x(st.v[2]), s.x(st.v[3]), eps_x);
        }
        else if (cs[cI][0] < 0 && cs[cI][3] < -1) b *= -cs[cI][3];
        b *= info[cI][0];
        barrierV[cI] = b;
    }
    E += std::accumulate(barrierV.begin(), barrierV.end(), 0.0);
    return 0;
}

// ======================================================================= Compute_Barrier_Gradient (IPC.h:943-1256)
// g: 3*nV, accumulated in place (nodeAttr.g +=)
static void compute_barrier_gradient(const Scene& s, bool elastic, const std::vector<I4>& cs, const std::vector<D2>& info,
    double dHat2, const double* kappa, double thickness, double* g)
{
    if (elastic) thickness = 0;
    const double thickness2 = thickness * thickness;
    dHat2 += 2 * std::sqrt(dHat2) * thickness;
    for (size_t cI = 0; cI < cs.size(); ++cI) {
        const Stencil st = decode(cs[cI]);
        const double w = info[cI][0];
        const double dist2 = stencil_dist2(s, st) - thickness2;
        const double bG = barrier_gradient(elastic, dist2, dHat2, kappa);
        double dg[12];
        auto add = [&](int slot, const double* v3, double sc) {
            double* o = g + 3 * st.v[slot];
            o[0] += sc * v3[0]; o[1] += sc * v3[1]; o[2] += sc * v3[2];
        };
        if (!mollified(st.kind)) {
            int nb = 0;
            switch (st.kind) {
            case K_PT: pt_grad(s.x(st.v[0]), s.x(st.v[1]), s.x(st.v[2]), s.x(st.v[3]), dg); nb = 4; break;
            case K_EE: ee_grad(s.x(st.v[0]), s.x(st.v[1]), s.x(st.v[2]), s.x(st.v[3]), dg); nb = 4; break;
            case K_PE: pe_grad(s.x(st.v[0]), s.x(st.v[1]), s.x(st.v[2]), dg); nb = 3; break;
            default: pp_grad(s.x(st.v[0]), s.x(st.v[1]), dg); nb = 2; break;
            }
            const double sc = st.mult * w * bG; // IPC.h:1202,1227,1249
            for (int k = 0; k < nb; ++k) add(k, dg + 3 * k, sc);
        }
        else {
            const V3 a0 = s.x(st.v[0]), a1 = s.x(st.v[1]), b0 = s.x(st.v[2]), b1 = s.x(st.v[3]);
            const double b = barrier(elastic, dist2, dHat2, kappa);
            const double eps_x = ee_mollifier_threshold(s.x0(st.v[0]), s.x0(st.v[1]), s.x0(st.v[2]), s.x0(st.v[3]));
            const double e = ee_mollifier(a0, a1, b0, b1, eps_x);
            double eg[12];
            ee_mollifier_grad(a0, a1, b0, b1, eps_x, eg);
            for (int k = 0; k < 4; ++k) add(k, eg + 3 * k, w * b);
            if (st.kind == K_EE_M) { // IPC.h:1081
                ee_grad(a0, a1, b0, b1, dg);
                for (int k = 0; k < 4; ++k) add(k, dg + 3 * k, w * e * bG);
            }
            else if (st.kind == K_PE_M) { // IPC.h:1122-1130: rows {0,2,3}
                pe_grad(a0, b0, b1, dg);
                add(0, dg, e * w * bG); add(2, dg + 3, e * w * bG); add(3, dg + 6, e * w * bG);
            }
            else { // IPC.h:1167-1174: rows {0,2}
                pp_grad(a0, b0, dg);
                add(0, dg, e * w * bG); add(2, dg + 3, e * w * bG);
            }
        }
    }
}

// ======================================================================= Compute_Barrier_Hessian (IPC.h:1258-1731)
struct Triplet { int row, col; double val; };

static inline int block_dim(const I4& c) { return (c[0] >= 0 || c[3] >= 0) ? 12 : (c[2] >= 0 ? 9 : 6); }

static void stencil_hessian(const Scene& s, bool elastic, const I4& c, double w, double dHat2, const double* kappa,
    double thickness2, bool projectSPD, double* H /*n*n*/, int* vids, int& nb)
{
    const Stencil st = decode(c);
    const double dist2 = stencil_dist2(s, st) - thickness2;
    const double bG = barrier_gradient(elastic, dist2, dHat2, kappa);
    const double bH = barrier_hessian(elastic, dist2, dHat2, kappa);
    double dg[12], dH[144];
    if (!mollified(st.kind)) {
        switch (st.kind) {
        case K_PT: nb = 4; pt_grad(s.x(st.v[0]), s.x(st.v[1]), s.x(st.v[2]), s.x(st.v[3]), dg);
            pt_hess(s.x(st.v[0]), s.x(st.v[1]), s.x(st.v[2]), s.x(st.v[3]), dH); break;
        case K_EE: nb = 4; ee_grad(s.x(st.v[0]), s.x(st.v[1]), s.x(st.v[2]), s.x(st.v[3]), dg);
            ee_hess(s.x(st.v[0]), s.x(st.v[1]), s.x(st.v[2]), s.x(st.v[3]), dH); break;
        case K_PE: nb = 3; pe_grad(s.x(st.v[0]), s.x(st.v[1]), s.x(st.v[2]), dg);
            pe_hess(s.x(st.v[0]), s.x(st.v[1]), s.x(st.v[2]), dH); break;
        default: nb = 2; pp_grad(s.x(st.v[0]), s.x(st.v[1]), dg); pp_hess(s.x(st.v[0]), s.x(st.v[1]), dH); break;
        }
        const int n = 3 * nb;
        const double m = st.mult; // IPC.h:1640-1642: ((m bH) g) g^T + (m bG) H
        for (int i = 0; i < n; ++i)
            for (int j = 0; j < n; ++j) H[i * n + j] = (((m * bH) * dg[i]) * dg[j] + (m * bG) * dH[i * n + j]) * w;
        for (int k = 0; k < nb; ++k) vids[k] = st.v[k];
        if (projectSPD) make_pd(n, H);
        return;
    }
    nb = 4;
    for (int k = 0; k < 4; ++k) vids[k] = st.v[k];
    const V3 a0 = s.x(st.v[0]), a1 = s.x(st.v[1]), b0 = s.x(st.v[2]), b1 = s.x(st.v[3]);
    const double b = barrier(elastic, dist2, dHat2, kappa);
    const double eps_x = ee_mollifier_threshold(s.x0(st.v[0]), s.x0(st.v[1]), s.x0(st.v[2]), s.x0(st.v[3]));
    const double e = ee_mollifier(a0, a1, b0, b1, eps_x);
    double eg[12], eH[144];
    ee_mollifier_grad(a0, a1, b0, b1, eps_x, eg);
    ee_mollifier_hess(a0, a1, b0, b1, eps_x, eH);
    // embed the distance gradient / Hessian into the 12-dof (ea0,ea1,eb0,eb1) frame
    double G[12] = {0}, K[144] = {0};
    int rows[4], nrow;
    if (st.kind == K_EE_M) {
        ee_grad(a0, a1, b0, b1, dg); ee_hess(a0, a1, b0, b1, dH);
        nrow = 4; rows[0] = 0; rows[1] = 1; rows[2] = 2; rows[3] = 3;
    }
    else if (st.kind == K_PE_M) { // IPC.h:1526-1537
        pe_grad(a0, b0, b1, dg); pe_hess(a0, b0, b1, dH);
        nrow = 3; rows[0] = 0; rows[1] = 2; rows[2] = 3;
    }
    else { // IPC.h:1588-1599
        pp_grad(a0, b0, dg); pp_hess(a0, b0, dH);
        nrow = 2; rows[0] = 0; rows[1] = 2;
    }
    const int nn = 3 * nrow;
    for (int I = 0; I < nrow; ++I)
        for (int r = 0; r < 3; ++r) {
            G[3 * rows[I] + r] = dg[3 * I + r];
            for (int J = 0; J < nrow; ++J)
                for (int cc = 0; cc < 3; ++cc) K[(3 * rows[I] + r) * 12 + 3 * rows[J] + cc] = dH[(3 * I + r) * nn + 3 * J + cc];
        }
    // IPC.h:1472-1475: bG (G eg^T + eg G^T) + b eH + e bH G G^T + e bG K, times w
    for (int i = 0; i < 12; ++i)
        for (int j = 0; j < 12; ++j)
            H[i * 12 + j] = (bG * (G[i] * eg[j] + eg[i] * G[j]) + b * eH[i * 12 + j] + (e * bH * G[i]) * G[j] + e * bG * K[i * 12 + j]) * w;
    if (projectSPD) make_pd(12, H);
}

static void compute_barrier_hessian(const Scene& s, bool elastic, const std::vector<I4>& cs, const std::vector<D2>& info,
    double dHat2, const double* kappa, double thickness, bool projectSPD, std::vector<Triplet>& triplets)
{
    if (elastic) thickness = 0;
    const double thickness2 = thickness * thickness;
    dHat2 += 2 * std::sqrt(dHat2) * thickness;
    std::vector<size_t> start(cs.size());
    size_t cur = triplets.size();
    for (size_t cI = 0; cI < cs.size(); ++cI) { // IPC.h:1370-1388
        start[cI] = cur;
        const int n = block_dim(cs[cI]);
        cur += (size_t)n * n;
    }
    triplets.resize(cur);
#pragma omp parallel for schedule(dynamic, 64)
    for (long cI = 0; cI < (long)cs.size(); ++cI) {
        double H[144];
        int vids[4], nb;
        stencil_hessian(s, elastic, cs[cI], info[cI][0], dHat2, kappa, thickness2, projectSPD, H, vids, nb);
        const int n = 3 * nb;
        Triplet* out = triplets.data() + start[cI];
        for (int i = 0; i < nb; ++i)
            for (int j = 0; j < nb; ++j)
                for (int a = 0; a < 3; ++a)
                    for (int b = 0; b < 3; ++b)
                        out[(i * 3 + a) * n + j * 3 + b] = Triplet{vids[i] * 3 + a, vids[j] * 3 + b, H[(i * 3 + a) * n + j * 3 + b]};
    }
}

// ======================================================================= ACCD (Math/Distance/CCD.h:279-483)
// CCD.h:279-328
static bool pt_accd(V3 p, V3 t0, V3 t1, V3 t2, V3 dp, V3 dt0, V3 dt1, V3 dt2, double eta, double thickness, double& toc)
{
    const V3 mov = (((dt0 + dt1) + dt2) + dp) / 4;
    dt0 = dt0 - mov; dt1 = dt1 - mov; dt2 = dt2 - mov; dp = dp - mov;
    const double maxDispMag = std::sqrt(norm2(dp)) + std::sqrt(std::max(std::max(norm2(dt0), norm2(dt1)), norm2(dt2)));
    if (maxDispMag == 0) return false;
    double dist2_cur = pt_dist2_unclassified(p, t0, t1, t2);
    double dist_cur = std::sqrt(dist2_cur);
    const double gap = eta * (dist2_cur - thickness * thickness) / (dist_cur + thickness);
    const double toc_prev = toc;
    toc = 0;
    while (true) {
        const double tocLowerBound = (1 - eta) * (dist2_cur - thickness * thickness) / ((dist_cur + thickness) * maxDispMag);
        p = p + tocLowerBound * dp; t0 = t0 + tocLowerBound * dt0; t1 = t1 + tocLowerBound * dt1; t2 = t2 + tocLowerBound * dt2;
        dist2_cur = pt_dist2_unclassified(p, t0, t1, t2);
        dist_cur = std::sqrt(dist2_cur);
        if (toc && ((dist2_cur - thickness * thickness) / (dist_cur + thickness) < gap)) break;
        toc += tocLowerBound;
        if (toc > toc_prev) return false;
    }
    return true;
}
// CCD.h:330-395
static bool ee_accd(V3 ea0, V3 ea1, V3 eb0, V3 eb1, V3 dea0, V3 dea1, V3 deb0, V3 deb1, double eta, double thickness, double& toc)
{
    const V3 mov = (((dea0 + dea1) + deb0) + deb1) / 4;
    dea0 = dea0 - mov; dea1 = dea1 - mov; deb0 = deb0 - mov; deb1 = deb1 - mov;
    const double maxDispMag = std::sqrt(std::max(norm2(dea0), norm2(dea1))) + std::sqrt(std::max(norm2(deb0), norm2(deb1)));
    if (maxDispMag == 0) return false;
    auto min_endpoint = [&]() {
        return std::min(std::min(norm2(ea0 - eb0), norm2(ea0 - eb1)), std::min(norm2(ea1 - eb0), norm2(ea1 - eb1)));
    };
    double dist2_cur = ee_dist2_unclassified(ea0, ea1, eb0, eb1);
    double dFunc = dist2_cur - thickness * thickness;
    if (dFunc <= 0) { dist2_cur = min_endpoint(); dFunc = dist2_cur - thickness * thickness; }
    double dist_cur = std::sqrt(dist2_cur);
    const double gap = eta * dFunc / (dist_cur + thickness);
    const double toc_prev = toc;
    toc = 0;
    while (true) {
        const double tocLowerBound = (1 - eta) * dFunc / ((dist_cur + thickness) * maxDispMag);
        ea0 = ea0 + tocLowerBound * dea0; ea1 = ea1 + tocLowerBound * dea1; eb0 = eb0 + tocLowerBound * deb0; eb1 = eb1 + tocLowerBound * deb1;
        dist2_cur = ee_dist2_unclassified(ea0, ea1, eb0, eb1);
        dFunc = dist2_cur - thickness * thickness;
        if (dFunc <= 0) { dist2_cur = min_endpoint(); dFunc = dist2_cur - thickness * thickness; }
        dist_cur = std::sqrt(dist2_cur);
        if (toc && (dFunc / (dist_cur + thickness) < gap)) break;
        toc += tocLowerBound;
        if (toc > toc_prev) return false;
    }
    return true;
}
// CCD.h:397-441
static bool pe_accd(V3 p, V3 e0, V3 e1, V3 dp, V3 de0, V3 de1, double eta, double thickness, double& toc)
{
    const V3 mov = ((dp + de0) + de1) / 3;
    de0 = de0 - mov; de1 = de1 - mov; dp = dp - mov;
    const double maxDispMag = std::sqrt(norm2(dp)) + std::sqrt(std::max(norm2(de0), norm2(de1)));
    if (maxDispMag == 0) return false;
    double dist2_cur = pe_dist2_unclassified(p, e0, e1);
    double dist_cur = std::sqrt(dist2_cur);
    const double gap = eta * (dist2_cur - thickness * thickness) / (dist_cur + thickness);
    const double toc_prev = toc;
    toc = 0;
    while (true) {
        const double tocLowerBound = (1 - eta) * (dist2_cur - thickness * thickness) / ((dist_cur + thickness) * maxDispMag);
        p = p + tocLowerBound * dp; e0 = e0 + tocLowerBound * de0; e1 = e1 + tocLowerBound * de1;
        dist2_cur = pe_dist2_unclassified(p, e0, e1);
        dist_cur = std::sqrt(dist2_cur);
        if (toc && (dist2_cur - thickness * thickness) / (dist_cur + thickness) < gap) break;
        toc += tocLowerBound;
        if (toc > toc_prev) return false;
    }
    return true;
}
// CCD.h:443-483
static bool pp_accd(V3 p0, V3 p1, V3 dp0, V3 dp1, double eta, double thickness, double& toc)
{
    const V3 mov = (dp0 + dp1) / 2;
    dp1 = dp1 - mov; dp0 = dp0 - mov;
    const double maxDispMag = std::sqrt(norm2(dp0)) + std::sqrt(norm2(dp1));
    if (maxDispMag == 0) return false;
    double dist2_cur = pp_dist2(p0, p1);
    double dist_cur = std::sqrt(dist2_cur);
    const double gap = eta * (dist2_cur - thickness * thickness) / (dist_cur + thickness);
    const double toc_prev = toc;
    toc = 0;
    while (true) {
        const double tocLowerBound = (1 - eta) * (dist2_cur - thickness * thickness) / ((dist_cur + thickness) * maxDispMag);
        p0 = p0 + tocLowerBound * dp0; p1 = p1 + tocLowerBound * dp1;
        dist2_cur = pp_dist2(p0, p1);
        dist_cur = std::sqrt(dist2_cur);
        if (toc && (dist2_cur - thickness * thickness) / (dist_cur + thickness) < gap) break;
        toc += tocLowerBound;
        if (toc > toc_prev) return false;
    }
    return true;
}

// ======================================================================= Compute_Intersection_Free_StepSize (IPC.h:1879-2244)
// returns 0 ok, 2 if a PT pair produced largestAlpha == 0 (reference dumps coordinates and exit(-1))
static int compute_step_size(const Scene& s, bool elastic, const double* searchDir, double thickness, bool use_hash,
    double& stepSize, Timers* tm, bool quiet, long* nPairs)
{
    if (elastic) thickness = 0;
    SpatialHash sh;
    double t0 = now_s();
    if (use_hash) sh.build_swept(s, searchDir, stepSize, 1.0, thickness, quiet);
    double t1 = now_s();
    int err = 0;
    long pairs = 0;
    auto dir = [&](int v) { return V3(searchDir + 3 * v); };

    std::vector<double> alphaPT(s.nBN);
#pragma omp parallel for schedule(dynamic, 64) reduction(+ : pairs)
    for (int svI = 0; svI < s.nBN; ++svI) {
        const int vI = s.BN[svI];
        const V3 p = s.x(vI), dp = dir(vI);
        alphaPT[svI] = stepSize;
        std::unordered_set<int> sP, sE, sT;
        if (use_hash) sh.query_point_for_primitives(svI, sP, sE, sT);
        auto visit_tri = [&](int sfI) {
            const int* t = s.BT + 3 * sfI;
            if (!pt_pair_ok(s, vI, t)) return;
            const V3 t0 = s.x(t[0]), t1 = s.x(t[1]), t2 = s.x(t[2]);
            const V3 dt0 = dir(t[0]), dt1 = dir(t[1]), dt2 = dir(t[2]);
            if (!pt_ccd_broadphase(p, t0, t1, t2, dp, dt0, dt1, dt2, thickness)) return;
            ++pairs;
            double largestAlpha = alphaPT[svI];
            if (pt_accd(p, t0, t1, t2, dp, dt0, dt1, dt2, 0.1, thickness, largestAlpha))
                if (alphaPT[svI] > largestAlpha) alphaPT[svI] = largestAlpha;
            if (largestAlpha == 0) {
#pragma omp atomic write
                err = 2;
            }
        };
        if (use_hash) for (int sfI : sT) visit_tri(sfI);
        else for (int sfI = 0; sfI < s.nBT; ++sfI) visit_tri(sfI);

        if (svI >= s.codim1) { // particle vs rod edges and later points (IPC.h:2098-2163)
            auto visit_edge = [&](int eI) {
                if (eI < s.nBE - s.nRod) return;
                const int* e = s.BE + 2 * eI;
                if (vI == e[0] || vI == e[1] || (s.DBC[vI] && s.DBC[e[0]] && s.DBC[e[1]])) return;
                const V3 e0 = s.x(e[0]), e1 = s.x(e[1]), de0 = dir(e[0]), de1 = dir(e[1]);
                if (!pe_ccd_broadphase(p, e0, e1, dp, de0, de1, thickness)) return;
                ++pairs;
                double largestAlpha = alphaPT[svI];
                if (pe_accd(p, e0, e1, dp, de0, de1, 0.1, thickness, largestAlpha))
                    if (alphaPT[svI] > largestAlpha) alphaPT[svI] = largestAlpha;
            };
            if (use_hash) for (int eI : sE) visit_edge(eI);
            else for (int eI = 0; eI < s.nBE; ++eI) visit_edge(eI);
            auto visit_point = [&](int svJ) {
                if (svJ <= svI) return;
                const int vJ = s.BN[svJ];
                if (s.DBC[vI] && s.DBC[vJ]) return;
                const V3 pJ = s.x(vJ), dpJ = dir(vJ);
                if (!pp_ccd_broadphase(p, pJ, dp, dpJ, thickness)) return;
                ++pairs;
                double largestAlpha = alphaPT[svI];
                if (pp_accd(p, pJ, dp, dpJ, 0.1, thickness, largestAlpha))
                    if (alphaPT[svI] > largestAlpha) alphaPT[svI] = largestAlpha;
            };
            if (use_hash) for (int svJ : sP) visit_point(svJ);
            else for (int svJ = 0; svJ < s.nBN; ++svJ) visit_point(svJ);
        }
    }
    if (s.nBN) stepSize = std::min(stepSize, *std::min_element(alphaPT.begin(), alphaPT.end()));
    double t2 = now_s();

    std::vector<double> alphaEE(s.nBE);
#pragma omp parallel for schedule(dynamic, 64) reduction(+ : pairs)
    for (int eI = 0; eI < s.nBE; ++eI) {
        const int* a = s.BE + 2 * eI;
        const V3 ea0 = s.x(a[0]), ea1 = s.x(a[1]), dea0 = dir(a[0]), dea1 = dir(a[1]);
        alphaEE[eI] = stepSize;
        auto visit = [&](int eJ) {
            const int* b = s.BE + 2 * eJ;
            if (!ee_pair_ok(s, eI, eJ, a, b)) return;
            const V3 eb0 = s.x(b[0]), eb1 = s.x(b[1]), deb0 = dir(b[0]), deb1 = dir(b[1]);
            if (!ee_ccd_broadphase(ea0, ea1, eb0, eb1, dea0, dea1, deb0, deb1, thickness)) return;
            ++pairs;
            double largestAlpha = alphaEE[eI];
            if (ee_accd(ea0, ea1, eb0, eb1, dea0, dea1, deb0, deb1, 0.1, thickness, largestAlpha))
                if (alphaEE[eI] > largestAlpha) alphaEE[eI] = largestAlpha;
        };
        if (use_hash) {
            std::unordered_set<int> sE;
            sh.query_edge_for_edges(eI, sE);
            for (int eJ : sE) visit(eJ);
        }
        else for (int eJ = eI + 1; eJ < s.nBE; ++eJ) visit(eJ);
    }
    if (s.nBE) stepSize = std::min(stepSize, *std::min_element(alphaEE.begin(), alphaEE.end()));
    double t3 = now_s();
    if (tm) { tm->t[0] = t1 - t0; tm->t[1] = t2 - t1; tm->t[2] = t3 - t2; }
    if (nPairs) *nPairs = pairs;
    return err;
}

// ======================================================================= Compute_Min_Dist2 (IPC.h:2246-2388)
static void compute_min_dist2(const Scene& s, const std::vector<I4>& cs, double thickness, std::vector<double>& dist2, double& minDist2)
{
    dist2.resize(cs.size());
    for (size_t cI = 0; cI < cs.size(); ++cI) dist2[cI] = stencil_dist2(s, decode(cs[cI]));
    minDist2 = *std::min_element(dist2.begin(), dist2.end());
    minDist2 -= thickness * thickness;
}

// ======================================================================= friction (FEM/FRICTION.h, FEM/FRICTION_UTILS.h)
// SURVEY 8(f)-1.  Lagged friction on the non-mollified part of the contact constraint set.
struct FrictionSet {
    std::vector<I4> cs;
    std::vector<std::array<double, 2>> closest;   // closestPoint (beta / gamma / yita)
    std::vector<std::array<double, 6>> basis;     // tanBasis, column-major 3x2: col0 xyz, col1 xyz
    std::vector<double> normalForce;
};
static inline V3 normalized(const V3& a) // Eigen::normalized(): v / sqrt(squaredNorm) when the norm is positive
{
    const double z = norm2(a);
    return z > 0 ? a / std::sqrt(z) : a;
}
// FRICTION_UTILS.h:10-39
static inline double f0_SF(double x2, double epsvh)
{
    if (x2 >= epsvh * epsvh) return std::sqrt(x2);
    return x2 * (-std::sqrt(x2) / 3.0 + epsvh) / (epsvh * epsvh) + epsvh / 3.0;
}
static inline double f1_SF_div(double x2, double epsvh)
{
    if (x2 >= epsvh * epsvh) return 1 / std::sqrt(x2);
    return (-std::sqrt(x2) + 2.0 * epsvh) / (epsvh * epsvh);
}
static inline double f2_SF_term(double, double epsvh) { return -1 / (epsvh * epsvh); }
static inline void set_basis(std::array<double, 6>& B, const V3& c0, const V3& c1)
{
    B = {c0.x, c0.y, c0.z, c1.x, c1.y, c1.z};
}
// FRICTION_UTILS.h:41-52, 107-118, 184-194, 229-245
static inline void pt_tangent_basis(const V3&, const V3& v1, const V3& v2, const V3& v3, std::array<double, 6>& B)
{
    const V3 v12 = v2 - v1;
    set_basis(B, normalized(v12), normalized(cross(cross(v12, v3 - v1), v12)));
}
static inline void ee_tangent_basis(const V3& v0, const V3& v1, const V3& v2, const V3& v3, std::array<double, 6>& B)
{
    const V3 v01 = v1 - v0;
    set_basis(B, normalized(v01), normalized(cross(cross(v01, v3 - v2), v01)));
}
static inline void pe_tangent_basis(const V3& v0, const V3& v1, const V3& v2, std::array<double, 6>& B)
{
    const V3 v12 = v2 - v1;
    set_basis(B, normalized(v12), normalized(cross(v12, v0 - v1)));
}
static inline void pp_tangent_basis(const V3& v0, const V3& v1, std::array<double, 6>& B)
{
    const V3 v01 = v1 - v0;
    const V3 xC = cross(V3(1, 0, 0), v01), yC = cross(V3(0, 1, 0), v01);
    if (norm2(xC) > norm2(yC)) set_basis(B, normalized(xC), normalized(cross(v01, xC)));
    else set_basis(B, normalized(yC), normalized(cross(v01, yC)));
}
// FRICTION_UTILS.h:54-66, 120-142, 196-204
static inline void pt_closest_point(const V3& v0, const V3& v1, const V3& v2, const V3& v3, double& b1, double& b2)
{
    const V3 r0 = v2 - v1, r1 = v3 - v1, rel = v0 - v1;
    ldlt2_solve(dot(r0, r0), dot(r1, r0), dot(r1, r1), dot(r0, rel), dot(r1, rel), b1, b2);
}
static inline void ee_closest_point(const V3& v0, const V3& v1, const V3& v2, const V3& v3, double& g1, double& g2)
{
    const V3 e20 = v0 - v2, e01 = v1 - v0, e23 = v3 - v2;
    ldlt2_solve(norm2(e01), -dot(e23, e01), norm2(e23), -dot(e20, e01), dot(e20, e23), g1, g2);
}
static inline double pe_closest_point(const V3& v0, const V3& v1, const V3& v2)
{
    const V3 e12 = v2 - v1;
    return dot(v0 - v1, e12) / norm2(e12);
}
// vertex coefficients of the relative displacement u = sum_k coef_k dx_k (FRICTION_UTILS.h: *_RelDX / *_TT)
static inline int friction_coefs(const I4& c, const std::array<double, 2>& cp, int* v, double* coef, int& mult)
{
    mult = 1;
    if (c[0] >= 0) { // EE
        v[0] = c[0]; v[1] = c[1]; v[2] = c[2]; v[3] = c[3];
        coef[0] = 1.0 - cp[0]; coef[1] = cp[0]; coef[2] = cp[1] - 1.0; coef[3] = -cp[1];
        return 4;
    }
    v[0] = -c[0] - 1; v[1] = c[1]; v[2] = c[2]; v[3] = c[3];
    if (c[2] < 0) { coef[0] = 1.0; coef[1] = -1.0; mult = -c[3]; return 2; }                                   // PP
    if (c[3] < 0) { coef[0] = 1.0; coef[1] = cp[0] - 1.0; coef[2] = -cp[0]; mult = -c[3]; return 3; }          // PE
    coef[0] = 1.0; coef[1] = -1.0 + cp[0] + cp[1]; coef[2] = -cp[0]; coef[3] = -cp[1];                         // PT
    return 4;
}
// FRICTION.h:16-124
static void compute_friction_basis(const Scene& s, bool elastic, const std::vector<I4>& contactCs, const std::vector<D2>& info,
    double dHat2, const double* kappa, double thickness, FrictionSet& F)
{
    if (elastic) thickness = 0;
    const double thickness2 = thickness * thickness;
    dHat2 += 2 * std::sqrt(dHat2) * thickness;
    F.cs.clear();
    for (const auto& c : contactCs)
        if (!(c[0] >= 0 && (c[2] < 0 || c[3] < 0))) F.cs.push_back(c);
    const size_t n = F.cs.size();
    F.closest.assign(n, {0.0, 0.0});
    F.basis.resize(n);
    F.normalForce.resize(n);
    for (size_t cI = 0; cI < n; ++cI) {
        const I4& c = F.cs[cI];
        double dist2;
        if (c[0] >= 0) {
            const V3 a0 = s.x(c[0]), a1 = s.x(c[1]), b0 = s.x(c[2]), b1 = s.x(c[3]);
            ee_closest_point(a0, a1, b0, b1, F.closest[cI][0], F.closest[cI][1]);
            ee_tangent_basis(a0, a1, b0, b1, F.basis[cI]);
            dist2 = ee_dist2(a0, a1, b0, b1);
        }
        else {
            const V3 p = s.x(-c[0] - 1);
            if (c[2] < 0) { pp_tangent_basis(p, s.x(c[1]), F.basis[cI]); dist2 = pp_dist2(p, s.x(c[1])); }
            else if (c[3] < 0) {
                F.closest[cI][0] = pe_closest_point(p, s.x(c[1]), s.x(c[2]));
                pe_tangent_basis(p, s.x(c[1]), s.x(c[2]), F.basis[cI]);
                dist2 = pe_dist2(p, s.x(c[1]), s.x(c[2]));
            }
            else {
                pt_closest_point(p, s.x(c[1]), s.x(c[2]), s.x(c[3]), F.closest[cI][0], F.closest[cI][1]);
                pt_tangent_basis(p, s.x(c[1]), s.x(c[2]), s.x(c[3]), F.basis[cI]);
                dist2 = pt_dist2(p, s.x(c[1]), s.x(c[2]), s.x(c[3]));
            }
        }
        const double bGrad = barrier_gradient(elastic, dist2 - thickness2, dHat2, kappa);
        // the reference indexes stencilInfo with the FILTERED index (FRICTION.h:112); weights are 1 when !elasticIPC
        F.normalForce[cI] = -bGrad * 2 * std::sqrt(dist2) * info[cI][0];
    }
}
// FRICTION.h:126-170
static void compute_friction_coef(const std::vector<I4>& cs, const std::vector<int>& compNodeRange, const std::vector<double>& muComp,
    std::vector<double>& normalForce, double& mu)
{
    mu = 1;
    auto comp = [&](int v) {
        for (size_t k = 0; k < compNodeRange.size(); ++k) if (v < compNodeRange[k]) return (int)k;
        return -1;
    };
#pragma omp parallel for schedule(static)
    for (long cI = 0; cI < (long)cs.size(); ++cI) {
        const I4& c = cs[cI];
        const int c0 = comp(c[0] >= 0 ? c[0] : -c[0] - 1), c1 = comp(c[0] >= 0 ? c[2] : c[1]);
        normalForce[cI] *= muComp[c0 + c1 * compNodeRange.size()];
    }
}
static inline void friction_rel(const Scene& s, const double* Xn, const FrictionSet& F, size_t cI, int* v, double* coef, int& nb, int& mult,
    double* u2, V3& rel3)
{
    nb = friction_coefs(F.cs[cI], F.closest[cI], v, coef, mult);
    rel3 = V3(0, 0, 0);
    for (int k = 0; k < nb; ++k) rel3 = rel3 + coef[k] * (s.x(v[k]) - V3(Xn + 3 * v[k]));
    const auto& B = F.basis[cI];
    u2[0] = dot(V3(B[0], B[1], B[2]), rel3);
    u2[1] = dot(V3(B[3], B[4], B[5]), rel3);
}
// FRICTION.h:172-252
static void compute_friction_potential(const Scene& s, const double* Xn, const FrictionSet& F, double epsvh2, double mu, double& E)
{
    const double epsvh = std::sqrt(epsvh2);
    std::vector<double> EI(F.cs.size());
    for (size_t cI = 0; cI < F.cs.size(); ++cI) {
        int v[4], nb, mult; double coef[4], u[2]; V3 r;
        friction_rel(s, Xn, F, cI, v, coef, nb, mult, u, r);
        EI[cI] = f0_SF(u[0] * u[0] + u[1] * u[1], epsvh) * F.normalForce[cI];
        if (F.cs[cI][3] < -1) EI[cI] *= -F.cs[cI][3];
    }
    E += mu * std::accumulate(EI.begin(), EI.end(), 0.0);
}
// FRICTION.h:254-379
static void compute_friction_gradient(const Scene& s, const double* Xn, const FrictionSet& F, double epsvh2, double mu, double* g)
{
    const double epsvh = std::sqrt(epsvh2);
    for (size_t cI = 0; cI < F.cs.size(); ++cI) {
        int v[4], nb, mult; double coef[4], u[2]; V3 r;
        friction_rel(s, Xn, F, cI, v, coef, nb, mult, u, r);
        const double sc = f1_SF_div(u[0] * u[0] + u[1] * u[1], epsvh) * mult * mu * F.normalForce[cI];
        const auto& B = F.basis[cI];
        const V3 t3(sc * (B[0] * u[0] + B[3] * u[1]), sc * (B[1] * u[0] + B[4] * u[1]), sc * (B[2] * u[0] + B[5] * u[1]));
        for (int k = 0; k < nb; ++k) { g[3 * v[k]] += coef[k] * t3.x; g[3 * v[k] + 1] += coef[k] * t3.y; g[3 * v[k] + 2] += coef[k] * t3.z; }
    }
}
// FRICTION.h:381-663
static void compute_friction_hessian(const Scene& s, const double* Xn, const FrictionSet& F, double epsvh2, double mu, bool projectSPD,
    std::vector<Triplet>& triplets)
{
    const double epsvh = std::sqrt(epsvh2);
    std::vector<size_t> start(F.cs.size());
    size_t cur = triplets.size();
    for (size_t cI = 0; cI < F.cs.size(); ++cI) { start[cI] = cur; const int n = block_dim(F.cs[cI]); cur += (size_t)n * n; }
    triplets.resize(cur);
#pragma omp parallel for schedule(dynamic, 256)
    for (long cI = 0; cI < (long)F.cs.size(); ++cI) {
        int v[4], nb, mult; double coef[4], u[2]; V3 r;
        friction_rel(s, Xn, F, cI, v, coef, nb, mult, u, r);
        const double x2 = u[0] * u[0] + u[1] * u[1], xn = std::sqrt(x2);
        const double f1 = f1_SF_div(x2, epsvh), f2 = f2_SF_term(x2, epsvh);
        const double sc = mult * mu * F.normalForce[cI];
        double inner[4];
        if (x2 >= epsvh2) {
            const double ub[2] = {-u[1], u[0]}, k = sc * f1 / x2;
            inner[0] = k * ub[0] * ub[0]; inner[1] = inner[2] = k * ub[0] * ub[1]; inner[3] = k * ub[1] * ub[1];
        }
        else if (xn == 0) { inner[0] = inner[3] = sc * f1; inner[1] = inner[2] = 0; }
        else {
            inner[0] = (f2 / xn) * u[0] * u[0] + f1; inner[1] = inner[2] = (f2 / xn) * u[0] * u[1]; inner[3] = (f2 / xn) * u[1] * u[1] + f1;
            if (projectSPD) make_pd(2, inner);
            for (int k = 0; k < 4; ++k) inner[k] *= sc;
        }
        // H = TT^T inner TT, TT = coef (x) basis^T
        const auto& B = F.basis[cI];
        double M[9]; // basis inner basis^T (3x3)
        for (int a = 0; a < 3; ++a)
            for (int b = 0; b < 3; ++b)
                M[3 * a + b] = B[a] * (inner[0] * B[b] + inner[1] * B[3 + b]) + B[3 + a] * (inner[2] * B[b] + inner[3] * B[3 + b]);
        const int n = 3 * nb;
        Triplet* out = triplets.data() + start[cI];
        for (int i = 0; i < nb; ++i)
            for (int j = 0; j < nb; ++j)
                for (int a = 0; a < 3; ++a)
                    for (int b = 0; b < 3; ++b)
                        out[(i * 3 + a) * n + j * 3 + b] = Triplet{v[i] * 3 + a, v[j] * 3 + b, coef[i] * coef[j] * M[3 * a + b]};
    }
}

} // namespace cipc_oracle

// ======================================================================= C API (ctypes)
using namespace cipc_oracle;

namespace {
struct SceneHolder { Scene s; };
std::vector<I4> g_cs;
std::vector<D2> g_info;
std::vector<Triplet> g_trip;
Timers g_tm;

void set_cs(const int* cs, const double* info, int n, std::vector<I4>& c, std::vector<D2>& w)
{
    c.resize(n); w.resize(n);
    for (int i = 0; i < n; ++i) {
        c[i] = {cs[4 * i], cs[4 * i + 1], cs[4 * i + 2], cs[4 * i + 3]};
        w[i] = {info[2 * i], info[2 * i + 1]};
    }
}
} // namespace

extern "C" {

void* oracle_scene_create(int nV, const double* X, const double* X0, int nBN, const int* BN, int nBE, const int* BE,
    int nBT, const int* BT, int nRod, int codim0, int codim1, const uint8_t* DBC, int nnxPairs, const int* nnxPairList,
    const double* BNArea, const double* BEArea, const double* BTArea)
{
    SceneHolder* h = new SceneHolder();
    Scene& s = h->s;
    s.nV = nV; s.X = X; s.X0 = X0; s.nBN = nBN; s.BN = BN; s.nBE = nBE; s.BE = BE; s.nBT = nBT; s.BT = BT;
    s.nRod = nRod; s.codim0 = codim0; s.codim1 = codim1; s.DBC = DBC;
    for (int i = 0; i < nnxPairs; ++i) s.NNX[nnxPairList[2 * i]].insert(nnxPairList[2 * i + 1]);
    s.BNArea = BNArea; s.BEArea = BEArea; s.BTArea = BTArea;
    return h;
}
void oracle_scene_set_X(void* h, const double* X) { ((SceneHolder*)h)->s.X = X; }
void oracle_scene_destroy(void* h) { delete (SceneHolder*)h; }
int oracle_num_threads()
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
void oracle_set_num_threads(int n)
{
#ifdef _OPENMP
    omp_set_num_threads(n);
#else
    (void)n;
#endif
}

// returns number of constraints; fetch with oracle_fetch_constraints
int oracle_constraint_set(void* h, int elastic, double dHat2, double thickness, int use_hash, double* timers4)
{
    compute_constraint_set(((SceneHolder*)h)->s, elastic != 0, dHat2, thickness, use_hash != 0, g_cs, g_info, &g_tm, true);
    if (timers4) for (int i = 0; i < 4; ++i) timers4[i] = g_tm.t[i];
    return (int)g_cs.size();
}
void oracle_fetch_constraints(int* cs, double* info)
{
    for (size_t i = 0; i < g_cs.size(); ++i) {
        for (int k = 0; k < 4; ++k) cs[4 * i + k] = g_cs[i][k];
        info[2 * i] = g_info[i][0]; info[2 * i + 1] = g_info[i][1];
    }
}
int oracle_barrier(void* h, int elastic, const int* cs, const double* info, int n, double dHat2, const double* kappa,
    double thickness, double* E)
{
    std::vector<I4> c; std::vector<D2> w;
    set_cs(cs, info, n, c, w);
    return compute_barrier(((SceneHolder*)h)->s, elastic != 0, c, w, dHat2, kappa, thickness, *E);
}
void oracle_barrier_gradient(void* h, int elastic, const int* cs, const double* info, int n, double dHat2,
    const double* kappa, double thickness, double* g)
{
    std::vector<I4> c; std::vector<D2> w;
    set_cs(cs, info, n, c, w);
    compute_barrier_gradient(((SceneHolder*)h)->s, elastic != 0, c, w, dHat2, kappa, thickness, g);
}
// returns number of triplets appended; fetch with oracle_fetch_triplets
long oracle_barrier_hessian(void* h, int elastic, const int* cs, const double* info, int n, double dHat2,
    const double* kappa, double thickness, int projectSPD)
{
    std::vector<I4> c; std::vector<D2> w;
    set_cs(cs, info, n, c, w);
    g_trip.clear();
    compute_barrier_hessian(((SceneHolder*)h)->s, elastic != 0, c, w, dHat2, kappa, thickness, projectSPD != 0, g_trip);
    return (long)g_trip.size();
}
void oracle_fetch_triplets(int* rows, int* cols, double* vals)
{
    for (size_t i = 0; i < g_trip.size(); ++i) { rows[i] = g_trip[i].row; cols[i] = g_trip[i].col; vals[i] = g_trip[i].val; }
}
// ---- checker helpers for full-size parity (bench.py's parity leg, tests/test_gpu_cfg5_full.py): the triplet streams of two
// implementations (16-byte {int row; int col; double val} records, blocks in constraint order) are compared block by block
// without leaving native memory.  out[0] = max_c ||Ha - Hb||_F / ||Hb||_F, out[1] = number of (row, col) mismatches,
// out[2] = max_c |v^T (Ha - Hb) v| / ||Hb||_F with v = cos(1 + 0.37 k) (the golden fixtures' quadratic probe),
// out[3] = number of blocks, out[4] = total number of triplets implied by cs
const void* oracle_triplets_data(long* n) { if (n) *n = (long)g_trip.size(); return g_trip.data(); }
void oracle_compare_triplet_blocks(const void* a_, const void* b_, const int* cs, int nC, double* out)
{
    const Triplet* A = (const Triplet*)a_;
    const Triplet* B = (const Triplet*)b_;
    std::vector<long> off((size_t)nC + 1, 0);
    for (int i = 0; i < nC; ++i) {
        const int* c = cs + 4 * (size_t)i;
        const int n = (c[0] >= 0 || c[3] >= 0) ? 12 : (c[2] >= 0 ? 9 : 6);
        off[i + 1] = off[i] + n * n;
    }
    double worst = 0, worstQ = 0;
    long mism = 0;
#pragma omp parallel for schedule(static) reduction(max : worst, worstQ) reduction(+ : mism)
    for (int i = 0; i < nC; ++i) {
        const long o = off[i], m = off[i + 1] - off[i];
        const int n = m == 144 ? 12 : (m == 81 ? 9 : 6);
        double d2 = 0, b2 = 0, q = 0;
        for (long t = 0; t < m; ++t) {
            const Triplet &x = A[o + t], &y = B[o + t];
            if (x.row != y.row || x.col != y.col) ++mism;
            const double d = x.val - y.val;
            d2 += d * d; b2 += y.val * y.val;
            q += std::cos(1.0 + 0.37 * (double)(t / n)) * d * std::cos(1.0 + 0.37 * (double)(t % n));
        }
        if (b2 > 0) { worst = std::max(worst, std::sqrt(d2 / b2)); worstQ = std::max(worstQ, std::fabs(q) / std::sqrt(b2)); }
        else if (d2 > 0) worst = std::max(worst, 1.0);
    }
    out[0] = worst; out[1] = (double)mism; out[2] = worstQ; out[3] = (double)nC; out[4] = (double)off[nC];
}
// ---- boundary-primitive construction, restated literally (checker of cipc_build_boundary):
// Find_Surface_Primitives_And_Compute_Area (Utils/MESHIO.h:768-834) followed by the seg / rod / particle appends of
// Shell/IMPLICIT_EULER.h:245-277.  Same containers as the reference: std::map keyed by the oriented vertex pair (lexicographic
// operator<, Math/VECTOR.h:124-134), a dense per-vertex area accumulator, std::map<int, T> for the rod nodes.
namespace {
std::vector<int> g_bBN, g_bBE, g_bBT;
std::vector<double> g_bBNArea, g_bBEArea, g_bBTArea;
int g_bCodim[2];
}
void oracle_build_boundary(int nV, const double* X, int nTri, const int* tri, int nSeg, const int* seg, int nRod, const int* rod,
    const double* rodRadius, int nParticle, const int* particle, int* counts6)
{
    g_bBN.clear(); g_bBE.clear(); g_bBT.clear(); g_bBNArea.clear(); g_bBEArea.clear(); g_bBTArea.clear();
    typedef std::pair<int, int> E2;
    std::map<E2, double> boundaryEdgeSet;
    std::vector<double> isBoundaryNode(nV, 0.0);
    for (int t = 0; t < nTri; ++t) { // MESHIO.h:781-812
        const int* v = tri + 3 * (size_t)t;
        const V3 v0(X + 3 * (size_t)v[0]), v1(X + 3 * (size_t)v[1]), v2(X + 3 * (size_t)v[2]);
        g_bBTArea.push_back(0.5 * std::sqrt(norm2(cross(v1 - v0, v2 - v0))));
        g_bBT.push_back(v[0]); g_bBT.push_back(v[1]); g_bBT.push_back(v[2]);
        const int ea[3] = {v[0], v[1], v[2]}, eb[3] = {v[1], v[2], v[0]};
        for (int k = 0; k < 3; ++k) {
            auto finder = boundaryEdgeSet.find(E2(eb[k], ea[k]));
            if (finder == boundaryEdgeSet.end()) boundaryEdgeSet[E2(ea[k], eb[k])] = g_bBTArea.back() / 3;
            else finder->second += g_bBTArea.back() / 3;
        }
        for (int k = 0; k < 3; ++k) isBoundaryNode[v[k]] += g_bBTArea.back() / 3;
        g_bBTArea.back() /= 2;
    }
    for (const auto& i : boundaryEdgeSet) { // :816-819
        g_bBE.push_back(i.first.first); g_bBE.push_back(i.first.second);
        g_bBEArea.push_back(i.second / 2);
    }
    for (int vI = 0; vI < nV; ++vI) // :821-826
        if (isBoundaryNode[vI]) { g_bBN.push_back(vI); g_bBNArea.push_back(isBoundaryNode[vI]); }
    // Shell/IMPLICIT_EULER.h:245-277
    for (int i = 0; i < nSeg; ++i) { g_bBE.push_back(seg[2 * i]); g_bBE.push_back(seg[2 * i + 1]); }
    for (int i = 0; i < nSeg; ++i) { g_bBN.push_back(seg[2 * i]); g_bBN.push_back(seg[2 * i + 1]); }
    std::map<int, double> rodNodeArea;
    for (int i = 0; i < nRod; ++i) {
        g_bBE.push_back(rod[2 * i]); g_bBE.push_back(rod[2 * i + 1]);
        const V3 v0(X + 3 * (size_t)rod[2 * i]), v1(X + 3 * (size_t)rod[2 * i + 1]);
        g_bBEArea.push_back(std::sqrt(norm2(v0 - v1)) * M_PI * rodRadius[i] / 6);
        rodNodeArea[rod[2 * i]] += g_bBEArea.back() / 2;
        rodNodeArea[rod[2 * i + 1]] += g_bBEArea.back() / 2;
        g_bBEArea.back() /= 2;
    }
    g_bCodim[0] = (int)g_bBN.size();
    for (const auto& n : rodNodeArea) { g_bBN.push_back(n.first); g_bBNArea.push_back(n.second); }
    g_bCodim[1] = (int)g_bBN.size();
    for (int i = 0; i < nParticle; ++i) g_bBN.push_back(particle[i]);
    counts6[0] = (int)g_bBN.size(); counts6[1] = (int)(g_bBE.size() / 2); counts6[2] = (int)(g_bBT.size() / 3);
    counts6[3] = g_bCodim[0]; counts6[4] = g_bCodim[1]; counts6[5] = (int)g_bBNArea.size();
}
void oracle_fetch_boundary(int* BN, int* BE, int* BT, double* BNArea, double* BEArea, double* BTArea)
{
    std::copy(g_bBN.begin(), g_bBN.end(), BN); std::copy(g_bBE.begin(), g_bBE.end(), BE); std::copy(g_bBT.begin(), g_bBT.end(), BT);
    std::copy(g_bBNArea.begin(), g_bBNArea.end(), BNArea); std::copy(g_bBEArea.begin(), g_bBEArea.end(), BEArea);
    std::copy(g_bBTArea.begin(), g_bBTArea.end(), BTArea);
}
int oracle_step_size(void* h, int elastic, const double* searchDir, double thickness, int use_hash, double* stepSize,
    double* timers3, long* nPairs)
{
    int err = compute_step_size(((SceneHolder*)h)->s, elastic != 0, searchDir, thickness, use_hash != 0, *stepSize, &g_tm, true, nPairs);
    if (timers3) for (int i = 0; i < 3; ++i) timers3[i] = g_tm.t[i];
    return err;
}
void oracle_min_dist2(void* h, const int* cs, int n, double thickness, double* dist2, double* minDist2)
{
    std::vector<I4> c(n);
    for (int i = 0; i < n; ++i) c[i] = {cs[4 * i], cs[4 * i + 1], cs[4 * i + 2], cs[4 * i + 3]};
    std::vector<double> d;
    compute_min_dist2(((SceneHolder*)h)->s, c, thickness, d, *minDist2);
    std::memcpy(dist2, d.data(), sizeof(double) * n);
}

// ---- per-stencil probes used by the unit tests (kind: 0 PP, 1 PE, 2 PT, 3 EE, 4 EE cross-norm^2)
void oracle_dist_derivs(int kind, const double* x, double* d, double* g, double* H)
{
    const V3 a(x), b(x + 3), c(x + 6), e(x + 9);
    switch (kind) {
    case 0: *d = pp_dist2(a, b); pp_grad(a, b, g); pp_hess(a, b, H); break;
    case 1: *d = pe_dist2(a, b, c); pe_grad(a, b, c, g); pe_hess(a, b, c, H); break;
    case 2: *d = pt_dist2(a, b, c, e); pt_grad(a, b, c, e, g); pt_hess(a, b, c, e, H); break;
    case 3: *d = ee_dist2(a, b, c, e); ee_grad(a, b, c, e, g); ee_hess(a, b, c, e, H); break;
    default: *d = ee_cross_norm2(a, b, c, e); eecn2_grad(a, b, c, e, g); eecn2_hess(a, b, c, e, H); break;
    }
}
void oracle_mollifier(const double* x, double eps_x, double* e, double* g, double* H)
{
    const V3 a(x), b(x + 3), c(x + 6), d(x + 9);
    *e = ee_mollifier(a, b, c, d, eps_x);
    ee_mollifier_grad(a, b, c, d, eps_x, g);
    ee_mollifier_hess(a, b, c, d, eps_x, H);
}
int oracle_pt_type(const double* x) { return pt_type(V3(x), V3(x + 3), V3(x + 6), V3(x + 9)); }
int oracle_ee_type(const double* x) { return ee_type(V3(x), V3(x + 3), V3(x + 6), V3(x + 9)); }
int oracle_pe_type(const double* x) { double r; return pe_type(V3(x), V3(x + 3), V3(x + 6), r); }
double oracle_dist2_unclassified(int kind, const double* x)
{
    if (kind == 1) return pe_dist2_unclassified(V3(x), V3(x + 3), V3(x + 6));
    if (kind == 2) return pt_dist2_unclassified(V3(x), V3(x + 3), V3(x + 6), V3(x + 9));
    return ee_dist2_unclassified(V3(x), V3(x + 3), V3(x + 6), V3(x + 9));
}
// kind: 0 PP, 1 PE, 2 PT, 3 EE.  x = positions (4 x 3), dx = directions (4 x 3)
int oracle_accd(int kind, const double* x, const double* dx, double eta, double thickness, double* toc)
{
    const V3 a(x), b(x + 3), c(x + 6), d(x + 9), da(dx), db(dx + 3), dc(dx + 6), dd(dx + 9);
    switch (kind) {
    case 0: return pp_accd(a, b, da, db, eta, thickness, *toc);
    case 1: return pe_accd(a, b, c, da, db, dc, eta, thickness, *toc);
    case 2: return pt_accd(a, b, c, d, da, db, dc, dd, eta, thickness, *toc);
    default: return ee_accd(a, b, c, d, da, db, dc, dd, eta, thickness, *toc);
    }
}
void oracle_barrier_fn(int elastic, double d, double dHat, const double* kappa, double* out3)
{
    out3[0] = barrier(elastic != 0, d, dHat, kappa);
    out3[1] = barrier_gradient(elastic != 0, d, dHat, kappa);
    out3[2] = barrier_hessian(elastic != 0, d, dHat, kappa);
}
// ---- friction
static FrictionSet g_F;
int oracle_friction_basis(void* h, int elastic, const int* cs, const double* info, int n, double dHat2, const double* kappa, double thickness)
{
    std::vector<I4> c; std::vector<D2> w;
    set_cs(cs, info, n, c, w);
    compute_friction_basis(((SceneHolder*)h)->s, elastic != 0, c, w, dHat2, kappa, thickness, g_F);
    return (int)g_F.cs.size();
}
void oracle_fetch_friction(int* cs, double* closest, double* basis, double* nf)
{
    for (size_t i = 0; i < g_F.cs.size(); ++i) {
        for (int k = 0; k < 4; ++k) cs[4 * i + k] = g_F.cs[i][k];
        closest[2 * i] = g_F.closest[i][0]; closest[2 * i + 1] = g_F.closest[i][1];
        for (int k = 0; k < 6; ++k) basis[6 * i + k] = g_F.basis[i][k];
        nf[i] = g_F.normalForce[i];
    }
}
void oracle_set_friction(const int* cs, const double* closest, const double* basis, const double* nf, int n)
{
    g_F.cs.resize(n); g_F.closest.resize(n); g_F.basis.resize(n); g_F.normalForce.resize(n);
    for (int i = 0; i < n; ++i) {
        g_F.cs[i] = {cs[4 * i], cs[4 * i + 1], cs[4 * i + 2], cs[4 * i + 3]};
        g_F.closest[i] = {closest[2 * i], closest[2 * i + 1]};
        for (int k = 0; k < 6; ++k) g_F.basis[i][k] = basis[6 * i + k];
        g_F.normalForce[i] = nf[i];
    }
}
double oracle_friction_coef(int nComp, const int* compNodeRange, const double* muComp)
{
    std::vector<int> r(compNodeRange, compNodeRange + nComp);
    std::vector<double> m(muComp, muComp + (size_t)nComp * nComp);
    double mu;
    compute_friction_coef(g_F.cs, r, m, g_F.normalForce, mu);
    return mu;
}
void oracle_friction_potential(void* h, const double* Xn, double epsvh2, double mu, double* E)
{
    compute_friction_potential(((SceneHolder*)h)->s, Xn, g_F, epsvh2, mu, *E);
}
void oracle_friction_gradient(void* h, const double* Xn, double epsvh2, double mu, double* g)
{
    compute_friction_gradient(((SceneHolder*)h)->s, Xn, g_F, epsvh2, mu, g);
}
long oracle_friction_hessian(void* h, const double* Xn, double epsvh2, double mu, int projectSPD)
{
    g_trip.clear();
    compute_friction_hessian(((SceneHolder*)h)->s, Xn, g_F, epsvh2, mu, projectSPD != 0, g_trip);
    return (long)g_trip.size();
}
// probes of FRICTION_UTILS.h: kind 0 PP, 1 PE, 2 PT, 3 EE -> basis (6, column-major) and closest point (2)
void oracle_friction_utils(int kind, const double* x, double* basis, double* closest)
{
    const V3 a(x), b(x + 3), c(x + 6), d(x + 9);
    std::array<double, 6> B;
    closest[0] = closest[1] = 0;
    switch (kind) {
    case 0: pp_tangent_basis(a, b, B); break;
    case 1: pe_tangent_basis(a, b, c, B); closest[0] = pe_closest_point(a, b, c); break;
    case 2: pt_tangent_basis(a, b, c, d, B); pt_closest_point(a, b, c, d, closest[0], closest[1]); break;
    default: ee_tangent_basis(a, b, c, d, B); ee_closest_point(a, b, c, d, closest[0], closest[1]); break;
    }
    for (int k = 0; k < 6; ++k) basis[k] = B[k];
}
void oracle_friction_f(double x2, double epsvh, double* out3) { out3[0] = f0_SF(x2, epsvh); out3[1] = f1_SF_div(x2, epsvh); out3[2] = f2_SF_term(x2, epsvh); }
void oracle_make_pd(int n, double* H) { make_pd(n, H); }
void oracle_sym_eig(int n, const double* A, double* V, double* d) { sym_eig(n, A, V, d); }

} // extern "C"
