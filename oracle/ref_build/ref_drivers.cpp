// TEST INFRASTRUCTURE: builds oracle/_ref/libcipc_refdrv.so -- the reference's OWN contact-path
// drivers (Library/FEM/IPC.h: Compute_Constraint_Set, Compute_Barrier, _Gradient, _Hessian,
// Compute_Intersection_Free_StepSize, Compute_Min_Dist2; Library/Grid/SPATIAL_HASH.h;
// Library/FEM/FRICTION.h; Library/Math/UTILS.h; Library/Storage/*.hpp; Library/Math/VECTOR.h;
// Library/FEM/DATA_TYPE.h), included from where they lie under /root/reference and compiled against
// the stand-ins in ref_build/stub (Eigen, Cabana/Kokkos, pybind11; Utils/MESHIO.h shadowed).
// No reference source is copied: this file only #includes it and exports a C API with the same
// shape as the oracle's (oracle_* -> ref_*), so tests can run both on identical inputs.
//
// What this pins: every line of the six drivers + hash + friction as written by the reference's
// authors, with their data structures (unordered_map voxels, std::map merge, AoSoA storage) and
// their parallel structure (Par_Each -> OpenMP).  What it cannot pin: Eigen internals (stub).
#include <cstdio>
#include <cstring>
#include <chrono>
#include <string>
#include <map>
#include <vector>
#include <cmath>
#include <omp.h>
using std::log;

// TIMER_FLAG of Utils/PROFILER.h (needs Boost through Utils/PARAMETER.h): scope seconds by name
namespace cipc_ref_timer {
static std::map<std::string, double> g_scopes;
struct Scope {
    std::string name;
    std::chrono::steady_clock::time_point t0;
    explicit Scope(const char* n) : name(n), t0(std::chrono::steady_clock::now()) {}
    ~Scope() { g_scopes[name] += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); }
};
} // namespace cipc_ref_timer
#define TIMER_FLAG(name) cipc_ref_timer::Scope scoped_timer(name)
#define STORAGE_ENABLED_OPENMP 1

#include <Eigen/Eigen>
#include <FEM/IPC.h>
#include <FEM/FRICTION.h>

using namespace JGSL;
typedef double T;

namespace {
struct RefScene {
    int nV = 0;
    MESH_NODE<T, 3> X;
    MESH_NODE_ATTR<T, 3> nodeAttr;
    std::vector<int> boundaryNode, particle;
    std::vector<VECTOR<int, 2>> boundaryEdge, rod;
    std::vector<VECTOR<int, 3>> boundaryTri;
    std::map<int, std::set<int>> NNExclusion;
    std::vector<T> BNArea, BEArea, BTArea;
    VECTOR<int, 2> codimBNStartInd;
    std::vector<bool> DBCb;
};
std::vector<VECTOR<int, 4>> g_cs;
std::vector<VECTOR<T, 2>> g_info;
std::vector<Eigen::Triplet<T>> g_trip;
// friction state
std::vector<VECTOR<int, 4>> g_fcs;
std::vector<Eigen::Matrix<T, 2, 1>> g_closest;
std::vector<Eigen::Matrix<T, 3, 2>> g_basis;
std::vector<T> g_nf;

void set_cs(const int* cs, const double* info, int n, std::vector<VECTOR<int, 4>>& c, std::vector<VECTOR<T, 2>>& w)
{
    c.resize(n); w.resize(n);
    for (int i = 0; i < n; ++i) {
        c[i] = VECTOR<int, 4>(cs[4 * i], cs[4 * i + 1], cs[4 * i + 2], cs[4 * i + 3]);
        if (info) w[i] = VECTOR<T, 2>(info[2 * i], info[2 * i + 1]);
    }
}
void set_X(MESH_NODE<T, 3>& X, const double* x, int nV)
{
    for (int i = 0; i < nV; ++i) std::get<0>(X.Get_Unchecked(i)) = VECTOR<T, 3>(x[3 * i], x[3 * i + 1], x[3 * i + 2]);
}
} // namespace

extern "C" {

void* ref_scene_create(int nV, const double* X, const double* X0, int nBN, const int* BN, int nBE, const int* BE,
    int nBT, const int* BT, int nRod, int codim0, int codim1, const uint8_t* DBC, int nnxPairs, const int* nnxPairList,
    const double* BNArea, const double* BEArea, const double* BTArea)
{
    RefScene* s = new RefScene();
    s->nV = nV;
    s->X.Reserve(nV); s->nodeAttr.Reserve(nV);
    for (int i = 0; i < nV; ++i) {
        s->X.Append(VECTOR<T, 3>(X[3 * i], X[3 * i + 1], X[3 * i + 2]));
        s->nodeAttr.Append(VECTOR<T, 3>(X0[3 * i], X0[3 * i + 1], X0[3 * i + 2]), VECTOR<T, 3>(0, 0, 0), VECTOR<T, 3>(0, 0, 0), T(1));
    }
    s->boundaryNode.assign(BN, BN + nBN);
    for (int i = 0; i < nBE; ++i) s->boundaryEdge.emplace_back(BE[2 * i], BE[2 * i + 1]);
    for (int i = 0; i < nBT; ++i) s->boundaryTri.emplace_back(BT[3 * i], BT[3 * i + 1], BT[3 * i + 2]);
    // only rod.size() is read on the path (IPC.h:276): rod edges are the last nRod boundary edges
    for (int i = 0; i < nRod; ++i) s->rod.emplace_back(BE[2 * (nBE - nRod + i)], BE[2 * (nBE - nRod + i) + 1]);
    s->codimBNStartInd = VECTOR<int, 2>(codim0, codim1);
    s->DBCb.assign(nV, false);
    for (int i = 0; i < nV; ++i) s->DBCb[i] = DBC[i] != 0;
    for (int i = 0; i < nnxPairs; ++i) s->NNExclusion[nnxPairList[2 * i]].insert(nnxPairList[2 * i + 1]);
    if (BNArea) s->BNArea.assign(BNArea, BNArea + nBN); else s->BNArea.assign(nBN, 1.0);
    if (BEArea) s->BEArea.assign(BEArea, BEArea + nBE); else s->BEArea.assign(nBE, 1.0);
    if (BTArea) s->BTArea.assign(BTArea, BTArea + nBT); else s->BTArea.assign(nBT, 1.0);
    return s;
}
void ref_scene_set_X(void* h, const double* X) { RefScene* s = (RefScene*)h; set_X(s->X, X, s->nV); }
void ref_scene_destroy(void* h) { delete (RefScene*)h; }
int ref_num_threads() { return omp_get_max_threads(); }
void ref_set_num_threads(int n) { omp_set_num_threads(n); }

// seconds accumulated per TIMER_FLAG scope since the last reset; returns the value of `name` (0 if never entered)
double ref_timer(const char* name) { auto it = cipc_ref_timer::g_scopes.find(name); return it == cipc_ref_timer::g_scopes.end() ? 0.0 : it->second; }
void ref_timer_reset() { cipc_ref_timer::g_scopes.clear(); }

int ref_constraint_set(void* h, int elastic, double dHat2, double thickness, int /*use_hash*/, double* /*timers4*/)
{
    RefScene* s = (RefScene*)h;
    std::vector<VECTOR<int, 2>> cs_PTEE;
    if (elastic)
        Compute_Constraint_Set<T, 3, false, true>(s->X, s->nodeAttr, s->boundaryNode, s->boundaryEdge, s->boundaryTri, s->particle, s->rod,
            s->NNExclusion, s->BNArea, s->BEArea, s->BTArea, s->codimBNStartInd, s->DBCb, dHat2, thickness, false, g_cs, cs_PTEE, g_info);
    else
        Compute_Constraint_Set<T, 3, false, false>(s->X, s->nodeAttr, s->boundaryNode, s->boundaryEdge, s->boundaryTri, s->particle, s->rod,
            s->NNExclusion, s->BNArea, s->BEArea, s->BTArea, s->codimBNStartInd, s->DBCb, dHat2, thickness, false, g_cs, cs_PTEE, g_info);
    return (int)g_cs.size();
}
void ref_fetch_constraints(int* cs, double* info)
{
    for (size_t i = 0; i < g_cs.size(); ++i) {
        for (int k = 0; k < 4; ++k) cs[4 * i + k] = g_cs[i][k];
        info[2 * i] = g_info[i][0]; info[2 * i + 1] = g_info[i][1];
    }
}
int ref_barrier(void* h, int elastic, const int* cs, const double* info, int n, double dHat2, const double* kappa_in,
    double thickness, double* E)
{
    RefScene* s = (RefScene*)h;
    std::vector<VECTOR<int, 4>> c; std::vector<VECTOR<T, 2>> w;
    set_cs(cs, info, n, c, w);
    T kappa[3] = {kappa_in[0], kappa_in[1], kappa_in[2]};
    if (elastic) Compute_Barrier<T, 3, true>(s->X, s->nodeAttr, c, w, dHat2, kappa, thickness, *E);
    else Compute_Barrier<T, 3, false>(s->X, s->nodeAttr, c, w, dHat2, kappa, thickness, *E);
    return 0;
}
void ref_barrier_gradient(void* h, int elastic, const int* cs, const double* info, int n, double dHat2,
    const double* kappa_in, double thickness, double* g)
{
    RefScene* s = (RefScene*)h;
    std::vector<VECTOR<int, 4>> c; std::vector<VECTOR<T, 2>> w;
    set_cs(cs, info, n, c, w);
    T kappa[3] = {kappa_in[0], kappa_in[1], kappa_in[2]};
    for (int i = 0; i < s->nV; ++i)
        std::get<FIELDS<MESH_NODE_ATTR<T, 3>>::g>(s->nodeAttr.Get_Unchecked(i)) = VECTOR<T, 3>(g[3 * i], g[3 * i + 1], g[3 * i + 2]);
    if (elastic) Compute_Barrier_Gradient<T, 3, true>(s->X, c, w, dHat2, kappa, thickness, s->nodeAttr);
    else Compute_Barrier_Gradient<T, 3, false>(s->X, c, w, dHat2, kappa, thickness, s->nodeAttr);
    for (int i = 0; i < s->nV; ++i) {
        const VECTOR<T, 3>& gi = std::get<FIELDS<MESH_NODE_ATTR<T, 3>>::g>(s->nodeAttr.Get_Unchecked(i));
        g[3 * i] = gi[0]; g[3 * i + 1] = gi[1]; g[3 * i + 2] = gi[2];
    }
}
long ref_barrier_hessian(void* h, int elastic, const int* cs, const double* info, int n, double dHat2,
    const double* kappa_in, double thickness, int projectSPD)
{
    RefScene* s = (RefScene*)h;
    std::vector<VECTOR<int, 4>> c; std::vector<VECTOR<T, 2>> w;
    set_cs(cs, info, n, c, w);
    T kappa[3] = {kappa_in[0], kappa_in[1], kappa_in[2]};
    g_trip.clear();
    if (elastic) Compute_Barrier_Hessian<T, 3, true>(s->X, s->nodeAttr, c, w, dHat2, kappa, thickness, projectSPD != 0, g_trip);
    else Compute_Barrier_Hessian<T, 3, false>(s->X, s->nodeAttr, c, w, dHat2, kappa, thickness, projectSPD != 0, g_trip);
    return (long)g_trip.size();
}
void ref_fetch_triplets(int* rows, int* cols, double* vals)
{
    for (size_t i = 0; i < g_trip.size(); ++i) { rows[i] = g_trip[i].row(); cols[i] = g_trip[i].col(); vals[i] = g_trip[i].value(); }
}
// the reference's triplet vector in place (16-byte {int,int,double} records): full-size parity checks compare it with the
// CUDA path's stream without a copy (oracle_compare_triplet_blocks)
const void* ref_triplets_data(long* n) { static_assert(sizeof(Eigen::Triplet<T>) == 16, "triplet layout"); if (n) *n = (long)g_trip.size(); return g_trip.data(); }
void ref_triplets_release() { std::vector<Eigen::Triplet<T>>().swap(g_trip); }
int ref_step_size(void* h, int elastic, const double* searchDir, double thickness, int /*use_hash*/, double* stepSize,
    double* /*timers3*/, long* nPairs)
{
    RefScene* s = (RefScene*)h;
    std::vector<T> p(searchDir, searchDir + 3 * (size_t)s->nV);
    if (elastic)
        Compute_Intersection_Free_StepSize<T, 3, false, true>(s->X, s->boundaryNode, s->boundaryEdge, s->boundaryTri, s->particle, s->rod,
            s->NNExclusion, s->codimBNStartInd, s->DBCb, p, thickness, *stepSize);
    else
        Compute_Intersection_Free_StepSize<T, 3, false, false>(s->X, s->boundaryNode, s->boundaryEdge, s->boundaryTri, s->particle, s->rod,
            s->NNExclusion, s->codimBNStartInd, s->DBCb, p, thickness, *stepSize);
    if (nPairs) *nPairs = -1;
    return 0;
}
void ref_min_dist2(void* h, const int* cs, int n, double thickness, double* dist2, double* minDist2)
{
    RefScene* s = (RefScene*)h;
    std::vector<VECTOR<int, 4>> c; std::vector<VECTOR<T, 2>> w;
    set_cs(cs, nullptr, n, c, w);
    std::vector<T> d;
    Compute_Min_Dist2<T, 3>(s->X, c, thickness, d, *minDist2);
    std::memcpy(dist2, d.data(), sizeof(double) * n);
}

// ---- friction (FEM/FRICTION.h)
int ref_friction_basis(void* h, int elastic, const int* cs, const double* info, int n, double dHat2, const double* kappa_in, double thickness)
{
    RefScene* s = (RefScene*)h;
    std::vector<VECTOR<int, 4>> c; std::vector<VECTOR<T, 2>> w;
    set_cs(cs, info, n, c, w);
    T kappa[3] = {kappa_in[0], kappa_in[1], kappa_in[2]};
    if (elastic) Compute_Friction_Basis<T, 3, true>(s->X, c, w, g_fcs, g_closest, g_basis, g_nf, dHat2, kappa, thickness);
    else Compute_Friction_Basis<T, 3, false>(s->X, c, w, g_fcs, g_closest, g_basis, g_nf, dHat2, kappa, thickness);
    return (int)g_fcs.size();
}
void ref_fetch_friction(int* cs, double* closest, double* basis, double* nf)
{
    for (size_t i = 0; i < g_fcs.size(); ++i) {
        for (int k = 0; k < 4; ++k) cs[4 * i + k] = g_fcs[i][k];
        closest[2 * i] = g_closest[i][0]; closest[2 * i + 1] = g_closest[i][1];
        for (int j = 0; j < 2; ++j) for (int a = 0; a < 3; ++a) basis[6 * i + 3 * j + a] = g_basis[i](a, j);
        nf[i] = g_nf[i];
    }
}
void ref_set_friction(const int* cs, const double* closest, const double* basis, const double* nf, int n)
{
    g_fcs.resize(n); g_closest.resize(n); g_basis.resize(n); g_nf.resize(n);
    for (int i = 0; i < n; ++i) {
        g_fcs[i] = VECTOR<int, 4>(cs[4 * i], cs[4 * i + 1], cs[4 * i + 2], cs[4 * i + 3]);
        g_closest[i][0] = closest[2 * i]; g_closest[i][1] = closest[2 * i + 1];
        for (int j = 0; j < 2; ++j) for (int a = 0; a < 3; ++a) g_basis[i](a, j) = basis[6 * i + 3 * j + a];
        g_nf[i] = nf[i];
    }
}
double ref_friction_coef(int nComp, const int* compNodeRange, const double* muComp)
{
    std::vector<int> r(compNodeRange, compNodeRange + nComp);
    std::vector<T> m(muComp, muComp + (size_t)nComp * nComp);
    T mu = 0;
    Compute_Friction_Coef<T, 3>(g_fcs, r, m, g_nf, mu);
    return mu;
}
void ref_friction_potential(void* h, const double* Xn_in, double epsvh2, double mu, double* E)
{
    RefScene* s = (RefScene*)h;
    MESH_NODE<T, 3> Xn(s->nV);
    for (int i = 0; i < s->nV; ++i) Xn.Append(VECTOR<T, 3>(Xn_in[3 * i], Xn_in[3 * i + 1], Xn_in[3 * i + 2]));
    Compute_Friction_Potential<T, 3>(s->X, Xn, g_fcs, g_closest, g_basis, g_nf, epsvh2, mu, *E);
}
void ref_friction_gradient(void* h, const double* Xn_in, double epsvh2, double mu, double* g)
{
    RefScene* s = (RefScene*)h;
    MESH_NODE<T, 3> Xn(s->nV);
    for (int i = 0; i < s->nV; ++i) Xn.Append(VECTOR<T, 3>(Xn_in[3 * i], Xn_in[3 * i + 1], Xn_in[3 * i + 2]));
    for (int i = 0; i < s->nV; ++i)
        std::get<FIELDS<MESH_NODE_ATTR<T, 3>>::g>(s->nodeAttr.Get_Unchecked(i)) = VECTOR<T, 3>(g[3 * i], g[3 * i + 1], g[3 * i + 2]);
    Compute_Friction_Gradient<T, 3>(s->X, Xn, g_fcs, g_closest, g_basis, g_nf, epsvh2, mu, s->nodeAttr);
    for (int i = 0; i < s->nV; ++i) {
        const VECTOR<T, 3>& gi = std::get<FIELDS<MESH_NODE_ATTR<T, 3>>::g>(s->nodeAttr.Get_Unchecked(i));
        g[3 * i] = gi[0]; g[3 * i + 1] = gi[1]; g[3 * i + 2] = gi[2];
    }
}
long ref_friction_hessian(void* h, const double* Xn_in, double epsvh2, double mu, int projectSPD)
{
    RefScene* s = (RefScene*)h;
    MESH_NODE<T, 3> Xn(s->nV);
    for (int i = 0; i < s->nV; ++i) Xn.Append(VECTOR<T, 3>(Xn_in[3 * i], Xn_in[3 * i + 1], Xn_in[3 * i + 2]));
    g_trip.clear();
    Compute_Friction_Hessian<T, 3>(s->X, Xn, g_fcs, g_closest, g_basis, g_nf, epsvh2, mu, projectSPD != 0, g_trip);
    return (long)g_trip.size();
}

} // extern "C"
