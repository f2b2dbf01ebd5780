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
#include <algorithm>
using std::log;

// TIMER_FLAG of Utils/PROFILER.h (needs Boost through Utils/PARAMETER.h).
//   plain build: scope seconds by name.
//   CIPC_SHIM_BUILD (tests/shim_harness/shim_drivers.cpp compiles this file through codim-ipc_b200/shim): a scope tree with
//   the containers of the reference's JGSL::TIMER namespace (same names and types, Utils/PROFILER.h:34-40), so that the
//   shim's own add_profiler_scope -- written against the reference's profiler -- is the code that runs here.
#ifdef CIPC_SHIM_BUILD
namespace JGSL { namespace TIMER {
std::map<std::pair<std::string, int>, int> name_scope;
std::vector<std::pair<std::string, int>> scope_name(1, std::make_pair("Global", -1));
std::vector<std::chrono::duration<double>> scope_duration(1, std::chrono::duration<double>(0));
std::vector<std::chrono::duration<double>> global_duration(1, std::chrono::duration<double>(0));
std::vector<std::vector<int>> scope_edges(1, std::vector<int>());
std::vector<int> scope_stack(1, 0);
struct ScopedTimer { // opens (name, parent) as a child of the innermost open scope; adds its wall time on exit
    int id;
    std::chrono::steady_clock::time_point t0;
    explicit ScopedTimer(const std::string& name) : t0(std::chrono::steady_clock::now())
    {
        const auto key = std::make_pair(name, scope_stack.back());
        const auto it = name_scope.find(key);
        if (it != name_scope.end()) id = it->second;
        else {
            id = (int)scope_name.size();
            name_scope[key] = id;
            scope_name.push_back(key);
            scope_duration.emplace_back(0);
            global_duration.emplace_back(0);
            scope_edges.emplace_back();
            scope_edges[scope_stack.back()].push_back(id);
        }
        scope_stack.push_back(id);
    }
    ~ScopedTimer() { scope_duration[id] += std::chrono::steady_clock::now() - t0; scope_stack.pop_back(); }
};
} }
#define TIMER_FLAG(name) JGSL::TIMER::ScopedTimer scoped_timer(name)
#else
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
#endif
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
#ifdef CIPC_SHIM_BUILD
double ref_timer(const char* name)
{
    double t = 0;
    for (size_t i = 0; i < TIMER::scope_name.size(); ++i)
        if (TIMER::scope_name[i].first == name) t += TIMER::scope_duration[i].count();
    return t;
}
// name of the parent scope of `name` ("" if unknown): the sub-scopes must hang under their top-level scope
const char* ref_timer_parent(const char* name)
{
    for (size_t i = 1; i < TIMER::scope_name.size(); ++i)
        if (TIMER::scope_name[i].first == name) return TIMER::scope_name[TIMER::scope_name[i].second].first.c_str();
    return "";
}
void ref_timer_reset() { for (auto& d : TIMER::scope_duration) d = std::chrono::duration<double>(0); }
#else
double ref_timer(const char* name) { auto it = cipc_ref_timer::g_scopes.find(name); return it == cipc_ref_timer::g_scopes.end() ? 0.0 : it->second; }
const char* ref_timer_parent(const char*) { return ""; }
void ref_timer_reset() { cipc_ref_timer::g_scopes.clear(); }
#endif

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


// One contact stage of a Newton iteration in the reference's own call pattern (Shell/IMPLICIT_EULER.h:418-428,464,495,
// 97,118-123,597-600; INC_POTENTIAL.h:321,374): constraintSet / stencilInfo persist, the triplet vector and dist2 are fresh
// locals of every call.  times[7] = seconds of Compute_Constraint_Set, Compute_Barrier, _Gradient, _Hessian,
// Compute_Intersection_Free_StepSize, 2 x Compute_Min_Dist2 (calls only); results[6] = E, step, minDist2, nC, nTriplets,
// sum of the triplet values.  In the plain build this runs the reference's CPU templates, in the shim build the CUDA path.
void ref_contact_stage(void* h, double dHat2, const double* kappa_in, double thickness, const double* searchDir, double* times, double* results)
{
    RefScene* s = (RefScene*)h;
    T kappa[3] = {kappa_in[0], kappa_in[1], kappa_in[2]};
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto sec = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) { return std::chrono::duration<double>(b - a).count(); };
    std::vector<VECTOR<int, 2>> cs_PTEE;
    static std::vector<T> p;
    p.assign(searchDir, searchDir + 3 * (size_t)s->nV);
    auto t0 = now();
    Compute_Constraint_Set<T, 3, false, false>(s->X, s->nodeAttr, s->boundaryNode, s->boundaryEdge, s->boundaryTri, s->particle, s->rod,
        s->NNExclusion, s->BNArea, s->BEArea, s->BTArea, s->codimBNStartInd, s->DBCb, dHat2, thickness, false, g_cs, cs_PTEE, g_info);
    auto t1 = now();
    T E = 0;
    Compute_Barrier<T, 3, false>(s->X, s->nodeAttr, g_cs, g_info, dHat2, kappa, thickness, E);
    auto t2 = now();
    Compute_Barrier_Gradient<T, 3, false>(s->X, g_cs, g_info, dHat2, kappa, thickness, s->nodeAttr);
    auto t3 = now();
    double nTrip = 0, tsum = 0, tH;
    {
        std::vector<Eigen::Triplet<T>> triplets; // INC_POTENTIAL.h:321: a fresh vector per Hessian evaluation
        auto a = now();
        Compute_Barrier_Hessian<T, 3, false>(s->X, s->nodeAttr, g_cs, g_info, dHat2, kappa, thickness, true, triplets);
        tH = sec(a, now());
        nTrip = (double)triplets.size();
        const size_t step = triplets.size() / 4096 + 1;
        for (size_t i = 0; i < triplets.size(); i += step) tsum += triplets[i].value();
    }
    auto t4 = now();
    T alpha = 1;
    Compute_Intersection_Free_StepSize<T, 3, false, false>(s->X, s->boundaryNode, s->boundaryEdge, s->boundaryTri, s->particle, s->rod,
        s->NNExclusion, s->codimBNStartInd, s->DBCb, p, thickness, alpha);
    auto t5 = now();
    T minDist2 = 0;
    double tm[2] = {0, 0};
    for (int k = 0; k < 2; ++k) {
        std::vector<T> dist2; // IMPLICIT_EULER.h:122,597
        auto a = now();
        if (!g_cs.empty()) Compute_Min_Dist2<T, 3>(s->X, g_cs, thickness, dist2, minDist2);
        tm[k] = sec(a, now());
    }
    times[0] = sec(t0, t1); times[1] = sec(t1, t2); times[2] = sec(t2, t3); times[3] = tH; times[4] = sec(t4, t5); times[5] = tm[0]; times[6] = tm[1];
    results[0] = E; results[1] = alpha; results[2] = minDist2; results[3] = (double)g_cs.size(); results[4] = nTrip; results[5] = tsum;
}

#ifdef CIPC_SHIM_BUILD
// GPU (shim) and CPU (the reference's own templates, renamed *_CPU by the shim) side by side in ONE binary, on the real
// MESH_NODE / MESH_NODE_ATTR storage; compared here in C++.  rawTriplets selects what the Hessian comparison looks at: the
// process must run with CIPC_TRIPLETS=raw for the per-block comparison (out[4], out[5]); with merged delivery out[4] is the
// relative difference of the assembled matrices' action on a fixed vector instead.
// out: [0] constraint-set mismatches (sorted), [1] |E - E_cpu| / |E_cpu|, [2] max |g - g_cpu| / max |g_cpu|, [3] triplet count
// difference (raw) / 0, [4] max per-block ||H - H_cpu||_F / ||H_cpu||_F (raw) or ||A x - A_cpu x|| / ||A_cpu x|| (merged),
// [5] (row, col) mismatches (raw), [6] step_gpu, [7] step_cpu, [8] dist2 mismatches, [9] |minDist2 - cpu|,
// [10] friction-set mismatches, [11] max closest-point / basis / normal-force error, [12] friction potential rel. error,
// [13] friction gradient rel. error, [14] friction Hessian error (as [4]), [15] nC, [16] nF,
// [17] CSR hand-off (Compute_Barrier_Hessian_CSR): rel. error of A x against the CPU template's triplets, [18] structural
// violations of the CSR arrays (row pointer length / end, unsorted or repeated or out-of-range columns)
void shim_selfcheck(void* h, double dHat2, const double* kappa_in, double thickness, const double* searchDir, const double* Xn_in, double epsvh2,
    double mu, int rawTriplets, double* out)
{
    RefScene* s = (RefScene*)h;
    const int nV = s->nV;
    T kappa[3] = {kappa_in[0], kappa_in[1], kappa_in[2]};
    typedef std::vector<VECTOR<int, 4>> CS;
    auto sorted = [](CS c) {
        std::sort(c.begin(), c.end(), [](const VECTOR<int, 4>& a, const VECTOR<int, 4>& b) {
            for (int k = 0; k < 4; ++k) if (a[k] != b[k]) return a[k] < b[k];
            return false;
        });
        return c;
    };
    auto same_cs = [&](const CS& a, const CS& b) {
        if (a.size() != b.size()) return (double)std::max(a.size(), b.size());
        const CS x = sorted(a), y = sorted(b);
        double m = 0;
        for (size_t i = 0; i < x.size(); ++i) for (int k = 0; k < 4; ++k) m += x[i][k] != y[i][k];
        return m;
    };
    // y = A x with x_k = cos(1 + 0.37 k), A given by triplets (duplicates add)
    auto matvec = [&](const std::vector<Eigen::Triplet<T>>& t) {
        std::vector<double> y(3 * (size_t)nV, 0.0);
        for (const auto& e : t) y[e.row()] += e.value() * std::cos(1.0 + 0.37 * e.col());
        return y;
    };
    auto rel_vec = [](const std::vector<double>& a, const std::vector<double>& b) {
        double d = 0, n = 0;
        for (size_t i = 0; i < a.size(); ++i) { d += (a[i] - b[i]) * (a[i] - b[i]); n += b[i] * b[i]; }
        return n > 0 ? std::sqrt(d / n) : std::sqrt(d);
    };
    auto block_err = [&](const CS& cs, const std::vector<Eigen::Triplet<T>>& a, const std::vector<Eigen::Triplet<T>>& b, double& idxMism) {
        double worst = 0;
        size_t o = 0;
        idxMism = 0;
        for (size_t i = 0; i < cs.size(); ++i) {
            const int n = (cs[i][0] >= 0 || cs[i][3] >= 0) ? 144 : (cs[i][2] >= 0 ? 81 : 36);
            double d2 = 0, b2 = 0;
            for (int k = 0; k < n; ++k) {
                idxMism += a[o + k].row() != b[o + k].row() || a[o + k].col() != b[o + k].col();
                const double d = a[o + k].value() - b[o + k].value();
                d2 += d * d; b2 += b[o + k].value() * b[o + k].value();
            }
            if (b2 > 0) worst = std::max(worst, std::sqrt(d2 / b2));
            o += n;
        }
        return worst;
    };
    auto get_g = [&]() {
        std::vector<double> g(3 * (size_t)nV);
        for (int i = 0; i < nV; ++i) {
            const VECTOR<T, 3>& gi = std::get<FIELDS<MESH_NODE_ATTR<T, 3>>::g>(s->nodeAttr.Get_Unchecked(i));
            g[3 * i] = gi[0]; g[3 * i + 1] = gi[1]; g[3 * i + 2] = gi[2];
        }
        return g;
    };
    auto zero_g = [&]() {
        for (int i = 0; i < nV; ++i) std::get<FIELDS<MESH_NODE_ATTR<T, 3>>::g>(s->nodeAttr.Get_Unchecked(i)) = VECTOR<T, 3>(0, 0, 0);
    };
    auto max_abs = [](const std::vector<double>& a) { double m = 0; for (double v : a) m = std::max(m, std::fabs(v)); return m; };
    auto max_diff = [](const std::vector<double>& a, const std::vector<double>& b) { double m = 0; for (size_t i = 0; i < a.size(); ++i) m = std::max(m, std::fabs(a[i] - b[i])); return m; };
    for (int k = 0; k < 19; ++k) out[k] = 0;

    // ---- the six contact templates
    CS csG, csC;
    std::vector<VECTOR<T, 2>> infoG, infoC;
    std::vector<VECTOR<int, 2>> ptee;
    Compute_Constraint_Set<T, 3, false, false>(s->X, s->nodeAttr, s->boundaryNode, s->boundaryEdge, s->boundaryTri, s->particle, s->rod, s->NNExclusion,
        s->BNArea, s->BEArea, s->BTArea, s->codimBNStartInd, s->DBCb, dHat2, thickness, false, csG, ptee, infoG);
    Compute_Constraint_Set_CPU<T, 3, false, false>(s->X, s->nodeAttr, s->boundaryNode, s->boundaryEdge, s->boundaryTri, s->particle, s->rod, s->NNExclusion,
        s->BNArea, s->BEArea, s->BTArea, s->codimBNStartInd, s->DBCb, dHat2, thickness, false, csC, ptee, infoC);
    out[0] = same_cs(csG, csC);
    out[15] = (double)csG.size();
    // both sides continue on the GPU's set (its order differs from the CPU's unordered_set order)
    T EG = 0.25, EC = 0.25;
    Compute_Barrier<T, 3, false>(s->X, s->nodeAttr, csG, infoG, dHat2, kappa, thickness, EG);
    Compute_Barrier_CPU<T, 3, false>(s->X, s->nodeAttr, csG, infoG, dHat2, kappa, thickness, EC);
    out[1] = std::fabs(EG - EC) / std::max(std::fabs(EC), 1e-300);
    zero_g(); Compute_Barrier_Gradient<T, 3, false>(s->X, csG, infoG, dHat2, kappa, thickness, s->nodeAttr);
    const std::vector<double> gG = get_g();
    zero_g(); Compute_Barrier_Gradient_CPU<T, 3, false>(s->X, csG, infoG, dHat2, kappa, thickness, s->nodeAttr);
    const std::vector<double> gC = get_g();
    out[2] = max_diff(gG, gC) / std::max(max_abs(gC), 1e-300);
    {
        std::vector<Eigen::Triplet<T>> tG(3, Eigen::Triplet<T>(0, 0, 0.0)), tC(3, Eigen::Triplet<T>(0, 0, 0.0)); // appended behind existing entries
        Compute_Barrier_Hessian<T, 3, false>(s->X, s->nodeAttr, csG, infoG, dHat2, kappa, thickness, true, tG);
        Compute_Barrier_Hessian_CPU<T, 3, false>(s->X, s->nodeAttr, csG, infoG, dHat2, kappa, thickness, true, tC);
        if (rawTriplets) {
            out[3] = (double)tG.size() - (double)tC.size();
            if (tG.size() == tC.size()) {
                std::vector<Eigen::Triplet<T>> a(tG.begin() + 3, tG.end()), b(tC.begin() + 3, tC.end());
                out[4] = block_err(csG, a, b, out[5]);
            }
        }
        else out[4] = rel_vec(matvec(tG), matvec(tC));
        // optional CSR hand-off: the same matrix, assembled on the device
        std::vector<int> ptr, col;
        std::vector<T> val;
        Compute_Barrier_Hessian_CSR<T, 3, false>(s->X, s->nodeAttr, csG, infoG, dHat2, kappa, thickness, true, ptr, col, val);
        std::vector<Eigen::Triplet<T>> tRef(tC.begin() + 3, tC.end());
        std::vector<double> y(3 * (size_t)nV, 0.0);
        double bad = (ptr.size() != 3 * (size_t)nV + 1) ? 1 : 0;
        if (!bad) {
            bad += ptr[0] != 0 || (size_t)ptr.back() != val.size() || col.size() != val.size();
            for (int r = 0; r < 3 * nV && !bad; ++r)
                for (int k = ptr[r]; k < ptr[r + 1]; ++k) {
                    if (col[k] < 0 || col[k] >= 3 * nV || (k > ptr[r] && col[k] <= col[k - 1])) { ++bad; break; }
                    y[r] += val[k] * std::cos(1.0 + 0.37 * col[k]);
                }
        }
        out[17] = rel_vec(y, matvec(tRef));
        out[18] = bad;
    }
    std::vector<T> p(searchDir, searchDir + 3 * (size_t)nV);
    T aG = 1, aC = 1;
    Compute_Intersection_Free_StepSize<T, 3, false, false>(s->X, s->boundaryNode, s->boundaryEdge, s->boundaryTri, s->particle, s->rod, s->NNExclusion,
        s->codimBNStartInd, s->DBCb, p, thickness, aG);
    Compute_Intersection_Free_StepSize_CPU<T, 3, false, false>(s->X, s->boundaryNode, s->boundaryEdge, s->boundaryTri, s->particle, s->rod, s->NNExclusion,
        s->codimBNStartInd, s->DBCb, p, thickness, aC);
    out[6] = aG; out[7] = aC;
    if (!csG.empty()) {
        std::vector<T> dG, dC;
        T mG = 0, mC = 0;
        Compute_Min_Dist2<T, 3>(s->X, csG, thickness, dG, mG);
        Compute_Min_Dist2_CPU<T, 3>(s->X, csG, thickness, dC, mC);
        for (size_t i = 0; i < dC.size(); ++i) out[8] += dG[i] != dC[i];
        out[9] = std::fabs(mG - mC);
    }
    // ---- the five friction templates
    if (!Xn_in || csG.empty()) return;
    MESH_NODE<T, 3> Xn(nV);
    for (int i = 0; i < nV; ++i) Xn.Append(VECTOR<T, 3>(Xn_in[3 * i], Xn_in[3 * i + 1], Xn_in[3 * i + 2]));
    CS fG, fC;
    std::vector<Eigen::Matrix<T, 2, 1>> cpG, cpC;
    std::vector<Eigen::Matrix<T, 3, 2>> tbG, tbC;
    std::vector<T> nfG, nfC;
    Compute_Friction_Basis<T, 3, false>(s->X, csG, infoG, fG, cpG, tbG, nfG, dHat2, kappa, thickness);
    Compute_Friction_Basis_CPU<T, 3, false>(s->X, csG, infoG, fC, cpC, tbC, nfC, dHat2, kappa, thickness);
    out[16] = (double)fG.size();
    out[10] = (fG.size() != fC.size()) ? (double)std::max(fG.size(), fC.size()) : 0;
    if (out[10] == 0) {
        double nfScale = 0;
        for (size_t i = 0; i < fC.size(); ++i) nfScale = std::max(nfScale, std::fabs(nfC[i]));
        for (size_t i = 0; i < fC.size(); ++i) {
            for (int k = 0; k < 4; ++k) out[10] += fG[i][k] != fC[i][k];
            const bool pp = fC[i][0] < 0 && fC[i][2] < 0, pe = fC[i][0] < 0 && fC[i][2] >= 0 && fC[i][3] < 0; // unused closest-point slots stay uninitialised in the reference
            if (!pp) out[11] = std::max(out[11], std::fabs(cpG[i][0] - cpC[i][0]));
            if (!pp && !pe) out[11] = std::max(out[11], std::fabs(cpG[i][1] - cpC[i][1]));
            for (int r = 0; r < 3; ++r) for (int c = 0; c < 2; ++c) out[11] = std::max(out[11], std::fabs(tbG[i](r, c) - tbC[i](r, c)));
            out[11] = std::max(out[11], std::fabs(nfG[i] - nfC[i]) / std::max(nfScale, 1e-300));
        }
    }
    T FG = 0, FC = 0;
    Compute_Friction_Potential<T, 3>(s->X, Xn, fG, cpG, tbG, nfG, epsvh2, mu, FG);
    Compute_Friction_Potential_CPU<T, 3>(s->X, Xn, fG, cpG, tbG, nfG, epsvh2, mu, FC);
    out[12] = std::fabs(FG - FC) / std::max(std::fabs(FC), 1e-300);
    zero_g(); Compute_Friction_Gradient<T, 3>(s->X, Xn, fG, cpG, tbG, nfG, epsvh2, mu, s->nodeAttr);
    const std::vector<double> hG = get_g();
    zero_g(); Compute_Friction_Gradient_CPU<T, 3>(s->X, Xn, fG, cpG, tbG, nfG, epsvh2, mu, s->nodeAttr);
    const std::vector<double> hC = get_g();
    out[13] = max_diff(hG, hC) / std::max(max_abs(hC), 1e-300);
    {
        std::vector<Eigen::Triplet<T>> tG, tC;
        Compute_Friction_Hessian<T, 3>(s->X, Xn, fG, cpG, tbG, nfG, epsvh2, mu, true, tG);
        Compute_Friction_Hessian_CPU<T, 3>(s->X, Xn, fG, cpG, tbG, nfG, epsvh2, mu, true, tC);
        double idx = 0;
        if (rawTriplets && tG.size() == tC.size()) { out[14] = block_err(fG, tG, tC, idx); out[14] += idx; }
        else out[14] = rel_vec(matvec(tG), matvec(tC));
    }
}
#endif

} // extern "C"
