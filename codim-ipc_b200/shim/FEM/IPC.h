// shim/FEM/IPC.h -- drop-in replacement of the reference's Library/FEM/IPC.h contact entry points.
//
// Build the reference with `-I <this repo>/codim-ipc_b200/shim -I <this repo>/include` placed BEFORE
// `-I Library` (CMakeLists.txt:24) and link libcipc_b200.so: every `#include <FEM/IPC.h>` of the reference
// (FEM/FEM_EXPORTER.h:10, FEM/TimeStepper/IMPLICIT_EULER.h:6, ADMM.h:6, SHAPE_UP.h:6) then lands here.
// No reference source is edited and every caller / pybind export compiles unchanged.
//
// How it works: the reference's own header is pulled in underneath with its six contact templates
// renamed to *_CPU (so everything else it defines -- Compute_Inversion_Free_StepSize, the root finders,
// Check_Edge_Tri_Intersect, the 2-D branches -- stays available), and six templates with the reference's
// exact names and signatures are defined on top:
//
//   <double, dim=3, shell=false, elasticIPC=false>  -> CUDA path through the C ABI (include/cipc_b200.h).
//        This is the instantiation every Projects/FEMShell scene runs (initialize_OIPC).  Any CUDA
//        failure prints a message and exit(-1)s like the reference's own invariant checks; there is
//        no CPU fallback for this instantiation.
//   anything else (float, dim=2, shell pairing, elasticIPC) -> the reference's own CPU template, untouched
//        (those instantiations are outside the scope of the graft).
//
// Reference signatures: FEM/IPC.h:19-36, 742-748, 943-948, 1258-1265, 1879-1890, 2246-2249.
//
// Environment knobs (read once): CIPC_DEVICE (CUDA ordinal, default 0) or CIPC_DEVICES (comma-separated ordinals: all of them
// behind this one calling thread, cipc_create_multi); CIPC_TRIPLETS = merged (default) | raw; CIPC_MALLOC_REUSE=1 (see state()).
//   merged: Compute_Barrier_Hessian / Compute_Friction_Hessian append ONE triplet per distinct (row, col) of their matrix --
//           the duplicates that the only consumer of the vector, sysMtr.Construct_From_Triplet = Eigen setFromTriplets
//           (Shell/INC_POTENTIAL.h:382, Math/CSR_MATRIX.h:49-56), would sum anyway are summed on the device (cipc_*_hessian_merged).
//           The assembled system matrix is the same to summation order (<= 1e-9); 0.6 GB instead of 14.5 GB cross PCIe at 1M
//           triangles and Eigen sorts 119M instead of 909M triplets.
//   raw:    the reference's exact 144 / 81 / 36 triplets per stencil, in constraint order (what the parity tests compare).
// Timer tree: the reference's sub-scopes (Compute_Constraint_Set_Build_Hash / _PT / _EE / _Merge,
// Compute_Intersection_Free_StepSize_Build_Hash / _PT / _EE; IPC.h:44,148,361,571,1903,1959,2170) are filled from the
// device-side CUDA-event times of the stages (cipc_stage_ms) under the top-level scope, see add_profiler_scope below.
#pragma once

#define Compute_Constraint_Set Compute_Constraint_Set_CPU
#define Compute_Barrier Compute_Barrier_CPU
#define Compute_Barrier_Gradient Compute_Barrier_Gradient_CPU
#define Compute_Barrier_Hessian Compute_Barrier_Hessian_CPU
#define Compute_Intersection_Free_StepSize Compute_Intersection_Free_StepSize_CPU
#define Compute_Min_Dist2 Compute_Min_Dist2_CPU
#include_next <FEM/IPC.h>
#undef Compute_Constraint_Set
#undef Compute_Barrier
#undef Compute_Barrier_Gradient
#undef Compute_Barrier_Hessian
#undef Compute_Intersection_Free_StepSize
#undef Compute_Min_Dist2

#include <cipc_b200.h>

#include <malloc.h>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <set>
#include <string>
#include <type_traits>
#include <vector>

namespace JGSL {
namespace cipc_shim {

inline void die(cipc_ctx* ctx, int st, const char* where)
{
    if (st == CIPC_OK) return;
    if (st == CIPC_ERR_NONPOSITIVE_DIST) printf("non-positive distance detected during barrier evaluation!\n"); // IPC.h:773-776
    else if (st == CIPC_ERR_ZERO_STEP) printf("zero step size from CCD!\n");                                    // IPC.h:2014-2032
    else printf("cipc_b200: %s failed with status %d: %s\n", where, st, ctx ? cipc_last_error(ctx) : "no context");
    exit(-1);
}

struct State {
    cipc_ctx* ctx = nullptr;
    bool merged = true; // CIPC_TRIPLETS
    // identity of the constraint set that is resident on the device (what Compute_Constraint_Set last handed out):
    // address, size and a 64-bit hash of the whole array (cipc_hash_bytes, threaded: ~2 ms per 100 MB)
    const void* csPtr = nullptr;
    size_t csSize = 0;
    uint64_t csHash = 0;
    std::vector<double> stage3, stage2; // packed staging
    std::vector<int32_t> nnx;
    std::vector<uint8_t> dbc;
};
inline State& state()
{
    static State s;
    if (!s.ctx) {
        const char* dev = getenv("CIPC_DEVICE");
        const char* devs = getenv("CIPC_DEVICES"); // "0,1,2,3": one context over several GPUs of this box (cipc_create_multi)
        int st;
        if (devs && *devs) {
            std::vector<int> list;
            for (const char* p = devs; *p;) {
                list.push_back(atoi(p));
                while (*p && *p != ',') ++p;
                if (*p == ',') ++p;
            }
            st = cipc_create_multi((int)list.size(), list.data(), &s.ctx);
        }
        else st = cipc_create(dev ? atoi(dev) : 0, 0, 1, &s.ctx);
        if (st != CIPC_OK) { printf("cipc_b200: no CUDA device; the contact path has no CPU fallback\n"); exit(-1); }
        const char* tm = getenv("CIPC_TRIPLETS");
        s.merged = !(tm && std::string(tm) == "raw");
        // CIPC_MALLOC_REUSE=1: keep freed memory in the process (glibc: no mmap for large blocks, no trimming), so that the
        // triplet vector a Newton iteration allocates (INC_POTENTIAL.h:321, ~2 GB at 1M triangles) lands on pages the
        // previous iteration already faulted in instead of paying ~30 ms of first-touch page faults again
        const char* mr = getenv("CIPC_MALLOC_REUSE");
        if (mr && mr[0] == '1') { mallopt(M_MMAP_MAX, 0); mallopt(M_TRIM_THRESHOLD, 2147483647); }
    }
    return s;
}

// ---- the reference's timer tree (Utils/PROFILER.h:34-40,70-97): a child scope of the scope that is open right now gets the
// device time of a stage added.  Builds that replace TIMER_FLAG (tests) define CIPC_SHIM_ADD_SCOPE themselves.
#ifndef CIPC_SHIM_ADD_SCOPE
inline void add_profiler_scope(const char* name, double seconds)
{
    using namespace TIMER;
    const auto key = std::make_pair(std::string(name), scope_stack.back());
    int id;
    const auto it = name_scope.find(key);
    if (it != name_scope.end()) id = it->second;
    else { // same bookkeeping as ScopedTimer's constructor
        id = (int)scope_name.size();
        name_scope[key] = id;
        scope_name.push_back(key);
        scope_duration.emplace_back(0);
        global_duration.emplace_back(0);
        scope_edges.emplace_back();
        scope_edges[scope_stack.back()].push_back(id);
    }
    scope_duration[id] += std::chrono::duration<double>(seconds);
}
#define CIPC_SHIM_ADD_SCOPE(name, seconds) ::JGSL::cipc_shim::add_profiler_scope(name, seconds)
#endif
inline double stage_s(State& s, const char* name)
{
    const double ms = cipc_stage_ms(s.ctx, name);
    return ms > 0 ? 1e-3 * ms : 0.0;
}
// pair enumeration is one launch for all candidate kinds: its time is split by candidate counts between the reference's
// _PT scope (point queries: PT + codimensional PE / PP) and _EE scope; narrow phase / ACCD are timed per kind
inline void report_scopes(State& s, const char* base, const char* hash, const char* pairs, const char* kPT, const char* kEE, const char* merge,
    double nPT, double nEE)
{
    const double f = (nPT + nEE) > 0 ? nPT / (nPT + nEE) : 0.5, tp = stage_s(s, pairs);
    const std::string b(base);
    CIPC_SHIM_ADD_SCOPE((b + "_Build_Hash").c_str(), stage_s(s, hash));
    CIPC_SHIM_ADD_SCOPE((b + "_PT").c_str(), tp * f + stage_s(s, kPT));
    CIPC_SHIM_ADD_SCOPE((b + "_EE").c_str(), tp * (1 - f) + stage_s(s, kEE));
    if (merge) CIPC_SHIM_ADD_SCOPE((b + "_Merge").c_str(), stage_s(s, merge));
}

// Appends n records to a vector of trivially-copyable elements WITHOUT value-initialising them and returns the first new
// slot.  vector::resize would zero-fill the new range on one thread (14.5 GB of raw triplets at 1M triangles) before the
// delivery overwrites it; here the delivery threads are the first to touch the pages.  libstdc++ and libc++ both lay a
// vector out as three pointers (begin, end, end of storage); -DCIPC_SHIM_STD_RESIZE selects the portable resize().
template <class V>
inline typename V::value_type* grow_uninitialized(V& v, size_t n)
{
    typedef typename V::value_type E;
    const size_t old = v.size();
#if (defined(__GLIBCXX__) || defined(_LIBCPP_VERSION)) && !defined(CIPC_SHIM_STD_RESIZE)
    static_assert(std::is_trivially_copyable<E>::value && sizeof(V) == 3 * sizeof(void*), "three-pointer vector of plain records");
    v.reserve(old + n);
    struct Raw { E* b; E* e; E* c; };
    Raw& r = reinterpret_cast<Raw&>(v);
    r.e += n;
#else
    v.resize(old + n);
#endif
    return v.data() + old;
}
// size n with unspecified contents (every element is overwritten by the delivery that follows)
template <class V>
inline void resize_uninitialized(V& v, size_t n)
{
    v.clear();
    grow_uninitialized(v, n);
}

// runs body(i) for i in [0, n) on the library's host thread pool (element accessors of the AoSoA storage are pure address
// arithmetic, distinct i touch distinct records)
template <class F>
inline void parallel_nodes(size_t n, F body)
{
    cipc_host_parallel_for(n, 8192, [](size_t b, size_t e, void* u) { F& f = *static_cast<F*>(u); for (size_t i = b; i < e; ++i) f(i); }, &body);
}

// positions: MESH_NODE<T,3> = BASE_STORAGE<VECTOR<T,3>>; element i is 4 contiguous doubles.  When the
// elements are contiguous (32-byte stride, the Cabana AoSoA of a single VECTOR member) upload in place,
// otherwise gather through Get_Unchecked.
template <class NODES>
inline void upload_positions(State& s, NODES& X, int (*setter)(cipc_ctx*, const double*, int), const char* what)
{
    const size_t n = X.size;
    const double* p0 = std::get<0>(X.Get_Unchecked(0)).data;
    bool contiguous = true;
    const size_t probe[3] = {1, n / 2, n - 1};
    for (size_t k : probe)
        if (k < n && std::get<0>(X.Get_Unchecked(k)).data != p0 + 4 * k) contiguous = false;
    if (contiguous) { die(s.ctx, setter(s.ctx, p0, 32), what); return; }
    s.stage3.resize(3 * n);
    double* st = s.stage3.data();
    parallel_nodes(n, [&X, st](size_t i) {
        const double* d = std::get<0>(X.Get_Unchecked(i)).data;
        st[3 * i] = d[0]; st[3 * i + 1] = d[1]; st[3 * i + 2] = d[2];
    });
    die(s.ctx, setter(s.ctx, s.stage3.data(), 24), what);
}
template <class ATTR>
inline void upload_rest(State& s, ATTR& nodeAttr)
{
    const size_t n = nodeAttr.size;
    s.stage3.resize(3 * n);
    double* st = s.stage3.data();
    parallel_nodes(n, [&nodeAttr, st](size_t i) {
        const double* d = std::get<FIELDS<ATTR>::x0>(nodeAttr.Get_Unchecked(i)).data;
        st[3 * i] = d[0]; st[3 * i + 1] = d[1]; st[3 * i + 2] = d[2];
    });
    die(s.ctx, cipc_set_rest_positions(s.ctx, s.stage3.data(), 24), "cipc_set_rest_positions");
}
inline void upload_topology(State& s, size_t nV, const std::vector<int>& boundaryNode, const std::vector<VECTOR<int, 2>>& boundaryEdge,
    const std::vector<VECTOR<int, 3>>& boundaryTri, size_t nRod, const std::map<int, std::set<int>>& NNExclusion,
    const VECTOR<int, 2>& codimBNStartInd, const std::vector<bool>& DBCb)
{
    s.dbc.resize(nV);
    for (size_t i = 0; i < nV; ++i) s.dbc[i] = DBCb[i] ? 1 : 0;
    s.nnx.clear();
    for (const auto& kv : NNExclusion)
        for (int m : kv.second) { s.nnx.push_back(kv.first); s.nnx.push_back(m); }
    const int32_t cd[2] = {codimBNStartInd[0], codimBNStartInd[1]};
    static_assert(sizeof(VECTOR<int, 2>) == 16 && sizeof(VECTOR<int, 3>) == 16, "VECTOR<int,k> is a 16-byte record");
    die(s.ctx, cipc_set_topology(s.ctx, (int)nV, (int)boundaryNode.size(), boundaryNode.data(), (int)boundaryEdge.size(),
            boundaryEdge.empty() ? nullptr : boundaryEdge[0].data, 4, (int)boundaryTri.size(), boundaryTri.empty() ? nullptr : boundaryTri[0].data, 4,
            (int)nRod, cd, s.dbc.data(), (int)(s.nnx.size() / 2), s.nnx.data(), nullptr, nullptr, nullptr),
        "cipc_set_topology");
}
// the barrier calls receive the constraint set by const reference; skip the upload when it is (by address, size and the hash
// of its full contents) the set Compute_Constraint_Set just produced, which is what every reference call site passes
inline bool constraints_resident(State& s, const std::vector<VECTOR<int, 4>>& cs)
{
    return cs.data() == s.csPtr && cs.size() == s.csSize && (cs.empty() || cipc_hash_bytes(cs.data(), cs.size() * sizeof(VECTOR<int, 4>)) == s.csHash);
}
inline void remember_constraints(State& s, const std::vector<VECTOR<int, 4>>& cs)
{
    s.csPtr = cs.data(); s.csSize = cs.size();
    s.csHash = cs.empty() ? 0 : cipc_hash_bytes(cs.data(), cs.size() * sizeof(VECTOR<int, 4>));
}
inline void ensure_constraints(State& s, const std::vector<VECTOR<int, 4>>& cs, const std::vector<VECTOR<double, 2>>& info)
{
    if (constraints_resident(s, cs)) return;
    static_assert(sizeof(VECTOR<double, 2>) == 32 && sizeof(VECTOR<int, 4>) == 16, "VECTOR<T,dim> stores T data[4]");
    die(s.ctx, cipc_set_constraints_strided(s.ctx, cs.empty() ? nullptr : cs[0].data, info.empty() ? nullptr : info[0].data, 32, (int)cs.size()),
        "cipc_set_constraints");
    remember_constraints(s, cs); // the caller's set is the resident one now
}

template <class T, int dim, bool shell, bool elasticIPC>
constexpr bool on_gpu = std::is_same<T, double>::value && dim == 3 && !shell && !elasticIPC;

} // namespace cipc_shim

// ------------------------------------------------------------------ FEM/IPC.h:19-36
template <class T, int dim, bool shell = false, bool elasticIPC = false>
void Compute_Constraint_Set(MESH_NODE<T, dim>& X, MESH_NODE_ATTR<T, dim>& nodeAttr, const std::vector<int>& boundaryNode,
    const std::vector<VECTOR<int, 2>>& boundaryEdge, const std::vector<VECTOR<int, 3>>& boundaryTri, const std::vector<int>& particle,
    const std::vector<VECTOR<int, 2>>& rod, const std::map<int, std::set<int>>& NNExclusion, const std::vector<T>& BNArea,
    const std::vector<T>& BEArea, const std::vector<T>& BTArea, const VECTOR<int, 2>& codimBNStartInd, const std::vector<bool>& DBCb,
    T dHat2, T thickness, bool getPTEE, std::vector<VECTOR<int, dim + 1>>& constraintSet, std::vector<VECTOR<int, 2>>& cs_PTEE,
    std::vector<VECTOR<T, 2>>& stencilInfo)
{
    if constexpr (!cipc_shim::on_gpu<T, dim, shell, elasticIPC>) {
        Compute_Constraint_Set_CPU<T, dim, shell, elasticIPC>(X, nodeAttr, boundaryNode, boundaryEdge, boundaryTri, particle, rod, NNExclusion,
            BNArea, BEArea, BTArea, codimBNStartInd, DBCb, dHat2, thickness, getPTEE, constraintSet, cs_PTEE, stencilInfo);
    }
    else {
        if (getPTEE) { // false at every call site of the reference
            Compute_Constraint_Set_CPU<T, dim, shell, elasticIPC>(X, nodeAttr, boundaryNode, boundaryEdge, boundaryTri, particle, rod,
                NNExclusion, BNArea, BEArea, BTArea, codimBNStartInd, DBCb, dHat2, thickness, getPTEE, constraintSet, cs_PTEE, stencilInfo);
            return;
        }
        TIMER_FLAG("Compute_Constraint_Set");
        cipc_shim::State& s = cipc_shim::state();
        cipc_shim::upload_topology(s, X.size, boundaryNode, boundaryEdge, boundaryTri, rod.size(), NNExclusion, codimBNStartInd, DBCb);
        cipc_shim::upload_positions(s, X, cipc_set_positions, "cipc_set_positions");
        cipc_shim::upload_rest(s, nodeAttr);
        int n = 0;
        cipc_shim::die(s.ctx, cipc_constraint_set(s.ctx, 0, dHat2, thickness, &n), "cipc_constraint_set");
        // the reference resize(0)s and refills both containers (IPC.h:587-590); here every element is written by the delivery
        static_assert(sizeof(VECTOR<T, 2>) == 32, "VECTOR<T,dim> stores T data[4]");
        cipc_shim::resize_uninitialized(constraintSet, (size_t)n);
        cipc_shim::resize_uninitialized(stencilInfo, (size_t)n);
        cipc_shim::die(s.ctx, cipc_get_constraints_strided(s.ctx, n ? constraintSet[0].data : nullptr, n ? stencilInfo[0].data : nullptr, 32),
            "cipc_get_constraints");
        cipc_shim::remember_constraints(s, constraintSet);
        cipc_shim::report_scopes(s, "Compute_Constraint_Set", "ccs_hash_build", "ccs_pairs", "ccs_narrow_pt", "ccs_narrow_ee", "ccs_merge",
            (double)(cipc_counter(s.ctx, "candidates_pt") + cipc_counter(s.ctx, "candidates_pe") + cipc_counter(s.ctx, "candidates_pp")),
            (double)cipc_counter(s.ctx, "candidates_ee"));
    }
}

// ------------------------------------------------------------------ FEM/IPC.h:742-748
template <class T, int dim, bool elasticIPC = false>
void Compute_Barrier(MESH_NODE<T, dim>& X, MESH_NODE_ATTR<T, dim>& nodeAttr, const std::vector<VECTOR<int, dim + 1>>& constraintSet,
    const std::vector<VECTOR<T, 2>>& stencilInfo, T dHat2, T kappa[], T thickness, T& E)
{
    if constexpr (!cipc_shim::on_gpu<T, dim, false, elasticIPC>) {
        Compute_Barrier_CPU<T, dim, elasticIPC>(X, nodeAttr, constraintSet, stencilInfo, dHat2, kappa, thickness, E);
    }
    else {
        TIMER_FLAG("Compute_Barrier");
        cipc_shim::State& s = cipc_shim::state();
        cipc_shim::upload_positions(s, X, cipc_set_positions, "cipc_set_positions");
        cipc_shim::ensure_constraints(s, constraintSet, stencilInfo);
        cipc_shim::die(s.ctx, cipc_barrier_energy(s.ctx, 0, dHat2, kappa, thickness, &E), "cipc_barrier_energy");
    }
}

// ------------------------------------------------------------------ FEM/IPC.h:943-948
template <class T, int dim, bool elasticIPC = false>
void Compute_Barrier_Gradient(MESH_NODE<T, dim>& X, const std::vector<VECTOR<int, dim + 1>>& constraintSet,
    const std::vector<VECTOR<T, 2>>& stencilInfo, T dHat2, T kappa[], T thickness, MESH_NODE_ATTR<T, dim>& nodeAttr)
{
    if constexpr (!cipc_shim::on_gpu<T, dim, false, elasticIPC>) {
        Compute_Barrier_Gradient_CPU<T, dim, elasticIPC>(X, constraintSet, stencilInfo, dHat2, kappa, thickness, nodeAttr);
    }
    else {
        TIMER_FLAG("Compute_Barrier_Gradient");
        cipc_shim::State& s = cipc_shim::state();
        cipc_shim::upload_positions(s, X, cipc_set_positions, "cipc_set_positions");
        cipc_shim::ensure_constraints(s, constraintSet, stencilInfo);
        const size_t n = X.size;
        s.stage3.resize(3 * n);
        double* z = s.stage3.data();
        cipc_shim::parallel_nodes(n, [z](size_t i) { z[3 * i] = 0.0; z[3 * i + 1] = 0.0; z[3 * i + 2] = 0.0; }); // the C ABI accumulates
        cipc_shim::die(s.ctx, cipc_barrier_gradient(s.ctx, 0, dHat2, kappa, thickness, s.stage3.data(), 24), "cipc_barrier_gradient");
        const double* st = s.stage3.data();
        cipc_shim::parallel_nodes(n, [&nodeAttr, st](size_t i) { // nodeAttr.g += (IPC.h:1034-1042); g lives inside the AoSoA record of node i
            VECTOR<T, dim>& g = std::get<FIELDS<MESH_NODE_ATTR<T, dim>>::g>(nodeAttr.Get_Unchecked(i));
            g[0] += st[3 * i]; g[1] += st[3 * i + 1]; g[2] += st[3 * i + 2];
        });
    }
}

// ------------------------------------------------------------------ FEM/IPC.h:1258-1265
template <class T, int dim, bool elasticIPC = false>
void Compute_Barrier_Hessian(MESH_NODE<T, dim>& X, MESH_NODE_ATTR<T, dim>& nodeAttr, const std::vector<VECTOR<int, dim + 1>>& constraintSet,
    const std::vector<VECTOR<T, 2>>& stencilInfo, T dHat2, T kappa[], T thickness, bool projectSPD, std::vector<Eigen::Triplet<T>>& triplets)
{
    if constexpr (!cipc_shim::on_gpu<T, dim, false, elasticIPC>) {
        Compute_Barrier_Hessian_CPU<T, dim, elasticIPC>(X, nodeAttr, constraintSet, stencilInfo, dHat2, kappa, thickness, projectSPD, triplets);
    }
    else {
        TIMER_FLAG("Compute_Barrier_Hessian");
        static_assert(sizeof(Eigen::Triplet<T>) == sizeof(cipc_triplet), "Eigen::Triplet<double> is {int,int,double}");
        cipc_shim::State& s = cipc_shim::state();
        cipc_shim::upload_positions(s, X, cipc_set_positions, "cipc_set_positions");
        cipc_shim::ensure_constraints(s, constraintSet, stencilInfo);
        int64_t n = 0;
        // the vector is new in every Newton iteration (INC_POTENTIAL.h:321): reserve it from the previous iteration's count and
        // let its first-touch page faults overlap the device work
        static size_t lastCount = 0;
        if (lastCount) {
            const size_t guess = lastCount + lastCount / 32;
            triplets.reserve(triplets.size() + guess);
            cipc_host_prefault_async(triplets.data() + triplets.size(), guess * sizeof(Eigen::Triplet<T>));
        }
        if (s.merged) cipc_shim::die(s.ctx, cipc_barrier_hessian_merged(s.ctx, 0, dHat2, kappa, thickness, projectSPD ? 1 : 0, &n), "cipc_barrier_hessian_merged");
        else cipc_shim::die(s.ctx, cipc_barrier_hessian(s.ctx, 0, dHat2, kappa, thickness, projectSPD ? 1 : 0, &n), "cipc_barrier_hessian");
        // the new entries are APPENDED (IPC.h:1371,1388), without the zero-fill of resize()
        if (n) cipc_shim::die(s.ctx, cipc_get_triplets(s.ctx, reinterpret_cast<cipc_triplet*>(cipc_shim::grow_uninitialized(triplets, (size_t)n))), "cipc_get_triplets");
        lastCount = (size_t)n;
    }
}

// ------------------------------------------------------------------ optional CSR hand-off
// The contact part of the system matrix assembled ON THE DEVICE (duplicates summed, columns sorted) and delivered in the three
// arrays the reference's CSR_MATRIX<T>::Construct_From_CSR takes (Math/CSR_MATRIX.h:33-47), instead of triplets that
// Construct_From_Triplet / setFromTriplets would have to sort (Math/CSR_MATRIX.h:49-56, Shell/INC_POTENTIAL.h:373-394).  A caller
// that wants it builds with -DCIPC_SHIM_CSR and replaces the Compute_Barrier_Hessian call next to INC_POTENTIAL.h:382 (see
// INTEGRATION.md 4d):  ptr has 3 X.size + 1 entries.
template <class T, int dim, bool elasticIPC = false>
void Compute_Barrier_Hessian_CSR(MESH_NODE<T, dim>& X, MESH_NODE_ATTR<T, dim>& nodeAttr, const std::vector<VECTOR<int, dim + 1>>& constraintSet,
    const std::vector<VECTOR<T, 2>>& stencilInfo, T dHat2, T kappa[], T thickness, bool projectSPD, std::vector<int>& ptr, std::vector<int>& col,
    std::vector<T>& val)
{
    static_assert(cipc_shim::on_gpu<T, dim, false, elasticIPC>, "the CSR hand-off exists for the instantiation the CUDA path serves (double, dim = 3, OIPC)");
    static_assert(sizeof(int) == sizeof(int32_t), "CSR index type");
    (void)nodeAttr;
    TIMER_FLAG("Compute_Barrier_Hessian_CSR");
    cipc_shim::State& s = cipc_shim::state();
    cipc_shim::upload_positions(s, X, cipc_set_positions, "cipc_set_positions");
    cipc_shim::ensure_constraints(s, constraintSet, stencilInfo);
    int64_t n = 0, nnz = 0;
    cipc_shim::die(s.ctx, cipc_csr_begin(s.ctx), "cipc_csr_begin");
    cipc_shim::die(s.ctx, cipc_barrier_hessian_dev(s.ctx, 0, dHat2, kappa, thickness, projectSPD ? 1 : 0, &n), "cipc_barrier_hessian_dev");
    cipc_shim::die(s.ctx, cipc_csr_add(s.ctx), "cipc_csr_add");
    cipc_shim::die(s.ctx, cipc_csr_finish(s.ctx, &nnz), "cipc_csr_finish");
    cipc_shim::resize_uninitialized(ptr, (size_t)dim * X.size + 1);
    cipc_shim::resize_uninitialized(col, (size_t)nnz);
    cipc_shim::resize_uninitialized(val, (size_t)nnz);
    cipc_shim::die(s.ctx, cipc_get_csr(s.ctx, ptr.data(), col.data(), val.data()), "cipc_get_csr");
}

// ------------------------------------------------------------------ FEM/IPC.h:1879-1890
template <class T, int dim, bool shell = false, bool elasticIPC = false>
void Compute_Intersection_Free_StepSize(MESH_NODE<T, dim>& X, const std::vector<int>& boundaryNode,
    const std::vector<VECTOR<int, 2>>& boundaryEdge, const std::vector<VECTOR<int, 3>>& boundaryTri, const std::vector<int>& particle,
    const std::vector<VECTOR<int, 2>>& rod, const std::map<int, std::set<int>>& NNExclusion, const VECTOR<int, 2>& codimBNStartInd,
    const std::vector<bool>& DBCb, const std::vector<T>& searchDir, T thickness, T& stepSize)
{
    if constexpr (!cipc_shim::on_gpu<T, dim, shell, elasticIPC>) {
        Compute_Intersection_Free_StepSize_CPU<T, dim, shell, elasticIPC>(X, boundaryNode, boundaryEdge, boundaryTri, particle, rod, NNExclusion,
            codimBNStartInd, DBCb, searchDir, thickness, stepSize);
    }
    else {
        TIMER_FLAG("Compute_Intersection_Free_StepSize");
        cipc_shim::State& s = cipc_shim::state();
        cipc_shim::upload_topology(s, X.size, boundaryNode, boundaryEdge, boundaryTri, rod.size(), NNExclusion, codimBNStartInd, DBCb);
        cipc_shim::upload_positions(s, X, cipc_set_positions, "cipc_set_positions");
        cipc_shim::die(s.ctx, cipc_set_search_dir(s.ctx, searchDir.data()), "cipc_set_search_dir");
        cipc_shim::die(s.ctx, cipc_step_size(s.ctx, 0, thickness, &stepSize), "cipc_step_size");
        const double nAll = (double)cipc_counter(s.ctx, "ccd_pairs"), nEE = (double)cipc_counter(s.ctx, "ccd_pairs_ee");
        cipc_shim::report_scopes(s, "Compute_Intersection_Free_StepSize", "ccd_hash_build", "ccd_pairs", "ccd_accd_pt", "ccd_accd_ee", nullptr, nAll - nEE, nEE);
    }
}

// ------------------------------------------------------------------ FEM/IPC.h:2246-2249
template <class T, int dim, bool elasticIPC = false>
void Compute_Min_Dist2(MESH_NODE<T, dim>& X, const std::vector<VECTOR<int, dim + 1>>& constraintSet, T thickness, std::vector<T>& dist2,
    T& minDist2)
{
    if constexpr (!cipc_shim::on_gpu<T, dim, false, elasticIPC>) {
        Compute_Min_Dist2_CPU<T, dim, elasticIPC>(X, constraintSet, thickness, dist2, minDist2);
    }
    else {
        TIMER_FLAG("Compute_Min_Dist");
        if (constraintSet.empty()) return; // IPC.h:2253: dist2 and minDist2 stay untouched
        cipc_shim::State& s = cipc_shim::state();
        cipc_shim::upload_positions(s, X, cipc_set_positions, "cipc_set_positions");
        if (!cipc_shim::constraints_resident(s, constraintSet)) {
            s.stage2.assign(2 * constraintSet.size(), 1.0); // stencilInfo is not an argument here and not read by the distance pass
            cipc_shim::die(s.ctx, cipc_set_constraints(s.ctx, constraintSet[0].data, s.stage2.data(), (int)constraintSet.size()), "cipc_set_constraints");
            s.csPtr = nullptr; // the resident weights are placeholders: the next barrier call uploads its own
        }
        cipc_shim::resize_uninitialized(dist2, constraintSet.size());
        cipc_shim::die(s.ctx, cipc_min_dist2(s.ctx, thickness, dist2.data(), &minDist2), "cipc_min_dist2");
    }
}

} // namespace JGSL
