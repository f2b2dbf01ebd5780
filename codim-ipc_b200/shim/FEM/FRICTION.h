// shim/FEM/FRICTION.h -- drop-in replacement of the reference's Library/FEM/FRICTION.h lagged-friction entry
// points, same mechanism as shim/FEM/IPC.h: with `-I codim-ipc_b200/shim -I include` before
// `-I Library`, every `#include <FEM/FRICTION.h>` of the reference (FEM/TimeStepper/IMPLICIT_EULER.h:7, ADMM.h:7,
// SHAPE_UP.h:7) lands here.  The reference's own header is included underneath with its five templates renamed to
// *_CPU; five templates with the reference's exact names and signatures are defined on top:
//
//   <double, dim=3 (, elasticIPC=false)>  -> CUDA path through the C ABI (include/cipc_b200.h); no CPU fallback
//   anything else                          -> the reference's own CPU template, untouched (outside the graft)
//
// Reference signatures: FEM/FRICTION.h:16-25, 126-130, 172-180, 254-262, 381-390.
// The friction set (constraintSet, closestPoint, tanBasis, normalForce) stays resident on the device between the
// calls of one Newton solve; the containers the caller passes are checked by address, size and a 64-bit hash of their
// full contents (cipc_hash_bytes) and uploaded again only when they are not the ones this shim filled last.
#pragma once

#include <FEM/IPC.h> // the contact shim: cipc_shim::State, upload helpers, the C ABI

#define Compute_Friction_Basis Compute_Friction_Basis_CPU
#define Compute_Friction_Coef Compute_Friction_Coef_CPU
#define Compute_Friction_Potential Compute_Friction_Potential_CPU
#define Compute_Friction_Gradient Compute_Friction_Gradient_CPU
#define Compute_Friction_Hessian Compute_Friction_Hessian_CPU
#include_next <FEM/FRICTION.h>
#undef Compute_Friction_Basis
#undef Compute_Friction_Coef
#undef Compute_Friction_Potential
#undef Compute_Friction_Gradient
#undef Compute_Friction_Hessian

namespace JGSL {
namespace cipc_shim {

struct FrictionState {
    const void* csPtr = nullptr;
    const void* nfPtr = nullptr;
    size_t n = 0;
    bool full = false; // closestPoint / tanBasis are resident too
    uint64_t csHash = 0, nfHash = 0;
};
inline FrictionState& fstate()
{
    static FrictionState f;
    return f;
}
inline void remember_friction(const std::vector<VECTOR<int, 4>>& cs, const std::vector<double>& nf, bool full)
{
    FrictionState& f = fstate();
    f.csPtr = cs.data(); f.nfPtr = nf.data(); f.n = cs.size(); f.full = full;
    f.csHash = cs.empty() ? 0 : cipc_hash_bytes(cs.data(), cs.size() * sizeof(VECTOR<int, 4>));
    f.nfHash = nf.empty() ? 0 : cipc_hash_bytes(nf.data(), nf.size() * sizeof(double));
}
inline bool friction_resident(const std::vector<VECTOR<int, 4>>& cs, const std::vector<double>& nf, bool needFull)
{
    const FrictionState& f = fstate();
    if (!(cs.data() == f.csPtr && nf.data() == f.nfPtr && cs.size() == f.n && nf.size() == f.n && (f.full || !needFull))) return false;
    if (cs.empty()) return true;
    return cipc_hash_bytes(cs.data(), cs.size() * sizeof(VECTOR<int, 4>)) == f.csHash && cipc_hash_bytes(nf.data(), nf.size() * sizeof(double)) == f.nfHash;
}
template <class CP, class TB>
inline void ensure_friction(State& s, const std::vector<VECTOR<int, 4>>& cs, const std::vector<CP>& closestPoint, const std::vector<TB>& tanBasis,
    const std::vector<double>& nf)
{
    static_assert(sizeof(CP) == 16 && sizeof(TB) == 48, "Eigen::Matrix<double,2,1> / <double,3,2> are packed fixed-size records");
    if (friction_resident(cs, nf, true)) return;
    die(s.ctx, cipc_set_friction_basis(s.ctx, cs.empty() ? nullptr : cs[0].data, closestPoint.empty() ? nullptr : closestPoint[0].data(),
                   tanBasis.empty() ? nullptr : tanBasis[0].data(), nf.data(), (int)cs.size()), "cipc_set_friction_basis");
    remember_friction(cs, nf, true);
}

template <class T, int dim, bool elasticIPC = false>
constexpr bool friction_on_gpu = std::is_same<T, double>::value && dim == 3 && !elasticIPC;

} // namespace cipc_shim

// ------------------------------------------------------------------ FEM/FRICTION.h:16-25
template <class T, int dim = 3, bool elasticIPC = false>
void Compute_Friction_Basis(MESH_NODE<T, dim>& X, const std::vector<VECTOR<int, dim + 1>>& contactConstraintSet,
    const std::vector<VECTOR<T, 2>>& stencilInfo, std::vector<VECTOR<int, dim + 1>>& constraintSet,
    std::vector<Eigen::Matrix<T, dim - 1, 1>>& closestPoint, std::vector<Eigen::Matrix<T, dim, dim - 1>>& tanBasis,
    std::vector<T>& normalForce, T dHat2, T kappa[], T thickness)
{
    if constexpr (!cipc_shim::friction_on_gpu<T, dim, elasticIPC>) {
        Compute_Friction_Basis_CPU<T, dim, elasticIPC>(X, contactConstraintSet, stencilInfo, constraintSet, closestPoint, tanBasis, normalForce,
            dHat2, kappa, thickness);
    }
    else {
        TIMER_FLAG("Compute_Friction_Basis");
        static_assert(sizeof(Eigen::Matrix<T, 2, 1>) == 16 && sizeof(Eigen::Matrix<T, 3, 2>) == 48, "packed fixed-size Eigen records");
        cipc_shim::State& s = cipc_shim::state();
        cipc_shim::upload_positions(s, X, cipc_set_positions, "cipc_set_positions");
        cipc_shim::ensure_constraints(s, contactConstraintSet, stencilInfo);
        int n = 0;
        cipc_shim::die(s.ctx, cipc_friction_basis(s.ctx, 0, dHat2, kappa, thickness, &n), "cipc_friction_basis");
        constraintSet.resize(n); closestPoint.resize(n); tanBasis.resize(n); normalForce.resize(n); // FRICTION.h:35-46
        cipc_shim::die(s.ctx, cipc_get_friction_basis(s.ctx, n ? constraintSet[0].data : nullptr, n ? closestPoint[0].data() : nullptr,
                                  n ? tanBasis[0].data() : nullptr, normalForce.data()), "cipc_get_friction_basis");
        cipc_shim::remember_friction(constraintSet, normalForce, true);
    }
}

// ------------------------------------------------------------------ FEM/FRICTION.h:126-130
template <class T, int dim = 3>
void Compute_Friction_Coef(const std::vector<VECTOR<int, dim + 1>>& constraintSet, const std::vector<int>& compNodeRange,
    const std::vector<T>& muComp, std::vector<T>& normalForce, T& mu)
{
    if constexpr (!cipc_shim::friction_on_gpu<T, dim>) {
        Compute_Friction_Coef_CPU<T, dim>(constraintSet, compNodeRange, muComp, normalForce, mu);
    }
    else {
        cipc_shim::State& s = cipc_shim::state();
        bool full = true;
        if (!cipc_shim::friction_resident(constraintSet, normalForce, false)) {
            // only the stencils and the forces are needed here; closest points / bases follow with the next call
            cipc_shim::die(s.ctx, cipc_set_friction_basis(s.ctx, constraintSet.empty() ? nullptr : constraintSet[0].data, nullptr, nullptr,
                                      normalForce.data(), (int)constraintSet.size()), "cipc_set_friction_basis");
            full = false;
        }
        else full = cipc_shim::fstate().full;
        cipc_shim::die(s.ctx, cipc_friction_coef(s.ctx, (int)compNodeRange.size(), compNodeRange.data(), muComp.data(), &mu), "cipc_friction_coef");
        cipc_shim::die(s.ctx, cipc_get_friction_basis(s.ctx, nullptr, nullptr, nullptr, normalForce.data()), "cipc_get_friction_basis");
        cipc_shim::remember_friction(constraintSet, normalForce, full);
    }
}

// ------------------------------------------------------------------ FEM/FRICTION.h:172-180
template <class T, int dim = 3>
void Compute_Friction_Potential(MESH_NODE<T, dim>& X, MESH_NODE<T, dim>& Xn, const std::vector<VECTOR<int, dim + 1>>& constraintSet,
    const std::vector<Eigen::Matrix<T, dim - 1, 1>>& closestPoint, const std::vector<Eigen::Matrix<T, dim, dim - 1>>& tanBasis,
    const std::vector<T>& normalForce, T epsvh2, T mu, T& E)
{
    if constexpr (!cipc_shim::friction_on_gpu<T, dim>) {
        Compute_Friction_Potential_CPU<T, dim>(X, Xn, constraintSet, closestPoint, tanBasis, normalForce, epsvh2, mu, E);
    }
    else {
        TIMER_FLAG("Compute_Friction_Potential");
        cipc_shim::State& s = cipc_shim::state();
        cipc_shim::upload_positions(s, X, cipc_set_positions, "cipc_set_positions");
        cipc_shim::upload_positions(s, Xn, cipc_set_prev_positions, "cipc_set_prev_positions");
        cipc_shim::ensure_friction(s, constraintSet, closestPoint, tanBasis, normalForce);
        cipc_shim::die(s.ctx, cipc_friction_energy(s.ctx, epsvh2, mu, &E), "cipc_friction_energy");
    }
}

// ------------------------------------------------------------------ FEM/FRICTION.h:254-262
template <class T, int dim = 3>
void Compute_Friction_Gradient(MESH_NODE<T, dim>& X, MESH_NODE<T, dim>& Xn, const std::vector<VECTOR<int, dim + 1>>& constraintSet,
    const std::vector<Eigen::Matrix<T, dim - 1, 1>>& closestPoint, const std::vector<Eigen::Matrix<T, dim, dim - 1>>& tanBasis,
    const std::vector<T>& normalForce, T epsvh2, T mu, MESH_NODE_ATTR<T, dim>& nodeAttr)
{
    if constexpr (!cipc_shim::friction_on_gpu<T, dim>) {
        Compute_Friction_Gradient_CPU<T, dim>(X, Xn, constraintSet, closestPoint, tanBasis, normalForce, epsvh2, mu, nodeAttr);
    }
    else {
        TIMER_FLAG("Compute_Friction_Gradient");
        cipc_shim::State& s = cipc_shim::state();
        cipc_shim::upload_positions(s, X, cipc_set_positions, "cipc_set_positions");
        cipc_shim::upload_positions(s, Xn, cipc_set_prev_positions, "cipc_set_prev_positions");
        cipc_shim::ensure_friction(s, constraintSet, closestPoint, tanBasis, normalForce);
        const size_t n = X.size;
        s.stage3.resize(3 * n);
        double* z = s.stage3.data();
        cipc_shim::parallel_nodes(n, [z](size_t i) { z[3 * i] = 0.0; z[3 * i + 1] = 0.0; z[3 * i + 2] = 0.0; }); // the C ABI accumulates
        cipc_shim::die(s.ctx, cipc_friction_gradient(s.ctx, epsvh2, mu, s.stage3.data(), 24), "cipc_friction_gradient");
        const double* st = s.stage3.data();
        cipc_shim::parallel_nodes(n, [&nodeAttr, st](size_t i) { // nodeAttr.g += (FRICTION.h:294-297)
            VECTOR<T, dim>& g = std::get<FIELDS<MESH_NODE_ATTR<T, dim>>::g>(nodeAttr.Get_Unchecked(i));
            g[0] += st[3 * i]; g[1] += st[3 * i + 1]; g[2] += st[3 * i + 2];
        });
    }
}

// ------------------------------------------------------------------ FEM/FRICTION.h:381-390
template <class T, int dim = 3>
void Compute_Friction_Hessian(MESH_NODE<T, dim>& X, MESH_NODE<T, dim>& Xn, const std::vector<VECTOR<int, dim + 1>>& constraintSet,
    const std::vector<Eigen::Matrix<T, dim - 1, 1>>& closestPoint, const std::vector<Eigen::Matrix<T, dim, dim - 1>>& tanBasis,
    const std::vector<T>& normalForce, T epsvh2, T mu, bool projectSPD, std::vector<Eigen::Triplet<T>>& triplets)
{
    if constexpr (!cipc_shim::friction_on_gpu<T, dim>) {
        Compute_Friction_Hessian_CPU<T, dim>(X, Xn, constraintSet, closestPoint, tanBasis, normalForce, epsvh2, mu, projectSPD, triplets);
    }
    else {
        TIMER_FLAG("Compute_Friction_Hessian");
        static_assert(sizeof(Eigen::Triplet<T>) == sizeof(cipc_triplet), "Eigen::Triplet<double> is {int,int,double}");
        cipc_shim::State& s = cipc_shim::state();
        cipc_shim::upload_positions(s, X, cipc_set_positions, "cipc_set_positions");
        cipc_shim::upload_positions(s, Xn, cipc_set_prev_positions, "cipc_set_prev_positions");
        cipc_shim::ensure_friction(s, constraintSet, closestPoint, tanBasis, normalForce);
        int64_t n = 0;
        if (s.merged) cipc_shim::die(s.ctx, cipc_friction_hessian_merged(s.ctx, epsvh2, mu, projectSPD ? 1 : 0, &n), "cipc_friction_hessian_merged");
        else cipc_shim::die(s.ctx, cipc_friction_hessian(s.ctx, epsvh2, mu, projectSPD ? 1 : 0, &n), "cipc_friction_hessian");
        // the new entries are APPENDED (FRICTION.h:404-421), without the zero-fill of resize()
        if (n) cipc_shim::die(s.ctx, cipc_get_triplets(s.ctx, reinterpret_cast<cipc_triplet*>(cipc_shim::grow_uninitialized(triplets, (size_t)n))), "cipc_get_triplets");
    }
}

} // namespace JGSL
