// friction.cuh -- lagged friction on the non-mollified part of the contact constraint set
// (reference: FEM/FRICTION.h:16-663, FEM/FRICTION_UTILS.h).  Included by cipc_b200.cu
// after the stencil decoding helpers (ldx / ldd / YHdr / gadd / BarrierParams live there).
//
// Per friction stencil the reference keeps (constraint, closestPoint[2], tanBasis 3x2, normalForce).  With
//   u    = B^T * sum_k coef_k (x_k - xn_k)            (relative sliding in the tangent plane, 2-vector)
//   TT   = coef (x) B^T                                (2 x 3nb)
// the potential is  mu * lam * f0(|u|),  the gradient  TT^T (f1(|u|)/|u| mu lam u)  and the Hessian
// TT^T M TT  with the 2x2 matrix M of FRICTION.h:440-465.  M is positive semi-definite in closed form
// (eigenpairs below), so the block is  s * sum_{k<2} y_k y_k^T  with  y_k = sqrt(|s| l_k) coef (x) (B q_k)
// and goes through the same factor -> triplet expansion as the barrier Hessian (two vectors per stencil).
#pragma once

namespace cipc {

__device__ __forceinline__ dv3 fr_normalized(const dv3& a) // Eigen normalized(): v / sqrt(|v|^2) when the norm is positive
{
    const double z = norm2(a);
    return z > 0.0 ? a / sqrt(z) : a;
}
// FRICTION_UTILS.h:10-39
__device__ __forceinline__ double fr_f0(double x2, double epsvh)
{
    if (x2 >= epsvh * epsvh) return sqrt(x2);
    return x2 * (-sqrt(x2) / 3.0 + epsvh) / (epsvh * epsvh) + epsvh / 3.0;
}
__device__ __forceinline__ double fr_f1_div(double x2, double epsvh)
{
    if (x2 >= epsvh * epsvh) return 1.0 / sqrt(x2);
    return (-sqrt(x2) + 2.0 * epsvh) / (epsvh * epsvh);
}

// a stencil is kept for friction unless it is a mollified one (FRICTION.h:37-41)
__device__ __forceinline__ bool fr_keep(const int4 c) { return !(c.x >= 0 && (c.z < 0 || c.w < 0)); }

struct FrStencil {
    int nb, mult;
    int v[4];
    double coef[4]; // u3 = sum_k coef[k] dx_k  (rows of TT: coef[k] * B^T)
};
// FRICTION_UTILS.h: *_RelDX / *_TT
__device__ __forceinline__ FrStencil fr_decode(const int4 c, const double2 cp)
{
    FrStencil s;
    s.mult = 1;
    s.v[1] = c.y; s.v[2] = c.z; s.v[3] = c.w;
    s.coef[2] = s.coef[3] = 0.0;
    if (c.x >= 0) { // EE
        s.nb = 4; s.v[0] = c.x;
        s.coef[0] = 1.0 - cp.x; s.coef[1] = cp.x; s.coef[2] = cp.y - 1.0; s.coef[3] = -cp.y;
        return s;
    }
    s.v[0] = -c.x - 1;
    s.coef[0] = 1.0;
    if (c.z < 0) { s.nb = 2; s.coef[1] = -1.0; s.mult = -c.w; }                                          // PP
    else if (c.w < 0) { s.nb = 3; s.coef[1] = cp.x - 1.0; s.coef[2] = -cp.x; s.mult = -c.w; }          // PE
    else { s.nb = 4; s.coef[1] = -1.0 + cp.x + cp.y; s.coef[2] = -cp.x; s.coef[3] = -cp.y; }           // PT
    return s;
}
// relative displacement in 3-D, evaluated in the reference's association order (FRICTION_UTILS.h:68-78,144-154,206-215,247-254)
__device__ __forceinline__ dv3 fr_rel3(const double4* __restrict__ X, const double4* __restrict__ Xn, const int4 c, const FrStencil& s,
    const double2 cp)
{
    dv3 d[4];
    for (int k = 0; k < s.nb; ++k) d[k] = ldd(X, s.v[k]) - ldd(Xn, s.v[k]);
    if (c.x >= 0) return (d[0] + cp.x * (d[1] - d[0])) - (d[2] + cp.y * (d[3] - d[2]));
    if (c.z < 0) return d[0] - d[1];
    if (c.w < 0) return d[0] - (d[1] + cp.x * (d[2] - d[1]));
    return d[0] - ((d[1] + cp.x * (d[2] - d[1])) + cp.y * (d[3] - d[1]));
}
struct FrBasis {
    dv3 b0, b1;
};
__device__ __forceinline__ FrBasis fr_load_basis(const double* __restrict__ B, u32 i)
{
    const double2* q = reinterpret_cast<const double2*>(B + (size_t)i * 6); // 48-byte records, 16-byte aligned
    const double2 a = q[0], b = q[1], c = q[2];
    FrBasis r;
    r.b0 = dv3(a.x, a.y, b.x);
    r.b1 = dv3(b.y, c.x, c.y);
    return r;
}

// ---- Compute_Friction_Basis (FRICTION.h:16-124): order-preserving compaction + closest point, tangent basis, lagged normal force
__global__ void k_friction_flags(const int4* __restrict__ cs, u32 n, u32* keep)
{
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) keep[i] = fr_keep(cs[i]) ? 1u : 0u;
}
__global__ void __launch_bounds__(128) k_friction_basis(const double4* __restrict__ X, const int4* __restrict__ cs, const double2* __restrict__ info,
    const u32* __restrict__ slot, u32 n, BarrierParams bp, int4* __restrict__ fcs, double2* __restrict__ fcp, double* __restrict__ fB,
    double* __restrict__ fnf)
{
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int4 c = cs[i];
    if (!fr_keep(c)) return;
    const u32 q = slot[i];
    double2 cp = make_double2(0.0, 0.0);
    dv3 b0, b1;
    xd dist2;
    if (c.x >= 0) {
        const xv3 v0 = ldx(X, c.x), v1 = ldx(X, c.y), v2 = ldx(X, c.z), v3 = ldx(X, c.w);
        const xv3 e20 = v0 - v2, e01 = v1 - v0, e23 = v3 - v2;
        xd g1, g2;
        ldlt2_solve(norm2(e01), -dot(e23, e01), norm2(e23), -dot(e20, e01), dot(e20, e23), g1, g2); // FRICTION_UTILS.h:120-142
        cp = make_double2(g1.v, g2.v);
        const dv3 d01 = to_d(e01);
        b0 = fr_normalized(d01);
        b1 = fr_normalized(cross(cross(d01, to_d(e23)), d01)); // :107-118
        dist2 = ee_dist2(v0, v1, v2, v3);
    }
    else {
        const xv3 p = ldx(X, -c.x - 1), v1 = ldx(X, c.y);
        if (c.z < 0) {
            const dv3 v01 = to_d(v1 - p); // :229-245
            const dv3 xC = cross(dv3(1.0, 0.0, 0.0), v01), yC = cross(dv3(0.0, 1.0, 0.0), v01);
            if (norm2(xC) > norm2(yC)) { b0 = fr_normalized(xC); b1 = fr_normalized(cross(v01, xC)); }
            else { b0 = fr_normalized(yC); b1 = fr_normalized(cross(v01, yC)); }
            dist2 = pp_dist2(p, v1);
        }
        else if (c.w < 0) {
            const xv3 v2 = ldx(X, c.z);
            const xv3 e12 = v2 - v1;
            cp.x = (dot(p - v1, e12) / norm2(e12)).v; // :196-204
            const dv3 d12 = to_d(e12);
            b0 = fr_normalized(d12);
            b1 = fr_normalized(cross(d12, to_d(p - v1))); // :184-194
            dist2 = pe_dist2(p, v1, v2);
        }
        else {
            const xv3 v2 = ldx(X, c.z), v3 = ldx(X, c.w);
            const xv3 r0 = v2 - v1, r1 = v3 - v1, rel = p - v1;
            xd s1, s2;
            ldlt2_solve(dot(r0, r0), dot(r1, r0), dot(r1, r1), dot(r0, rel), dot(r1, rel), s1, s2); // :54-66
            cp = make_double2(s1.v, s2.v);
            const dv3 d12 = to_d(r0);
            b0 = fr_normalized(d12);
            b1 = fr_normalized(cross(cross(d12, to_d(r1)), d12)); // :41-52
            dist2 = pt_dist2(p, v1, v2, v3);
        }
    }
    const double bG = barrier_g(bp.elastic, dist2.v - bp.thickness2, bp.dHat2, bp.k0);
    fcs[q] = c;
    fcp[q] = cp;
    double2* o = reinterpret_cast<double2*>(fB + (size_t)q * 6);
    o[0] = make_double2(b0.x, b0.y); o[1] = make_double2(b0.z, b1.x); o[2] = make_double2(b1.y, b1.z);
    // the reference indexes stencilInfo with the FILTERED index (FRICTION.h:112); weights are 1 when !elasticIPC
    fnf[q] = -bG * 2.0 * sqrt(dist2.v) * info[q].x;
}
// ---- Compute_Friction_Coef (FRICTION.h:126-170): normalForce *= muComp[comp(v0) + comp(v1) * nComp]
__global__ void k_friction_coef(const int4* __restrict__ fcs, u32 n, const int* __restrict__ compNodeRange, int nComp,
    const double* __restrict__ muComp, double* __restrict__ fnf, int* errFlag)
{
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int4 c = fcs[i];
    const int va = c.x >= 0 ? c.x : -c.x - 1, vb = c.x >= 0 ? c.z : c.y;
    int ca = -1, cb = -1;
    for (int k = nComp - 1; k >= 0; --k) {
        if (va < compNodeRange[k]) ca = k;
        if (vb < compNodeRange[k]) cb = k;
    }
    if (ca < 0 || cb < 0) { *errFlag = CIPC_ERR_ARG; return; } // reference: "can't find node compI" + exit(-1)
    fnf[i] *= muComp[ca + cb * nComp];
}

// ---- Compute_Friction_Potential (FRICTION.h:172-252)
__global__ void __launch_bounds__(RED_BT) k_friction_energy(const double4* __restrict__ X, const double4* __restrict__ Xn,
    const int4* __restrict__ fcs, const double2* __restrict__ fcp, const double* __restrict__ fB, const double* __restrict__ fnf, u32 n,
    double epsvh, double* partial)
{
    double acc = 0;
    for (u32 i = blockIdx.x * RED_BT + threadIdx.x; i < n; i += RED_GRID * RED_BT) {
        const int4 c = fcs[i];
        const double2 cp = fcp[i];
        const FrStencil s = fr_decode(c, cp);
        const dv3 r = fr_rel3(X, Xn, c, s, cp);
        const FrBasis B = fr_load_basis(fB, i);
        const double u0 = dot(r, B.b0), u1 = dot(r, B.b1);
        double e = fr_f0(u0 * u0 + u1 * u1, epsvh) * fnf[i];
        if (c.w < -1) e *= (double)(-c.w);
        acc += e;
    }
    acc = block_sum<RED_BT>(acc);
    if (threadIdx.x == 0) partial[blockIdx.x] = acc;
}
// ---- Compute_Friction_Gradient (FRICTION.h:254-379)
__global__ void __launch_bounds__(128) k_friction_gradient(const double4* __restrict__ X, const double4* __restrict__ Xn,
    const int4* __restrict__ fcs, const double2* __restrict__ fcp, const double* __restrict__ fB, const double* __restrict__ fnf, u32 n,
    double epsvh, double mu, double* g)
{
    const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int4 c = fcs[i];
    const double2 cp = fcp[i];
    const FrStencil s = fr_decode(c, cp);
    const dv3 r = fr_rel3(X, Xn, c, s, cp);
    const FrBasis B = fr_load_basis(fB, i);
    const double u0 = dot(r, B.b0), u1 = dot(r, B.b1);
    const double sc = fr_f1_div(u0 * u0 + u1 * u1, epsvh) * (double)s.mult * mu * fnf[i];
    const double t3[3] = {sc * (B.b0.x * u0 + B.b1.x * u1), sc * (B.b0.y * u0 + B.b1.y * u1), sc * (B.b0.z * u0 + B.b1.z * u1)};
    for (int k = 0; k < s.nb; ++k) gadd(g, s.v[k], t3, s.coef[k]);
}
// ---- Compute_Friction_Hessian (FRICTION.h:381-663), phase A: two factor vectors per stencil.
// M = s * (l_a qa qa^T + l_b qb qb^T), s = mult mu lam:
//   |u| >= eps_v h : l_a = f1/|u|^2 along ubar = (-u1, u0) (unnormalised), l_b = 0          (:440-444)
//   |u| == 0       : f1 * I                                                                    (:446-449)
//   otherwise      : f1 + f2 |u| = 2 (eps - |u|) / eps^2 >= 0 along u/|u|,  f1 > 0 along ubar/|u|   (:451-462; makePD is the identity)
// factors of friction stencil i: Y = [y_a | y_b] (2 x 3nb doubles), vertex ids, sign of the block
template <int CLS>
__device__ __forceinline__ void friction_factor_one(const double4* __restrict__ X, const double4* __restrict__ Xn, const int4* __restrict__ fcs,
    const double2* __restrict__ fcp, const double* __restrict__ fB, const double* __restrict__ fnf, u32 i, double epsvh, double epsvh2, double mu,
    double* Y, int* v, bool& neg)
{
    const int4 c = fcs[i];
    const double2 cp = fcp[i];
    const FrStencil s = fr_decode(c, cp);
    const dv3 r = fr_rel3(X, Xn, c, s, cp);
    const FrBasis B = fr_load_basis(fB, i);
    const double u0 = dot(r, B.b0), u1 = dot(r, B.b1);
    const double x2 = u0 * u0 + u1 * u1, xn = sqrt(x2);
    const double f1 = fr_f1_div(x2, epsvh), f2 = -1.0 / (epsvh * epsvh);
    const double sc = (double)s.mult * mu * fnf[i], asc = fabs(sc);
    double qa[2], qb[2];
    if (x2 >= epsvh2) {
        const double k = sqrt(asc * f1 / x2);
        qa[0] = -k * u1; qa[1] = k * u0; qb[0] = qb[1] = 0.0;
    }
    else if (xn == 0.0) {
        const double k = sqrt(asc * f1);
        qa[0] = k; qa[1] = 0.0; qb[0] = 0.0; qb[1] = k;
    }
    else {
        const double la = fmax(f1 + f2 * xn, 0.0), lb = fmax(f1, 0.0);
        const double ka = sqrt(asc * la) / xn, kb = sqrt(asc * lb) / xn;
        qa[0] = ka * u0; qa[1] = ka * u1; qb[0] = -kb * u1; qb[1] = kb * u0;
    }
    const double wa[3] = {B.b0.x * qa[0] + B.b1.x * qa[1], B.b0.y * qa[0] + B.b1.y * qa[1], B.b0.z * qa[0] + B.b1.z * qa[1]};
    const double wb[3] = {B.b0.x * qb[0] + B.b1.x * qb[1], B.b0.y * qb[0] + B.b1.y * qb[1], B.b0.z * qb[0] + B.b1.z * qb[1]};
    constexpr int NB = (CLS == 0) ? 4 : (CLS == 1 ? 3 : 2), NN = 3 * NB;
#pragma unroll
    for (int k = 0; k < NB; ++k)
#pragma unroll
        for (int a = 0; a < 3; ++a) { Y[3 * k + a] = s.coef[k] * wa[a]; Y[NN + 3 * k + a] = s.coef[k] * wb[a]; }
    v[0] = s.v[0]; v[1] = s.v[1]; v[2] = s.v[2]; v[3] = s.v[3];
    neg = sc < 0.0;
}
template <int CLS>
__global__ void __launch_bounds__(128) k_friction_factor(const double4* __restrict__ X, const double4* __restrict__ Xn,
    const int4* __restrict__ fcs, const double2* __restrict__ fcp, const double* __restrict__ fB, const double* __restrict__ fnf,
    const u32* __restrict__ off, const u32* __restrict__ idx, u32 n, double epsvh, double epsvh2, double mu, double* __restrict__ Yout,
    YHdr* __restrict__ hdr)
{
    const u32 q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n) return;
    const u32 i = idx[q];
    constexpr int NB = (CLS == 0) ? 4 : (CLS == 1 ? 3 : 2), NN = 3 * NB;
    double Y[2 * NN];
    YHdr h;
    bool neg;
    friction_factor_one<CLS>(X, Xn, fcs, fcp, fB, fnf, i, epsvh, epsvh2, mu, Y, h.v, neg);
    double2* o = reinterpret_cast<double2*>(Yout + (size_t)q * (2 * NN));
#pragma unroll
    for (int k = 0; k < NN; ++k) o[k] = make_double2(Y[2 * k], Y[2 * k + 1]);
    h.off = off[i];
    h.pad[0] = neg ? 1 : 0; h.pad[1] = h.pad[2] = 0;
    hdr[q] = h;
}
// Fused factor + expansion (device-resident triplet stream, cipc_friction_hessian_dev): same structure as k_hessian_fused.
// The friction factors are cheap (no eigenproblem), so this kernel is a pure store stream.
template <int CLS> struct FrFusedShape {
    static constexpr int NN = 3 * ((CLS == 0) ? 4 : (CLS == 1 ? 3 : 2)), YD = 2 * NN, YS = YD + 1;
    static constexpr int SMEM = FUSED_BD * YS * 8 + FUSED_BD * 32;
};
template <int CLS, bool BLK>
__global__ void __launch_bounds__(FUSED_BD) k_friction_fused(const double4* __restrict__ X, const double4* __restrict__ Xn,
    const int4* __restrict__ fcs, const double2* __restrict__ fcp, const double* __restrict__ fB, const double* __restrict__ fnf,
    const u32* __restrict__ off, const u32* __restrict__ idx, u32 n, double epsvh, double epsvh2, double mu, void* __restrict__ outp)
{
    constexpr int NN = FrFusedShape<CLS>::NN, YS = FrFusedShape<CLS>::YS;
    extern __shared__ __align__(16) unsigned char fr_fused_sm[];
    double* sY = reinterpret_cast<double*>(fr_fused_sm);
    int* sH = reinterpret_cast<int*>(fr_fused_sm + FUSED_BD * YS * 8);
    const u32 q0 = blockIdx.x * FUSED_BD;
    const u32 g = min((u32)FUSED_BD, n - q0);
    if (threadIdx.x < g) {
        const u32 i = idx[q0 + threadIdx.x];
        double Y[2 * NN];
        int* h = sH + threadIdx.x * 8;
        bool neg;
        friction_factor_one<CLS>(X, Xn, fcs, fcp, fB, fnf, i, epsvh, epsvh2, mu, Y, h + 1, neg);
        h[0] = (int)off[i];
        h[5] = neg ? 1 : 0;
        if (BLK) h[6] = swap_mask<NN / 3>(h + 1);
        double* y = sY + threadIdx.x * YS;
#pragma unroll
        for (int k = 0; k < 2 * NN; ++k) y[k] = Y[k];
    }
    __syncwarp();
    if (BLK) warp_expand_blocks<NN / 3, 2, YS, true>(sY, sH, threadIdx.x & ~31u, g, threadIdx.x & 31u, reinterpret_cast<double*>(outp)); // upper blocks (merge.cuh)
    else warp_expand_stencils<NN, 2, YS, true, false>(sY, sH, nullptr, threadIdx.x & ~31u, g, threadIdx.x & 31u, reinterpret_cast<cipc_triplet*>(outp));
}

} // namespace cipc
