// hess.cuh -- barrier-Hessian blocks with PSD projection through their exact low-rank structure.
//
// Reference semantics (Library/FEM/IPC.h:1390-1731): per stencil H = w m (b'' g g^T + b' K), g = grad d,
// K = hess d (d = squared distance), then makePD (Library/Math/UTILS.h:9-27: V max(L,0) V^T from a dense
// 12x12 / 9x9 / 6x6 eigen-decomposition).
//
// This file never forms the dense block to decompose it.  For a point-triangle or edge-edge stencil
// write d = s^2, s = w . n^ the signed distance to the plane spanned by u, v (y = (w,u,v) the
// difference variables of geom.cuh).  Then, with t1,t2 an orthonormal basis of the plane, zeta =
// (1,-a,-b) (w_tangential = a u + b v), eta_k = (0, t_k.p_u, t_k.p_v), p_u = n^ x v/|n|, p_v = u x n^/|n|:
//
//     grad s = zeta (x) n^          hess s = sum_k [ (zeta(x)t_k)(eta_k(x)n^)^T + sym ] - s sum_k (eta_k(x)n^)(eta_k(x)n^)^T
//
// so   H = B C B^T   with only FIVE basis vectors  B = [zeta(x)n^, eta_1(x)n^, eta_2(x)n^, zeta(x)t1, zeta(x)t2]
// (pulled back to the four vertices by the constant +-1 map) and a 5x5 coefficient matrix C.  H has
// rank 5 (3 positive, 2 negative eigenvalues).  Orthonormalising B (Cholesky of its Gram matrix, which
// is block diagonal because n^, t1, t2 are orthogonal) turns makePD into a 5x5 symmetric eigenproblem
// solved by Jacobi rotations entirely in registers:  H+ = B L^-T (L^T C L)+ L^-1 B^T.
// Point-edge stencils have the same form with 4 vectors (rank 4), point-point is closed form.
// The result equals the reference's dense projection up to rounding (tests: <= 1e-12 relative);
// flops drop from ~20K (12x12 QL) to ~3K, which moves this kernel from the FP64 pipe to the HBM roofline.
//
// Mollified stencils (rare) add the mollifier's own Hessian and keep the dense path (eig.cuh).
#pragma once
#include "geom.cuh"

#ifndef CIPC_JACOBI_RR
#define CIPC_JACOBI_RR 1
#endif

namespace cipc {

// ------------------------------------------------------------------ K x K symmetric Jacobi, registers only
// A (full symmetric storage) is replaced by its positive-semidefinite part V max(L,0) V^T.
template <int K>
CIPC_HD void jacobi_psd_small(double (&A)[K][K])
{
    double V[K][K];
#pragma unroll
    for (int i = 0; i < K; ++i)
#pragma unroll
        for (int j = 0; j < K; ++j) V[i][j] = (i == j) ? 1.0 : 0.0;
    double fro = 0.0;
#pragma unroll
    for (int i = 0; i < K; ++i)
#pragma unroll
        for (int j = 0; j < K; ++j) fro += A[i][j] * A[i][j];
    if (fro == 0.0) return;
    const double tol = 1e-30 * fro;
    // rotation angles in fp32 on entries scaled by 1/||A||, rotations orthogonal to fp64 rounding and applied as exact
    // similarity transforms (see hess4_factor): the sweeps converge to the fp64 tolerance above
    const double iscale = cipc_rsqrt(fro);
    for (int sweep = 0; sweep < 14; ++sweep) {
        double off = 0.0;
#pragma unroll
        for (int p = 0; p < K - 1; ++p)
#pragma unroll
            for (int q = p + 1; q < K; ++q) off += A[p][q] * A[p][q];
        if (off <= tol) break;
#pragma unroll
        for (int p = 0; p < K - 1; ++p)
#pragma unroll
            for (int q = p + 1; q < K; ++q) {
                const double apq = A[p][q];
                if (apq * apq > 1e-36 * fro) {
                    const double app = A[p][p], aqq = A[q][q];
                    const float df = (float)((aqq - app) * iscale), af = (float)(apq * iscale) * 2.0f;
#if defined(__CUDA_ARCH__)
                    const float h2 = fmaf(df, df, af * af);
                    const float tf = __fdividef(df >= 0.0f ? af : -af, fabsf(df) + h2 * rsqrtf(h2));
#else
                    const float tf = (df >= 0.0f ? af : -af) / (fabsf(df) + sqrtf(df * df + af * af));
#endif
                    const double t = (double)tf;
                    const double c = cipc_rsqrt(t * t + 1.0), s = t * c;
                    const double cc = c * c, ss = s * s, cs = c * s;
                    A[p][p] = cc * app - 2.0 * cs * apq + ss * aqq;
                    A[q][q] = ss * app + 2.0 * cs * apq + cc * aqq;
                    const double npq = cs * (app - aqq) + (cc - ss) * apq;
                    A[p][q] = npq;
                    A[q][p] = npq;
#pragma unroll
                    for (int k = 0; k < K; ++k) {
                        if (k != p && k != q) {
                            const double akp = A[k][p], akq = A[k][q];
                            const double np_ = c * akp - s * akq, nq_ = s * akp + c * akq;
                            A[k][p] = np_; A[p][k] = np_;
                            A[k][q] = nq_; A[q][k] = nq_;
                        }
                        const double vkp = V[k][p], vkq = V[k][q];
                        V[k][p] = c * vkp - s * vkq;
                        V[k][q] = s * vkp + c * vkq;
                    }
                }
            }
    }
    double lam[K];
#pragma unroll
    for (int i = 0; i < K; ++i) lam[i] = A[i][i] > 0.0 ? A[i][i] : 0.0;
#pragma unroll
    for (int i = 0; i < K; ++i)
#pragma unroll
        for (int j = i; j < K; ++j) {
            double s = 0.0;
#pragma unroll
            for (int k = 0; k < K; ++k) s += lam[k] * V[i][k] * V[j][k];
            A[i][j] = s;
            A[j][i] = s;
        }
}

// C <- L^-T (L^T C L)+ L^-1 where G = L L^T is the Gram matrix of the basis (symmetric positive definite)
template <int K>
CIPC_HD void project_coefficients(const double (&G)[K][K], double (&C)[K][K])
{
    double L[K][K];
#pragma unroll
    for (int i = 0; i < K; ++i)
#pragma unroll
        for (int j = 0; j < K; ++j) L[i][j] = 0.0;
#pragma unroll
    for (int j = 0; j < K; ++j) { // Cholesky, column by column
        double d = G[j][j];
#pragma unroll
        for (int k = 0; k < j; ++k) d -= L[j][k] * L[j][k];
        d = sqrt(d);
        L[j][j] = d;
        const double id = 1.0 / d;
#pragma unroll
        for (int i = j + 1; i < K; ++i) {
            double s = G[i][j];
#pragma unroll
            for (int k = 0; k < j; ++k) s -= L[i][k] * L[j][k];
            L[i][j] = s * id;
        }
    }
    // M = L^T C L
    double T[K][K], M[K][K];
#pragma unroll
    for (int i = 0; i < K; ++i)
#pragma unroll
        for (int j = 0; j < K; ++j) {
            double s = 0.0;
#pragma unroll
            for (int k = j; k < K; ++k) s += C[i][k] * L[k][j]; // (C L)_ij, L lower
            T[i][j] = s;
        }
#pragma unroll
    for (int i = 0; i < K; ++i)
#pragma unroll
        for (int j = i; j < K; ++j) {
            double s = 0.0;
#pragma unroll
            for (int k = i; k < K; ++k) s += L[k][i] * T[k][j]; // (L^T T)_ij
            M[i][j] = s;
            M[j][i] = s;
        }
    jacobi_psd_small<K>(M);
    // Li = L^-1 (lower)
    double Li[K][K];
#pragma unroll
    for (int i = 0; i < K; ++i)
#pragma unroll
        for (int j = 0; j < K; ++j) Li[i][j] = 0.0;
#pragma unroll
    for (int j = 0; j < K; ++j) {
        Li[j][j] = 1.0 / L[j][j];
#pragma unroll
        for (int i = j + 1; i < K; ++i) {
            double s = 0.0;
#pragma unroll
            for (int k = j; k < i; ++k) s += L[i][k] * Li[k][j];
            Li[i][j] = -s / L[i][i];
        }
    }
    // C = Li^T M Li
#pragma unroll
    for (int i = 0; i < K; ++i)
#pragma unroll
        for (int j = 0; j < K; ++j) {
            double s = 0.0;
#pragma unroll
            for (int k = j; k < K; ++k) s += M[i][k] * Li[k][j];
            T[i][j] = s;
        }
#pragma unroll
    for (int i = 0; i < K; ++i)
#pragma unroll
        for (int j = i; j < K; ++j) {
            double s = 0.0;
#pragma unroll
            for (int k = i; k < K; ++k) s += Li[k][i] * T[k][j];
            C[i][j] = s;
            C[j][i] = s;
        }
}

// ------------------------------------------------------------------ 4-point stencils (PT / EE)
// Emits the 16 blocks through `emit(I, J, B)` with B a row-major 3x3 (double[9]).
// alpha = w m b'', beta = w m b'.
template <class Emit>
CIPC_HD void hess4_lowrank(bool ee, const dv3* x, double alpha, double beta, bool project, Emit emit)
{
    dv3 w, u, v;
    if (!ee) { w = x[0] - x[1]; u = x[2] - x[1]; v = x[3] - x[1]; }
    else { w = x[2] - x[0]; u = x[1] - x[0]; v = x[3] - x[2]; }
    const dv3 n = cross(u, v);
    const double nn = norm2(n), inn = 1.0 / sqrt(nn);
    const dv3 nh(n.x * inn, n.y * inn, n.z * inn);
    const double s = dot(w, nh);
    const dv3 cu = cross(nh, v), cv = cross(u, nh);
    const dv3 pu(cu.x * inn, cu.y * inn, cu.z * inn), pv(cv.x * inn, cv.y * inn, cv.z * inn);
    const double a = -dot(pu, w), b = -dot(pv, w);
    const double iu = 1.0 / sqrt(norm2(u));
    const dv3 t1(u.x * iu, u.y * iu, u.z * iu);
    const dv3 t2 = cross(nh, t1);
    // y-space coefficient 3-vectors and their pull-back to the 4 vertices
    const double zy[3] = {1.0, -a, -b};
    const double e1y[3] = {0.0, dot(t1, pu), dot(t1, pv)};
    const double e2y[3] = {0.0, dot(t2, pu), dot(t2, pv)};
    double cz[4], c1[4], c2[4];
    if (!ee) {
        cz[0] = zy[0]; cz[1] = -zy[0] - zy[1] - zy[2]; cz[2] = zy[1]; cz[3] = zy[2];
        c1[0] = 0.0; c1[1] = -e1y[1] - e1y[2]; c1[2] = e1y[1]; c1[3] = e1y[2];
        c2[0] = 0.0; c2[1] = -e2y[1] - e2y[2]; c2[2] = e2y[1]; c2[3] = e2y[2];
    }
    else {
        cz[0] = -zy[0] - zy[1]; cz[1] = zy[1]; cz[2] = zy[0] - zy[2]; cz[3] = zy[2];
        c1[0] = -e1y[1]; c1[1] = e1y[1]; c1[2] = -e1y[2]; c1[3] = e1y[2];
        c2[0] = -e2y[1]; c2[1] = e2y[1]; c2[2] = -e2y[2]; c2[3] = e2y[2];
    }
    // coefficient matrix in the basis [cz(x)n, c1(x)n, c2(x)n, cz(x)t1, cz(x)t2]
    double C[5][5];
#pragma unroll
    for (int i = 0; i < 5; ++i)
#pragma unroll
        for (int j = 0; j < 5; ++j) C[i][j] = 0.0;
    const double gam = 2.0 * beta * s, mu = 2.0 * beta * s * s;
    C[0][0] = 4.0 * alpha * s * s + 2.0 * beta;
    C[1][1] = -mu; C[2][2] = -mu;
    C[1][3] = gam; C[3][1] = gam; C[2][4] = gam; C[4][2] = gam;
    if (project) {
        double G[5][5];
#pragma unroll
        for (int i = 0; i < 5; ++i)
#pragma unroll
            for (int j = 0; j < 5; ++j) G[i][j] = 0.0;
        double gzz = 0, gz1 = 0, gz2 = 0, g11 = 0, g12 = 0, g22 = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            gzz += cz[k] * cz[k]; gz1 += cz[k] * c1[k]; gz2 += cz[k] * c2[k];
            g11 += c1[k] * c1[k]; g12 += c1[k] * c2[k]; g22 += c2[k] * c2[k];
        }
        G[0][0] = gzz; G[0][1] = gz1; G[1][0] = gz1; G[0][2] = gz2; G[2][0] = gz2;
        G[1][1] = g11; G[1][2] = g12; G[2][1] = g12; G[2][2] = g22;
        G[3][3] = gzz; G[4][4] = gzz;
        project_coefficients<5>(G, C);
    }
    // blocks: (I,J) = R W R^T, R = [n t1 t2]
    const double R[3][3] = {{nh.x, t1.x, t2.x}, {nh.y, t1.y, t2.y}, {nh.z, t1.z, t2.z}};
#pragma unroll
    for (int I = 0; I < 4; ++I) {
        const double kI[3] = {cz[I], c1[I], c2[I]};
        // row vector kI^T C over the N-type indices, and its couplings to the two T-type vectors
        double rN[3], r3 = 0.0, r4 = 0.0;
#pragma unroll
        for (int j = 0; j < 3; ++j) rN[j] = kI[0] * C[0][j] + kI[1] * C[1][j] + kI[2] * C[2][j];
#pragma unroll
        for (int i = 0; i < 3; ++i) { r3 += kI[i] * C[i][3]; r4 += kI[i] * C[i][4]; }
#pragma unroll
        for (int J = 0; J < 4; ++J) {
            const double kJ[3] = {cz[J], c1[J], c2[J]};
            double c3 = 0.0, c4 = 0.0;
#pragma unroll
            for (int j = 0; j < 3; ++j) { c3 += C[3][j] * kJ[j]; c4 += C[4][j] * kJ[j]; }
            double W[3][3];
            W[0][0] = rN[0] * kJ[0] + rN[1] * kJ[1] + rN[2] * kJ[2];
            W[0][1] = r3 * cz[J]; W[0][2] = r4 * cz[J];
            W[1][0] = cz[I] * c3; W[2][0] = cz[I] * c4;
            const double zz = cz[I] * cz[J];
            W[1][1] = C[3][3] * zz; W[1][2] = C[3][4] * zz; W[2][1] = C[4][3] * zz; W[2][2] = C[4][4] * zz;
            double T[3][3], Bk[9];
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
                for (int q = 0; q < 3; ++q) T[r][q] = R[r][0] * W[0][q] + R[r][1] * W[1][q] + R[r][2] * W[2][q];
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
                for (int c = 0; c < 3; ++c) Bk[3 * r + c] = T[r][0] * R[c][0] + T[r][1] * R[c][1] + T[r][2] * R[c][2];
            emit(I, J, Bk);
        }
    }
}

// ------------------------------------------------------------------ factored form for the two-phase kernel
// Projected 4-point block as a sum of at most three outer products  H+ = sum_k y_k y_k^T  (the block has exactly
// three non-negative directions).  Y: 3 x 12 doubles.  Same algebra as hess4_lowrank, with the block structure of
// the Gram matrix (3x3 + two equal scalars) and the closed forms of eta_1, eta_2 written out by hand so that the
// whole computation stays in registers:
//     eta_1 = (0, -1/|u|, 0),  eta_2 = (0, (u^.v)/|n|, -|u|/|n|)        (t1 = u^, t2 = n^ x u^)
// STASH: the values that are only needed again after the eigen-iteration (basis coefficients, frame vectors, Cholesky
// factors: 28 doubles) are parked in the output slot Y (volatile accesses: no forwarding through registers) while the 5x5
// Jacobi runs and are read back before Y is written.  With Y in shared memory (k_hessian_fused) this takes ~56 registers
// off the iteration's live set: the kernel fits five CTAs per SM without spilling.
template <bool STASH = false>
CIPC_HD void hess4_factor(bool ee, const dv3* x, double alpha, double beta, double* Y /*36*/)
{
    dv3 w, u, v;
    if (!ee) { w = x[0] - x[1]; u = x[2] - x[1]; v = x[3] - x[1]; }
    else { w = x[2] - x[0]; u = x[1] - x[0]; v = x[3] - x[2]; }
    const dv3 n = cross(u, v);
    const double nn = norm2(n), inn = 1.0 / sqrt(nn);
    const dv3 nh(n.x * inn, n.y * inn, n.z * inn);
    const double s = dot(w, nh);
    const double uu = norm2(u), iu = 1.0 / sqrt(uu);
    const dv3 t1(u.x * iu, u.y * iu, u.z * iu);
    const dv3 t2 = cross(nh, t1);
    // tangential part of w in the (u,v) frame: w_t = a u + b v  via the dual basis
    const dv3 cu = cross(nh, v), cv = cross(u, nh);
    const double a = -dot(cu, w) * inn, b = -dot(cv, w) * inn;
    const double e1u = -iu;                       // eta_1 = (0, e1u, 0)
    const double e2u = dot(t1, v) * inn, e2v = -uu * iu * inn; // eta_2 = (0, e2u, e2v)
    double cz[4], c1[4], c2[4];
    if (!ee) {
        cz[0] = 1.0; cz[1] = -1.0 + a + b; cz[2] = -a; cz[3] = -b;
        c1[0] = 0.0; c1[1] = -e1u; c1[2] = e1u; c1[3] = 0.0;
        c2[0] = 0.0; c2[1] = -e2u - e2v; c2[2] = e2u; c2[3] = e2v;
    }
    else {
        cz[0] = -1.0 + a; cz[1] = -a; cz[2] = 1.0 + b; cz[3] = -b;
        c1[0] = -e1u; c1[1] = e1u; c1[2] = 0.0; c1[3] = 0.0;
        c2[0] = -e2u; c2[1] = e2u; c2[2] = -e2v; c2[3] = e2v;
    }
    double gzz = 0, gz1 = 0, gz2 = 0, g11 = 0, g12 = 0, g22 = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        gzz += cz[k] * cz[k]; gz1 += cz[k] * c1[k]; gz2 += cz[k] * c2[k];
        g11 += c1[k] * c1[k]; g12 += c1[k] * c2[k]; g22 += c2[k] * c2[k];
    }
    // Cholesky of the 3x3 Gram block
    const double l00 = sqrt(gzz), il00 = 1.0 / l00;
    const double l10 = gz1 * il00, l20 = gz2 * il00;
    const double l11 = sqrt(g11 - l10 * l10), il11 = 1.0 / l11;
    const double l21 = (g12 - l20 * l10) * il11;
    const double l22 = sqrt(g22 - l20 * l20 - l21 * l21), il22 = 1.0 / l22;
    // M = L^T C L in the orthonormalised basis (upper triangle, 5x5)
    const double lam0 = 4.0 * alpha * s * s + 2.0 * beta, mu = 2.0 * beta * s * s, gg = l00 * 2.0 * beta * s;
    double A[5][5];
    A[0][0] = lam0 * l00 * l00 - mu * (l10 * l10 + l20 * l20);
    A[0][1] = -mu * (l10 * l11 + l20 * l21);
    A[0][2] = -mu * l20 * l22;
    A[1][1] = -mu * (l11 * l11 + l21 * l21);
    A[1][2] = -mu * l21 * l22;
    A[2][2] = -mu * l22 * l22;
    A[0][3] = gg * l10; A[0][4] = gg * l20;
    A[1][3] = gg * l11; A[1][4] = gg * l21;
    A[2][3] = 0.0; A[2][4] = gg * l22;
    A[3][3] = 0.0; A[3][4] = 0.0; A[4][4] = 0.0;
    if (STASH) {
        volatile double* st = Y;
#pragma unroll
        for (int k = 0; k < 4; ++k) { st[k] = cz[k]; st[4 + k] = c1[k]; st[8 + k] = c2[k]; }
        st[12] = nh.x; st[13] = nh.y; st[14] = nh.z; st[15] = t1.x; st[16] = t1.y; st[17] = t1.z; st[18] = t2.x; st[19] = t2.y; st[20] = t2.z;
        st[21] = l10; st[22] = l20; st[23] = l11; st[24] = l21; st[25] = il00; st[26] = il11; st[27] = il22;
    }
    double V[5][5];
#pragma unroll
    for (int i = 0; i < 5; ++i)
#pragma unroll
        for (int j = 0; j < 5; ++j) V[i][j] = (i == j) ? 1.0 : 0.0;
    double fro = 0.0;
#pragma unroll
    for (int i = 0; i < 5; ++i)
#pragma unroll
        for (int j = i; j < 5; ++j) fro += (i == j ? 1.0 : 2.0) * A[i][j] * A[i][j];
    const double tol = 1e-26 * fro; // off-diagonal norm <= 1e-13 ||M||: eigenvector error far below the 1e-9 gate
    // Rotation angles are formed in fp32 (MUFU reciprocal / rsqrt on entries scaled by 1/||M||): an fp64 division and
    // square root cost more instructions than applying the rotation.  The rotation itself stays orthogonal to fp64
    // rounding (c = rsqrt(1 + t^2), s = t c in fp64) and is applied as an exact similarity transform (a_pq is updated,
    // not zeroed), so an angle error of 1e-7 only leaves a_pq' ~ 1e-7 a_pq: the sweeps still converge quadratically
    // down to the fp64 tolerance above, which is what bounds the result's accuracy.
    const double iscale = cipc_rsqrt(fro);
#if CIPC_JACOBI_RR
    // Round-robin ordering: five rounds of two DISJOINT index pairs.  The two rotations of a round commute and neither
    // changes the three entries the other's angle is formed from, so both angles are computed first -- two independent
    // fp32 / rsqrt latency chains the scheduler interleaves -- and then both rotations are applied.  Branch-free: a pair
    // whose off-diagonal entry is already negligible gets the identity rotation (t = 0).
    constexpr int RP[10] = {1, 2, 0, 3, 1, 0, 2, 0, 0, 1}, RQ[10] = {4, 3, 2, 4, 3, 4, 4, 1, 3, 2};
    for (int sweep = 0; sweep < 12; ++sweep) {
        double off = 0.0;
#pragma unroll
        for (int p = 0; p < 4; ++p)
#pragma unroll
            for (int q = p + 1; q < 5; ++q) off += A[p][q] * A[p][q];
        if (off <= tol) break;
        const double thr = 1e-34 * fro;
#pragma unroll
        for (int r = 0; r < 5; ++r) {
            double cr[2], sr[2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int p = RP[2 * r + h], q = RQ[2 * r + h];
                const double apq = A[p][q], app = A[p][p], aqq = A[q][q];
                const float df = (float)((aqq - app) * iscale), af = (float)(apq * iscale) * 2.0f;
#if defined(__CUDA_ARCH__)
                const float h2 = fmaxf(fmaf(df, df, af * af), 1e-37f);
                const float tf = __fdividef(df >= 0.0f ? af : -af, fabsf(df) + h2 * rsqrtf(h2));
#else
                const float h2 = fmaxf(df * df + af * af, 1e-37f);
                const float tf = (df >= 0.0f ? af : -af) / (fabsf(df) + sqrtf(h2));
#endif
                const double t = (apq * apq > thr) ? (double)tf : 0.0;
                cr[h] = cipc_rsqrt(t * t + 1.0);
                sr[h] = t * cr[h];
            }
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int p = RP[2 * r + h], q = RQ[2 * r + h];
                const double c = cr[h], sn = sr[h];
                const double apq = A[p][q], app = A[p][p], aqq = A[q][q];
                const double cc = c * c, ss = sn * sn, cs = c * sn;
                A[p][p] = cc * app - 2.0 * cs * apq + ss * aqq;
                A[q][q] = ss * app + 2.0 * cs * apq + cc * aqq;
                A[p][q] = cs * (app - aqq) + (cc - ss) * apq;
#pragma unroll
                for (int k = 0; k < 5; ++k) {
                    if (k != p && k != q) { // upper-triangle accessors with compile-time indices
                        double& akp = (k < p) ? A[k][p] : A[p][k];
                        double& akq = (k < q) ? A[k][q] : A[q][k];
                        const double vp = akp, vq = akq;
                        akp = c * vp - sn * vq;
                        akq = sn * vp + c * vq;
                    }
                    const double vkp = V[k][p], vkq = V[k][q];
                    V[k][p] = c * vkp - sn * vkq;
                    V[k][q] = sn * vkp + c * vkq;
                }
            }
        }
    }
#else
    for (int sweep = 0; sweep < 12; ++sweep) {
        double off = 0.0;
#pragma unroll
        for (int p = 0; p < 4; ++p)
#pragma unroll
            for (int q = p + 1; q < 5; ++q) off += A[p][q] * A[p][q];
        if (off <= tol) break;
#pragma unroll
        for (int p = 0; p < 4; ++p)
#pragma unroll
            for (int q = p + 1; q < 5; ++q) {
                const double apq = A[p][q];
                if (apq * apq > 1e-34 * fro) {
                    const double app = A[p][p], aqq = A[q][q];
                    const float df = (float)((aqq - app) * iscale), af = (float)(apq * iscale) * 2.0f;
#if defined(__CUDA_ARCH__)
                    const float h2 = fmaf(df, df, af * af);
                    const float tf = __fdividef(df >= 0.0f ? af : -af, fabsf(df) + h2 * rsqrtf(h2));
#else
                    const float tf = (df >= 0.0f ? af : -af) / (fabsf(df) + sqrtf(df * df + af * af));
#endif
                    const double t = (double)tf;
                    const double c = cipc_rsqrt(t * t + 1.0), sn = t * c;
                    const double cc = c * c, ss = sn * sn, cs = c * sn;
                    A[p][p] = cc * app - 2.0 * cs * apq + ss * aqq;
                    A[q][q] = ss * app + 2.0 * cs * apq + cc * aqq;
                    A[p][q] = cs * (app - aqq) + (cc - ss) * apq;
#pragma unroll
                    for (int k = 0; k < 5; ++k) {
                        if (k != p && k != q) { // upper-triangle accessors with compile-time indices
                            double& akp = (k < p) ? A[k][p] : A[p][k];
                            double& akq = (k < q) ? A[k][q] : A[q][k];
                            const double vp = akp, vq = akq;
                            akp = c * vp - sn * vq;
                            akq = sn * vp + c * vq;
                        }
                        const double vkp = V[k][p], vkq = V[k][q];
                        V[k][p] = c * vkp - sn * vkq;
                        V[k][q] = sn * vkp + c * vkq;
                    }
                }
            }
    }
#endif
    // bring the three largest eigenvalues (the only possibly positive ones) to columns 0..2
    double lam[5];
#pragma unroll
    for (int i = 0; i < 5; ++i) lam[i] = A[i][i];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = i + 1; j < 5; ++j) {
            const bool sw = lam[j] > lam[i];
            const double tl = lam[i];
            lam[i] = sw ? lam[j] : lam[i];
            lam[j] = sw ? tl : lam[j];
#pragma unroll
            for (int k = 0; k < 5; ++k) {
                const double tv = V[k][i];
                V[k][i] = sw ? V[k][j] : V[k][i];
                V[k][j] = sw ? tv : V[k][j];
            }
        }
    // the parked values come back before the slot is overwritten with the factors
    double zc[4], oc[4], tc[4], nx, ny, nz, ax, ay, az, bx, by, bz, m10, m20, m21, j00, j11, j22;
    if (STASH) {
        volatile double* st = Y;
#pragma unroll
        for (int k = 0; k < 4; ++k) { zc[k] = st[k]; oc[k] = st[4 + k]; tc[k] = st[8 + k]; }
        nx = st[12]; ny = st[13]; nz = st[14]; ax = st[15]; ay = st[16]; az = st[17]; bx = st[18]; by = st[19]; bz = st[20];
        m10 = st[21]; m20 = st[22]; m21 = st[24]; j00 = st[25]; j11 = st[26]; j22 = st[27];
    }
    else {
#pragma unroll
        for (int k = 0; k < 4; ++k) { zc[k] = cz[k]; oc[k] = c1[k]; tc[k] = c2[k]; }
        nx = nh.x; ny = nh.y; nz = nh.z; ax = t1.x; ay = t1.y; az = t1.z; bx = t2.x; by = t2.y; bz = t2.z;
        m10 = l10; m20 = l20; m21 = l21; j00 = il00; j11 = il11; j22 = il22;
    }
    double F[3][5];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const double sc = lam[k] > 0.0 ? sqrt(lam[k]) : 0.0;
        // f = L~^-T (sc v): back substitution with L~ = blockdiag(L_N, l00, l00)
        F[k][4] = sc * V[4][k] * j00; F[k][3] = sc * V[3][k] * j00;
        F[k][2] = sc * V[2][k] * j22;
        F[k][1] = (sc * V[1][k] - m21 * F[k][2]) * j11;
        F[k][0] = (sc * V[0][k] - m10 * F[k][1] - m20 * F[k][2]) * j00;
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const double ttx = F[k][3] * ax + F[k][4] * bx, tty = F[k][3] * ay + F[k][4] * by, ttz = F[k][3] * az + F[k][4] * bz;
#pragma unroll
        for (int I = 0; I < 4; ++I) {
            const double cn = F[k][0] * zc[I] + F[k][1] * oc[I] + F[k][2] * tc[I];
            Y[12 * k + 3 * I + 0] = cn * nx + zc[I] * ttx;
            Y[12 * k + 3 * I + 1] = cn * ny + zc[I] * tty;
            Y[12 * k + 3 * I + 2] = cn * nz + zc[I] * ttz;
        }
    }
}

// ------------------------------------------------------------------ point-edge stencil (rank 4)
// x = (p, e0, e1).  Basis [z(x)n, eps(x)n, z(x)u^, z(x)b^] with z = A^T(1,-t), eps = A^T(0,1).
template <class Emit>
CIPC_HD void hess_pe_lowrank(const dv3& p, const dv3& e0, const dv3& e1, double alpha, double beta, bool project, Emit emit)
{
    const dv3 w = p - e0, u = e1 - e0;
    const double L = norm2(u), iL = 1.0 / L, t = dot(w, u) * iL;
    const dv3 r(w.x - t * u.x, w.y - t * u.y, w.z - t * u.z);
    const double d = norm2(r), rn = sqrt(d), irn = 1.0 / rn, un = sqrt(L), iun = 1.0 / un;
    const dv3 nh(r.x * irn, r.y * irn, r.z * irn), uh(u.x * iun, u.y * iun, u.z * iun);
    const dv3 bh = cross(uh, nh);
    const double cz[3] = {1.0, -1.0 + t, -t};  // A^T (1,-t):  p: 1, e0: -1-(-t), e1: -t
    const double ce[3] = {0.0, -1.0, 1.0};     // A^T (0,1)
    double C[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) C[i][j] = 0.0;
    C[0][0] = 4.0 * alpha * d + 2.0 * beta;
    C[1][1] = -2.0 * beta * d * iL;
    C[1][2] = -2.0 * beta * rn * iun; C[2][1] = C[1][2];
    C[3][3] = 2.0 * beta;
    if (project) {
        double G[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) G[i][j] = 0.0;
        const double gzz = cz[0] * cz[0] + cz[1] * cz[1] + cz[2] * cz[2];
        const double gze = cz[1] * ce[1] + cz[2] * ce[2];
        G[0][0] = gzz; G[0][1] = gze; G[1][0] = gze; G[1][1] = 2.0; G[2][2] = gzz; G[3][3] = gzz;
        project_coefficients<4>(G, C);
    }
    // spatial triad R = [n, u^, b^]; basis spatial index: 0,0,1,2 ; coefficient vectors cz,ce,cz,cz
    const double R[3][3] = {{nh.x, uh.x, bh.x}, {nh.y, uh.y, bh.y}, {nh.z, uh.z, bh.z}};
#pragma unroll
    for (int I = 0; I < 3; ++I)
#pragma unroll
        for (int J = 0; J < 3; ++J) {
            double W[3][3];
            W[0][0] = cz[I] * (C[0][0] * cz[J] + C[0][1] * ce[J]) + ce[I] * (C[1][0] * cz[J] + C[1][1] * ce[J]);
            W[0][1] = (cz[I] * C[0][2] + ce[I] * C[1][2]) * cz[J];
            W[0][2] = (cz[I] * C[0][3] + ce[I] * C[1][3]) * cz[J];
            W[1][0] = cz[I] * (C[2][0] * cz[J] + C[2][1] * ce[J]);
            W[2][0] = cz[I] * (C[3][0] * cz[J] + C[3][1] * ce[J]);
            const double zz = cz[I] * cz[J];
            W[1][1] = C[2][2] * zz; W[1][2] = C[2][3] * zz; W[2][1] = C[3][2] * zz; W[2][2] = C[3][3] * zz;
            double T[3][3], Bk[9];
#pragma unroll
            for (int rr = 0; rr < 3; ++rr)
#pragma unroll
                for (int q = 0; q < 3; ++q) T[rr][q] = R[rr][0] * W[0][q] + R[rr][1] * W[1][q] + R[rr][2] * W[2][q];
#pragma unroll
            for (int rr = 0; rr < 3; ++rr)
#pragma unroll
                for (int c = 0; c < 3; ++c) Bk[3 * rr + c] = T[rr][0] * R[c][0] + T[rr][1] * R[c][1] + T[rr][2] * R[c][2];
            emit(I, J, Bk);
        }
}

// Projected point-edge block as at most two outer products (the block has two non-negative directions).
// Basis [cz(x)n, ce(x)n, cz(x)u^]; the fourth vector cz(x)b^ carries the eigenvalue 2 beta |cz|^2 < 0 and drops out.
CIPC_HD void hess_pe_factor(const dv3& p, const dv3& e0, const dv3& e1, double alpha, double beta, double* Y /*18*/)
{
    const dv3 w = p - e0, u = e1 - e0;
    const double L = norm2(u), iL = 1.0 / L, t = dot(w, u) * iL;
    const dv3 r(w.x - t * u.x, w.y - t * u.y, w.z - t * u.z);
    const double d = norm2(r), rn = sqrt(d), irn = 1.0 / rn, iun = cipc_rsqrt(L);
    const dv3 nh(r.x * irn, r.y * irn, r.z * irn), uh(u.x * iun, u.y * iun, u.z * iun);
    const double cz[3] = {1.0, -1.0 + t, -t};
    const double ce[3] = {0.0, -1.0, 1.0};
    const double gzz = cz[0] * cz[0] + cz[1] * cz[1] + cz[2] * cz[2], gze = cz[1] * ce[1] + cz[2] * ce[2];
    const double l00 = sqrt(gzz), il00 = 1.0 / l00, l10 = gze * il00, l11 = sqrt(2.0 - l10 * l10), il11 = 1.0 / l11;
    const double lam1 = 4.0 * alpha * d + 2.0 * beta, c11 = -2.0 * beta * d * iL, c12 = -2.0 * beta * rn * iun;
    double A[3][3];
    A[0][0] = l00 * l00 * lam1 + l10 * l10 * c11;
    A[0][1] = l10 * c11 * l11;
    A[0][2] = l10 * c12 * l00;
    A[1][1] = l11 * l11 * c11;
    A[1][2] = l11 * c12 * l00;
    A[2][2] = 0.0;
    double V[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    const double fro = A[0][0] * A[0][0] + A[1][1] * A[1][1] + 2.0 * (A[0][1] * A[0][1] + A[0][2] * A[0][2] + A[1][2] * A[1][2]);
    const double tol = 1e-26 * fro;
    for (int sweep = 0; sweep < 10; ++sweep) {
        const double off = A[0][1] * A[0][1] + A[0][2] * A[0][2] + A[1][2] * A[1][2];
        if (off <= tol) break;
#pragma unroll
        for (int pI = 0; pI < 2; ++pI)
#pragma unroll
            for (int q = pI + 1; q < 3; ++q) {
                const double apq = A[pI][q];
                if (apq * apq > 1e-34 * fro) {
                    const double theta = (A[q][q] - A[pI][pI]) / (2.0 * apq);
                    const double tt = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                    const double c = cipc_rsqrt(tt * tt + 1.0), sn = tt * c;
                    A[pI][pI] -= tt * apq;
                    A[q][q] += tt * apq;
                    A[pI][q] = 0.0;
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        if (k != pI && k != q) {
                            double& akp = (k < pI) ? A[k][pI] : A[pI][k];
                            double& akq = (k < q) ? A[k][q] : A[q][k];
                            const double vp = akp, vq = akq;
                            akp = c * vp - sn * vq;
                            akq = sn * vp + c * vq;
                        }
                        const double vkp = V[k][pI], vkq = V[k][q];
                        V[k][pI] = c * vkp - sn * vkq;
                        V[k][q] = sn * vkp + c * vkq;
                    }
                }
            }
    }
    double lam[3] = {A[0][0], A[1][1], A[2][2]};
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = i + 1; j < 3; ++j) {
            const bool sw = lam[j] > lam[i];
            const double tl = lam[i];
            lam[i] = sw ? lam[j] : lam[i];
            lam[j] = sw ? tl : lam[j];
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const double tv = V[k][i];
                V[k][i] = sw ? V[k][j] : V[k][i];
                V[k][j] = sw ? tv : V[k][j];
            }
        }
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const double sc = lam[k] > 0.0 ? sqrt(lam[k]) : 0.0;
        const double f2 = sc * V[2][k] * il00;
        const double f1 = sc * V[1][k] * il11;
        const double f0 = (sc * V[0][k] - l10 * f1) * il00;
#pragma unroll
        for (int I = 0; I < 3; ++I) {
            const double cn = f0 * cz[I] + f1 * ce[I], cu_ = f2 * cz[I];
            Y[9 * k + 3 * I + 0] = cn * nh.x + cu_ * uh.x;
            Y[9 * k + 3 * I + 1] = cn * nh.y + cu_ * uh.y;
            Y[9 * k + 3 * I + 2] = cn * nh.z + cu_ * uh.z;
        }
    }
}
// Projected point-point block: one outer product, y = sqrt(max(4 alpha |d|^2 + 2 beta, 0)/(|d|^2)) (d, -d)   (beta < 0)
CIPC_HD void hess_pp_factor(const dv3& a, const dv3& b, double alpha, double beta, double* Y /*6*/)
{
    const dv3 dl = a - b;
    const double d2 = norm2(dl);
    const double lam = 4.0 * alpha * d2 + 2.0 * beta;
    const double k = lam > 0.0 ? sqrt(lam / d2) : 0.0;
    Y[0] = k * dl.x; Y[1] = k * dl.y; Y[2] = k * dl.z; Y[3] = -Y[0]; Y[4] = -Y[1]; Y[5] = -Y[2];
}

// ------------------------------------------------------------------ point-point stencil (closed form)
// H = A^T (4 alpha dd^T + 2 beta I) A, A = [I -I]; projected: max(4 alpha |d|^2 + 2 beta, 0) d^d^^T (x) [[1,-1],[-1,1]]
template <class Emit>
CIPC_HD void hess_pp_closed(const dv3& a, const dv3& b, double alpha, double beta, bool project, Emit emit)
{
    const dv3 dl = a - b;
    const double d2 = norm2(dl);
    double M[9];
    if (project) {
        const double lam = 4.0 * alpha * d2 + 2.0 * beta;
        const double k = (lam > 0.0 ? lam : 0.0) / d2;
        const double tb = (beta > 0.0) ? 2.0 * beta : 0.0; // the two transverse eigenvalues 2 beta (negative for a barrier)
        M[0] = k * dl.x * dl.x; M[1] = k * dl.x * dl.y; M[2] = k * dl.x * dl.z;
        M[3] = M[1]; M[4] = k * dl.y * dl.y; M[5] = k * dl.y * dl.z;
        M[6] = M[2]; M[7] = M[5]; M[8] = k * dl.z * dl.z;
        if (tb != 0.0) {
            const double id2 = 1.0 / d2;
            M[0] += tb * (1.0 - dl.x * dl.x * id2); M[4] += tb * (1.0 - dl.y * dl.y * id2); M[8] += tb * (1.0 - dl.z * dl.z * id2);
            M[1] -= tb * dl.x * dl.y * id2; M[3] = M[1]; M[2] -= tb * dl.x * dl.z * id2; M[6] = M[2]; M[5] -= tb * dl.y * dl.z * id2; M[7] = M[5];
        }
    }
    else {
        const double k = 4.0 * alpha;
        M[0] = k * dl.x * dl.x + 2.0 * beta; M[1] = k * dl.x * dl.y; M[2] = k * dl.x * dl.z;
        M[3] = M[1]; M[4] = k * dl.y * dl.y + 2.0 * beta; M[5] = k * dl.y * dl.z;
        M[6] = M[2]; M[7] = M[5]; M[8] = k * dl.z * dl.z + 2.0 * beta;
    }
    double Mn[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) Mn[i] = -M[i];
    emit(0, 0, M); emit(0, 1, Mn); emit(1, 0, Mn); emit(1, 1, M);
}

} // namespace cipc
