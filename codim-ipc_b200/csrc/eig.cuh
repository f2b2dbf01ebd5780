// eig.cuh -- PSD projection of the small symmetric barrier-Hessian blocks (makePD,
// Library/Math/UTILS.h:9-27): keep the matrix when its smallest eigenvalue is >= 0, otherwise
// rebuild it from the non-negative part of its spectrum.
//
// The reference calls Eigen::SelfAdjointEigenSolver (tridiagonal QL); V max(L,0) V^T is a
// well-conditioned function of the matrix, so this path uses a cyclic Jacobi iteration instead
// (backward stable, branch-light, no dynamic deflation logic), agreeing to ~1e-14 ||H||.
#pragma once
#include "geom.cuh"

namespace cipc {

// A: N x N row-major symmetric, overwritten by its PSD projection.
template <int N>
CIPC_HD void psd_project_jacobi(double* A)
{
    double V[N * N];
    for (int i = 0; i < N * N; ++i) V[i] = 0.0;
    for (int i = 0; i < N; ++i) V[i * N + i] = 1.0;
    double fro = 0.0;
    for (int i = 0; i < N * N; ++i) fro += A[i] * A[i];
    if (fro == 0.0) return;
    const double tol = 1e-30 * fro;
    for (int sweep = 0; sweep < 16; ++sweep) {
        double off = 0.0;
        for (int p = 0; p < N; ++p)
            for (int q = p + 1; q < N; ++q) off += A[p * N + q] * A[p * N + q];
        if (off <= tol) break;
        for (int p = 0; p < N - 1; ++p)
            for (int q = p + 1; q < N; ++q) {
                const double apq = A[p * N + q];
                if (apq == 0.0) continue;
                const double app = A[p * N + p], aqq = A[q * N + q];
                if (apq * apq <= 1e-34 * fabs(app * aqq) && sweep > 2) { A[p * N + q] = A[q * N + p] = 0.0; continue; }
                const double theta = (aqq - app) / (2.0 * apq);
                const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < N; ++k) { // columns p,q
                    const double akp = A[k * N + p], akq = A[k * N + q];
                    A[k * N + p] = c * akp - s * akq;
                    A[k * N + q] = s * akp + c * akq;
                }
                for (int k = 0; k < N; ++k) { // rows p,q
                    const double apk = A[p * N + k], aqk = A[q * N + k];
                    A[p * N + k] = c * apk - s * aqk;
                    A[q * N + k] = s * apk + c * aqk;
                }
                for (int k = 0; k < N; ++k) {
                    const double vkp = V[k * N + p], vkq = V[k * N + q];
                    V[k * N + p] = c * vkp - s * vkq;
                    V[k * N + q] = s * vkp + c * vkq;
                }
            }
    }
    double lam[N];
    bool anyNeg = false;
    for (int i = 0; i < N; ++i) { lam[i] = A[i * N + i]; anyNeg |= (lam[i] < 0.0); }
    // NOTE: A now holds the (nearly) diagonalised matrix, so it must be rebuilt in both cases.
    for (int i = 0; i < N; ++i)
        for (int j = i; j < N; ++j) {
            double s = 0.0;
            for (int k = 0; k < N; ++k) {
                const double l = anyNeg ? (lam[k] > 0.0 ? lam[k] : 0.0) : lam[k];
                s += l * V[i * N + k] * V[j * N + k];
            }
            A[i * N + j] = s;
            A[j * N + i] = s;
        }
}


// Symmetric eigenproblem by Householder tridiagonalisation + implicit-shift QL (the classical EISPACK tred2 / tql2 pair,
// which is also what Eigen's SelfAdjointEigenSolver -- the reference's makePD, Math/UTILS.h:9-27 -- amounts to): ~5x fewer
// operations than the cyclic Jacobi iteration at n = 9.  Z(i, j): accessor of an n x n matrix holding the symmetric input
// (full storage); on return its columns are the orthonormal eigenvectors, d[0..n) the eigenvalues (unsorted).  d, e: scratch
// of length n.  Returns false when an eigenvalue needed more than 40 QL steps (never observed; callers fall back to Jacobi).
template <class ZAcc>
CIPC_HD bool sym_eig_ql(int n, ZAcc Z, double* d, double* e)
{
    // ---- Householder reduction to tridiagonal form, last row first; the reflectors stay in Z
    for (int j = 0; j < n; ++j) d[j] = Z(n - 1, j);
    for (int i = n - 1; i > 0; --i) {
        double scale = 0.0, h = 0.0;
        for (int k = 0; k < i; ++k) scale += fabs(d[k]);
        if (scale == 0.0) {
            e[i] = d[i - 1];
            for (int j = 0; j < i; ++j) { d[j] = Z(i - 1, j); Z(i, j) = 0.0; Z(j, i) = 0.0; }
        }
        else {
            const double inv = 1.0 / scale;
            for (int k = 0; k < i; ++k) { d[k] *= inv; h += d[k] * d[k]; }
            double f = d[i - 1];
            double g = sqrt(h);
            if (f > 0.0) g = -g;
            e[i] = scale * g;
            h -= f * g;
            d[i - 1] = f - g;
            for (int j = 0; j < i; ++j) e[j] = 0.0;
            for (int j = 0; j < i; ++j) { // e = A u (lower triangle only)
                f = d[j];
                Z(j, i) = f;
                g = e[j] + Z(j, j) * f;
                for (int k = j + 1; k < i; ++k) { const double zkj = Z(k, j); g += zkj * d[k]; e[k] += zkj * f; }
                e[j] = g;
            }
            f = 0.0;
            const double invh = 1.0 / h;
            for (int j = 0; j < i; ++j) { e[j] *= invh; f += e[j] * d[j]; }
            const double hh = f / (h + h);
            for (int j = 0; j < i; ++j) e[j] -= hh * d[j];
            for (int j = 0; j < i; ++j) { // rank-2 update of the leading block
                f = d[j];
                g = e[j];
                for (int k = j; k < i; ++k) Z(k, j) -= f * e[k] + g * d[k];
                d[j] = Z(i - 1, j);
                Z(i, j) = 0.0;
            }
        }
        d[i] = h;
    }
    // ---- accumulate the reflectors into Z
    for (int i = 0; i < n - 1; ++i) {
        Z(n - 1, i) = Z(i, i);
        Z(i, i) = 1.0;
        const double h = d[i + 1];
        if (h != 0.0) {
            const double invh = 1.0 / h;
            for (int k = 0; k <= i; ++k) d[k] = Z(k, i + 1) * invh;
            for (int j = 0; j <= i; ++j) {
                double g = 0.0;
                for (int k = 0; k <= i; ++k) g += Z(k, i + 1) * Z(k, j);
                for (int k = 0; k <= i; ++k) Z(k, j) -= g * d[k];
            }
        }
        for (int k = 0; k <= i; ++k) Z(k, i + 1) = 0.0;
    }
    for (int j = 0; j < n; ++j) { d[j] = Z(n - 1, j); Z(n - 1, j) = 0.0; }
    Z(n - 1, n - 1) = 1.0;
    // ---- implicit-shift QL on (d, e)
    for (int i = 1; i < n; ++i) e[i - 1] = e[i];
    e[n - 1] = 0.0;
    double f = 0.0, tst1 = 0.0;
    const double eps = 2.220446049250313e-16;
    bool ok = true;
    for (int l = 0; l < n; ++l) {
        tst1 = fmax(tst1, fabs(d[l]) + fabs(e[l]));
        int m = l;
        while (m < n - 1 && fabs(e[m]) > eps * tst1) ++m; // e[n-1] = 0 ends the search
        if (m > l) {
            int iter = 0;
            do {
                if (++iter > 40) { ok = false; break; }
                double g = d[l];
                double p = (d[l + 1] - g) / (2.0 * e[l]);
                double r = sqrt(p * p + 1.0);
                if (p < 0.0) r = -r;
                d[l] = e[l] / (p + r);
                d[l + 1] = e[l] * (p + r);
                const double dl1 = d[l + 1];
                double h = g - d[l];
                for (int i = l + 2; i < n; ++i) d[i] -= h;
                f += h;
                p = d[m];
                double c = 1.0, c2 = 1.0, c3 = 1.0, s = 0.0, s2 = 0.0;
                const double el1 = e[l + 1];
                for (int i = m - 1; i >= l; --i) {
                    c3 = c2; c2 = c; s2 = s;
                    g = c * e[i];
                    h = c * p;
                    // r = hypot(p, e[i]) without overflow concerns: the entries are O(||A||)
                    const double ei = e[i];
                    const double big = fmax(fabs(p), fabs(ei));
                    if (big == 0.0) r = 0.0;
                    else { const double a = p / big, b = ei / big; r = big * sqrt(a * a + b * b); }
                    e[i + 1] = s * r;
                    const double invr = (r != 0.0) ? 1.0 / r : 0.0;
                    s = ei * invr;
                    c = p * invr;
                    p = c * d[i] - s * g;
                    d[i + 1] = h + s * (c * g + s * d[i]);
                    for (int k = 0; k < n; ++k) {
                        const double zk1 = Z(k, i + 1), zk0 = Z(k, i);
                        Z(k, i + 1) = s * zk0 + c * zk1;
                        Z(k, i) = c * zk0 - s * zk1;
                    }
                }
                p = -s * s2 * c3 * el1 * e[l] / dl1;
                e[l] = s * p;
                d[l] = c * p;
            } while (fabs(e[l]) > eps * tst1);
        }
        d[l] += f;
        e[l] = 0.0;
    }
    return ok;
}

#ifndef CIPC_DENSE_QL
#define CIPC_DENSE_QL 1
#endif
#if defined(__CUDACC__)
// Dense fallback used on the device for stencils without a low-rank fast path (mollified stencils, stencils
// outside the barrier's support).  Every block annihilates rigid translations (H t = 0 for t = (c,c,..,c)), so it
// is first compressed with an orthonormal Helmert basis Q of the translation-free subspace:  M = Q H Q^T is
// (N-3) x (N-3), H+ = Q^T M+ Q.  The cyclic Jacobi iteration then runs on M with matrix and eigenvectors held in
// SHARED memory (element e of thread t at sm[e*BD + t]: bank-conflict free, 41 KB per 64 threads -> four CTAs per SM,
// register-limited), instead of thread-local arrays that spill to L1/L2.  H: N x N row-major (N = 3*nb) in local memory, replaced by its PSD projection.
__device__ inline void psd_project_reduced_smem(double* H, int nb, double* sm, int tid, int BD)
{
    const int N = 3 * nb, K = N - 3;
    double A[45];                   // upper triangle of the symmetric K x K matrix: thread-local (written once, read a few times)
    double* V = sm;                 // K*K matrix / eigenvectors: the array the eigen-solver works on stays in shared memory
    auto tri = [&](int i, int j) { const int a = i < j ? i : j, b = i < j ? j : i; return a * (2 * K - a + 1) / 2 + (b - a); };
#define SA(i, j) A[tri((i), (j))]
#define SV(i, j) V[((i) * K + (j)) * BD + tid]
    // Helmert rows h_k (k = 1..nb-1): k entries 1/sqrt(k(k+1)), then -k/sqrt(k(k+1)), then zeros
    double hc[3], hd[3];
    for (int k = 1; k < nb; ++k) { const double s = 1.0 / sqrt((double)(k * (k + 1))); hc[k - 1] = s; hd[k - 1] = -k * s; }
    auto helm = [&](int k, int I) { return I < k + 1 ? hc[k] : (I == k + 1 ? hd[k] : 0.0); }; // k = 0..nb-2
    double fro = 0.0;
    for (int k = 0; k < nb - 1; ++k)
        for (int a = 0; a < 3; ++a)
            for (int l = 0; l < nb - 1; ++l)
                for (int b = 0; b < 3; ++b) {
                    const int r = 3 * k + a, c = 3 * l + b;
                    SV(r, c) = (r == c) ? 1.0 : 0.0;
                    if (r > c) continue; // symmetric: only the upper triangle is formed and stored
                    double s = 0.0;
                    for (int I = 0; I <= k + 1; ++I) {
                        const double hI = helm(k, I);
                        for (int J = 0; J <= l + 1; ++J) s += hI * helm(l, J) * H[(3 * I + a) * N + 3 * J + b];
                    }
                    SA(r, c) = s;
                    fro += (r == c ? 1.0 : 2.0) * s * s;
                }
    double lam[9];
    bool solved = false;
#if CIPC_DENSE_QL
    if (fro > 0.0) {
        // eigen-decomposition by tridiagonalisation + implicit QL (sym_eig_ql) on a full copy of M held in V; the Jacobi
        // iteration below stays as the fallback for a QL run that does not converge (never observed)
        for (int i = 0; i < K; ++i)
            for (int j = 0; j < K; ++j) SV(i, j) = SA(i, j);
        double ev[9], sub[9];
        solved = sym_eig_ql(K, [&](int i, int j) -> double& { return V[(i * K + j) * BD + tid]; }, ev, sub);
        if (solved) for (int i = 0; i < K; ++i) lam[i] = ev[i] > 0.0 ? ev[i] : 0.0;
        else
            for (int i = 0; i < K; ++i)
                for (int j = 0; j < K; ++j) SV(i, j) = (i == j) ? 1.0 : 0.0;
    }
#endif
    if (fro > 0.0 && !solved) {
        const double tol = 1e-25 * fro; // off-diagonal norm <= 3e-13 ||M||: far below the 1e-9 parity gate
        const double iscale = rsqrt(fro);
        for (int sweep = 0; sweep < 14; ++sweep) {
            double off = 0.0;
            for (int p = 0; p < K; ++p)
                for (int q = p + 1; q < K; ++q) { const double v = SA(p, q); off += v * v; }
            if (off <= tol) break;
            for (int p = 0; p < K - 1; ++p)
                for (int q = p + 1; q < K; ++q) {
                    const double apq = SA(p, q);
                    if (apq * apq <= 1e-36 * fro) continue;
                    const double app = SA(p, p), aqq = SA(q, q);
                    // rotation angle in fp32 on scaled entries, rotation orthogonal to fp64 rounding, exact similarity
                    // transform (a_pq is updated, not zeroed): see hess4_factor
                    const float df = (float)((aqq - app) * iscale), af = (float)(apq * iscale) * 2.0f;
                    const float h2 = fmaf(df, df, af * af);
                    const double t = (double)__fdividef(df >= 0.0f ? af : -af, fabsf(df) + h2 * rsqrtf(h2));
                    const double c = rsqrt(t * t + 1.0), s = t * c;
                    for (int k = 0; k < K; ++k) {
                        if (k != p && k != q) {
                            const double akp = SA(k, p), akq = SA(k, q);
                            const double np_ = c * akp - s * akq, nq_ = s * akp + c * akq;
                            SA(k, p) = np_;
                            SA(k, q) = nq_;
                        }
                        const double vkp = SV(k, p), vkq = SV(k, q);
                        SV(k, p) = c * vkp - s * vkq;
                        SV(k, q) = s * vkp + c * vkq;
                    }
                    const double cc = c * c, ss = s * s, cs = c * s;
                    SA(p, p) = cc * app - 2.0 * cs * apq + ss * aqq;
                    SA(q, q) = ss * app + 2.0 * cs * apq + cc * aqq;
                    SA(p, q) = cs * (app - aqq) + (cc - ss) * apq;
                }
        }
    }
    // M+ = sum_{lambda>0} lambda v v^T, kept in A (upper part recomputed fully)
    if (!solved) for (int i = 0; i < K; ++i) { const double l = SA(i, i); lam[i] = l > 0.0 ? l : 0.0; }
    for (int i = 0; i < K; ++i)
        for (int j = i; j < K; ++j) {
            double s = 0.0;
            for (int k = 0; k < K; ++k) s += lam[k] * SV(i, k) * SV(j, k);
            SA(i, j) = s;
        }
    // H+ = Q^T M+ Q
    for (int I = 0; I < nb; ++I)
        for (int a = 0; a < 3; ++a)
            for (int J = 0; J < nb; ++J)
                for (int b = 0; b < 3; ++b) {
                    double s = 0.0;
                    for (int k = (I > 0 ? I - 1 : 0); k < nb - 1; ++k) {
                        const double hI = helm(k, I);
                        for (int l = (J > 0 ? J - 1 : 0); l < nb - 1; ++l) s += hI * helm(l, J) * SA(3 * k + a, 3 * l + b);
                    }
                    H[(3 * I + a) * N + 3 * J + b] = s;
                }
#undef SA
#undef SV
}
#endif

} // namespace cipc
