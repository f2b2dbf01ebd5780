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

} // namespace cipc
