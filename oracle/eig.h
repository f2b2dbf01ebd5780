// TEST INFRASTRUCTURE -- CPU oracle (see geom.h header).
//
// eig.h: dense symmetric eigen-decomposition and the PSD projection `makePD`
// (Math/UTILS.h:9-27).  The reference calls Eigen::SelfAdjointEigenSolver (Eigen is not vendored:
// CMakeLists.txt:37), whose published algorithm is Householder tridiagonalisation followed by
// implicit-shift QL/QR iteration with eigenvalues sorted ascending.  This file restates that
// algorithm (classical tred2/tql2 formulation) for n <= 12.  V * max(L,0) * V^T is a
// well-conditioned function of the input, so any backward-stable solver agrees to ~1e-14 ||H||;
// the GPU path uses a different solver (Jacobi), which makes this an independent check.
#pragma once
#include <cmath>
#include <algorithm>

namespace cipc_oracle {

// A: n x n row-major symmetric (only read).  V: eigenvectors as columns (row-major n x n),
// d: eigenvalues ascending.
// NT > 0 fixes the size at compile time (the reference's Eigen solver is a fixed-size instantiation too, so the compiler
// unrolls / vectorises its loops; a run-time n would under-state the reference's CPU speed in oracle/_ref timings)
template <int NT>
static inline void sym_eig_t(int n_, const double* A, double* V, double* d)
{
    const int n = NT > 0 ? NT : n_;
    double e[12];
    for (int i = 0; i < n * n; ++i) V[i] = A[i];
    // ---- Householder reduction to tridiagonal form (accumulating the transform in V)
    for (int j = 0; j < n; ++j) d[j] = V[(n - 1) * n + j];
    for (int i = n - 1; i > 0; --i) {
        double scale = 0.0, h = 0.0;
        for (int k = 0; k < i; ++k) scale += std::fabs(d[k]);
        if (scale == 0.0) {
            e[i] = d[i - 1];
            for (int j = 0; j < i; ++j) {
                d[j] = V[(i - 1) * n + j];
                V[i * n + j] = 0.0;
                V[j * n + i] = 0.0;
            }
        }
        else {
            for (int k = 0; k < i; ++k) { d[k] /= scale; h += d[k] * d[k]; }
            double f = d[i - 1];
            double g = std::sqrt(h);
            if (f > 0) g = -g;
            e[i] = scale * g;
            h -= f * g;
            d[i - 1] = f - g;
            for (int j = 0; j < i; ++j) e[j] = 0.0;
            for (int j = 0; j < i; ++j) {
                f = d[j];
                V[j * n + i] = f;
                g = e[j] + V[j * n + j] * f;
                for (int k = j + 1; k <= i - 1; ++k) {
                    g += V[k * n + j] * d[k];
                    e[k] += V[k * n + j] * f;
                }
                e[j] = g;
            }
            f = 0.0;
            for (int j = 0; j < i; ++j) { e[j] /= h; f += e[j] * d[j]; }
            const double hh = f / (h + h);
            for (int j = 0; j < i; ++j) e[j] -= hh * d[j];
            for (int j = 0; j < i; ++j) {
                f = d[j];
                g = e[j];
                for (int k = j; k <= i - 1; ++k) V[k * n + j] -= (f * e[k] + g * d[k]);
                d[j] = V[(i - 1) * n + j];
                V[i * n + j] = 0.0;
            }
        }
        d[i] = h;
    }
    for (int i = 0; i < n - 1; ++i) {
        V[(n - 1) * n + i] = V[i * n + i];
        V[i * n + i] = 1.0;
        const double h = d[i + 1];
        if (h != 0.0) {
            for (int k = 0; k <= i; ++k) d[k] = V[k * n + i + 1] / h;
            for (int j = 0; j <= i; ++j) {
                double g = 0.0;
                for (int k = 0; k <= i; ++k) g += V[k * n + i + 1] * V[k * n + j];
                for (int k = 0; k <= i; ++k) V[k * n + j] -= g * d[k];
            }
        }
        for (int k = 0; k <= i; ++k) V[k * n + i + 1] = 0.0;
    }
    for (int j = 0; j < n; ++j) { d[j] = V[(n - 1) * n + j]; V[(n - 1) * n + j] = 0.0; }
    V[(n - 1) * n + n - 1] = 1.0;
    e[0] = 0.0;
    // ---- implicit QL on the tridiagonal (d, e)
    for (int i = 1; i < n; ++i) e[i - 1] = e[i];
    e[n - 1] = 0.0;
    double f = 0.0, tst1 = 0.0;
    const double eps = 2.220446049250313e-16;
    for (int l = 0; l < n; ++l) {
        tst1 = std::max(tst1, std::fabs(d[l]) + std::fabs(e[l]));
        int m = l;
        while (m < n) {
            if (std::fabs(e[m]) <= eps * tst1) break;
            ++m;
        }
        if (m > l) {
            int iter = 0;
            do {
                ++iter;
                double g = d[l];
                double p = (d[l + 1] - g) / (2.0 * e[l]);
                double r = std::hypot(p, 1.0);
                if (p < 0) r = -r;
                d[l] = e[l] / (p + r);
                d[l + 1] = e[l] * (p + r);
                const double dl1 = d[l + 1];
                double h = g - d[l];
                for (int i = l + 2; i < n; ++i) d[i] -= h;
                f += h;
                p = d[m];
                double c = 1.0, c2 = c, c3 = c;
                const double el1 = e[l + 1];
                double s = 0.0, s2 = 0.0;
                for (int i = m - 1; i >= l; --i) {
                    c3 = c2;
                    c2 = c;
                    s2 = s;
                    g = c * e[i];
                    h = c * p;
                    r = std::hypot(p, e[i]);
                    e[i + 1] = s * r;
                    s = e[i] / r;
                    c = p / r;
                    p = c * d[i] - s * g;
                    d[i + 1] = h + s * (c * g + s * d[i]);
                    for (int k = 0; k < n; ++k) {
                        h = V[k * n + i + 1];
                        V[k * n + i + 1] = s * V[k * n + i] + c * h;
                        V[k * n + i] = c * V[k * n + i] - s * h;
                    }
                }
                p = -s * s2 * c3 * el1 * e[l] / dl1;
                e[l] = s * p;
                d[l] = c * p;
            } while (std::fabs(e[l]) > eps * tst1 && iter < 60);
        }
        d[l] = d[l] + f;
        e[l] = 0.0;
    }
    // ---- sort ascending
    for (int i = 0; i < n - 1; ++i) {
        int k = i;
        double p = d[i];
        for (int j = i + 1; j < n; ++j) if (d[j] < p) { k = j; p = d[j]; }
        if (k != i) {
            d[k] = d[i]; d[i] = p;
            for (int j = 0; j < n; ++j) std::swap(V[j * n + i], V[j * n + k]);
        }
    }
}

static inline void sym_eig(int n, const double* A, double* V, double* d)
{
    switch (n) {
    case 12: sym_eig_t<12>(n, A, V, d); break;
    case 9: sym_eig_t<9>(n, A, V, d); break;
    case 6: sym_eig_t<6>(n, A, V, d); break;
    default: sym_eig_t<0>(n, A, V, d); break;
    }
}

// Math/UTILS.h:9-27  makePD: leave untouched when the smallest eigenvalue is >= 0, otherwise
// zero the negative prefix of the ascending spectrum and rebuild V D V^T.
static inline void make_pd(int n, double* H)
{
    double V[144], d[12];
    sym_eig(n, H, V, d);
    if (d[0] >= 0) return;
    for (int i = 0; i < n; ++i) {
        if (d[i] < 0) d[i] = 0;
        else break;
    }
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) {
            double s = 0;
            for (int k = 0; k < n; ++k) s += V[i * n + k] * d[k] * V[j * n + k];
            H[i * n + j] = s;
        }
}

} // namespace cipc_oracle
