// Host-side probe of codim-ipc_b200/csrc/geom.cuh (compiled by tests with g++ -ffp-contract=off):
// lets the CPU test-suite compare the product's geometry code with the oracle without a GPU.
#define CIPC_HOST_TEST 1
#include "../../codim-ipc_b200/csrc/geom.cuh"
#include "../../codim-ipc_b200/csrc/hess.cuh"
#include "../../codim-ipc_b200/csrc/eig.cuh"
using namespace cipc;
static xv3 X(const double* p) { return xv3(xd(p[0]), xd(p[1]), xd(p[2])); }
static dv3 Dv(const double* p) { return dv3(p[0], p[1], p[2]); }
extern "C" {
// kind: 0 PP, 1 PE, 2 PT, 3 EE, 4 cross-norm^2
void probe_dist_derivs(int kind, const double* x, double* d, double* g, double* H)
{
    const dv3 v[4] = {Dv(x), Dv(x + 3), Dv(x + 6), Dv(x + 9)};
    const int rows[3] = {0, 1, 2};
    const int n = (kind == 0) ? 6 : (kind == 1 ? 9 : 12);
    for (int i = 0; i < n * n; ++i) H[i] = 0;
    switch (kind) {
    case 0: *d = pp_dist2(X(x), X(x + 3)); pp_derivs(v[0], v[1], g, H, 6, rows, 1.0); break;
    case 1: *d = pe_dist2(X(x), X(x + 3), X(x + 6)); pe_derivs(v[0], v[1], v[2], g, H, 9, rows, 1.0); break;
    case 2: *d = pt_dist2(X(x), X(x + 3), X(x + 6), X(x + 9)); d4_derivs(false, v, g, H, 1.0); break;
    case 3: *d = ee_dist2(X(x), X(x + 3), X(x + 6), X(x + 9)); d4_derivs(true, v, g, H, 1.0); break;
    default: *d = ee_cross_norm2(X(x), X(x + 3), X(x + 6), X(x + 9)); eecn2_derivs(v, g, H, 1.0); break;
    }
}
int probe_type(int kind, const double* x)
{
    if (kind == 1) return pe_type(X(x), X(x + 3), X(x + 6));
    if (kind == 2) return pt_type(X(x), X(x + 3), X(x + 6), X(x + 9));
    return ee_type(X(x), X(x + 3), X(x + 6), X(x + 9));
}
double probe_dist2_unclassified(int kind, const double* x)
{
    if (kind == 1) return pe_dist2_unclassified(X(x), X(x + 3), X(x + 6));
    if (kind == 2) return pt_dist2_unclassified(X(x), X(x + 3), X(x + 6), X(x + 9));
    return ee_dist2_unclassified(X(x), X(x + 3), X(x + 6), X(x + 9));
}
int probe_accd(int kind, const double* x, const double* dx, double eta, double thickness, double* toc)
{
    xd t;
    bool ok;
    const xd bound(*toc);
    switch (kind) {
    case 0: ok = pp_accd(X(x), X(x + 3), X(dx), X(dx + 3), eta, thickness, bound, t, nullptr); break;
    case 1: ok = pe_accd(X(x), X(x + 3), X(x + 6), X(dx), X(dx + 3), X(dx + 6), eta, thickness, bound, t, nullptr); break;
    case 2: ok = pt_accd(X(x), X(x + 3), X(x + 6), X(x + 9), X(dx), X(dx + 3), X(dx + 6), X(dx + 9), eta, thickness, bound, t, nullptr); break;
    default: ok = ee_accd(X(x), X(x + 3), X(x + 6), X(x + 9), X(dx), X(dx + 3), X(dx + 6), X(dx + 9), eta, thickness, bound, t, nullptr); break;
    }
    *toc = t.v;
    return ok;
}
int probe_broadphase(int kind, const double* x, const double* dx, double dist)
{
    switch (kind) {
    case 0: return pt_cd_broadphase(X(x), X(x + 3), X(x + 6), X(x + 9), dist);
    case 1: return ee_cd_broadphase(X(x), X(x + 3), X(x + 6), X(x + 9), dist);
    case 2: return pe_cd_broadphase(X(x), X(x + 3), X(x + 6), dist);
    case 3: return pt_ccd_broadphase(X(x), X(x + 3), X(x + 6), X(x + 9), X(dx), X(dx + 3), X(dx + 6), X(dx + 9), dist);
    case 4: return ee_ccd_broadphase(X(x), X(x + 3), X(x + 6), X(x + 9), X(dx), X(dx + 3), X(dx + 6), X(dx + 9), dist);
    case 5: return pe_ccd_broadphase(X(x), X(x + 3), X(x + 6), X(dx), X(dx + 3), X(dx + 6), dist);
    default: return pp_ccd_broadphase(X(x), X(x + 3), X(dx), X(dx + 3), dist);
    }
}
// low-rank Hessian paths of hess.cuh: kind 0 PP, 1 PE, 2 PT, 3 EE;  H = alpha g g^T + beta K (projected if project)
void probe_hess_lowrank(int kind, const double* x, double alpha, double beta, int project, double* H)
{
    const dv3 v[4] = {Dv(x), Dv(x + 3), Dv(x + 6), Dv(x + 9)};
    const int n = (kind == 0) ? 6 : (kind == 1 ? 9 : 12);
    auto emit = [&](int I, int J, const double* B) {
        for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) H[(3 * I + a) * n + 3 * J + b] = B[3 * a + b];
    };
    if (kind == 0) hess_pp_closed(v[0], v[1], alpha, beta, project != 0, emit);
    else if (kind == 1) hess_pe_lowrank(v[0], v[1], v[2], alpha, beta, project != 0, emit);
    else hess4_lowrank(kind == 3, v, alpha, beta, project != 0, emit);
}
// factored form (two-phase kernel): H+ = sum_k y_k y_k^T
void probe_hess_factor(int kind, const double* x, double alpha, double beta, double* H)
{
    const dv3 v[4] = {Dv(x), Dv(x + 3), Dv(x + 6), Dv(x + 9)};
    const int n = (kind == 0) ? 6 : (kind == 1 ? 9 : 12), ny = (kind == 0) ? 1 : (kind == 1 ? 2 : 3);
    double Y[36];
    if (kind == 0) hess_pp_factor(v[0], v[1], alpha, beta, Y);
    else if (kind == 1) hess_pe_factor(v[0], v[1], v[2], alpha, beta, Y);
    else hess4_factor(kind == 3, v, alpha, beta, Y);
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) {
            double s = 0;
            for (int k = 0; k < ny; ++k) s += Y[k * n + i] * Y[k * n + j];
            H[i * n + j] = s;
        }
}
void probe_psd_jacobi(int n, double* H)
{
    if (n == 6) psd_project_jacobi<6>(H);
    else if (n == 9) psd_project_jacobi<9>(H);
    else psd_project_jacobi<12>(H);
}
// Z: n x n row-major symmetric in, eigenvectors (columns) out; d: eigenvalues
int probe_sym_eig_ql(int n, double* Z, double* d)
{
    double e[16];
    return sym_eig_ql(n, [&](int i, int j) -> double& { return Z[i * n + j]; }, d, e) ? 1 : 0;
}
void probe_barrier(int elastic, double d, double dHat, double k0, double* out3)
{
    out3[0] = barrier_b(elastic, d, dHat, k0); out3[1] = barrier_g(elastic, d, dHat, k0); out3[2] = barrier_H(elastic, d, dHat, k0);
}
}
