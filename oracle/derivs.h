// TEST INFRASTRUCTURE -- CPU oracle (see geom.h header).
//
// derivs.h: gradients and Hessians of the squared distances and of the edge-edge cross-norm
// mollifier.  The reference ships MATLAB-generated straight-line code for these
// (Math/Distance/POINT_EDGE.h:60-113,266-587 g_PE3D/H_PE3D; POINT_TRIANGLE.h:24-86,103-549
// g_PT/H_PT; EDGE_EDGE.h:24-104,121-732 g_EE/H_EE; EDGE_EDGE_MOLLIFIER.h:20-79,97-366
// g_EECN2/H_EECN2).  Those are analytic derivatives of the closed forms in geom.h; the tolerance
// on them is 1e-9 relative, so the oracle evaluates the same derivatives from a compact
// vector-calculus derivation instead of restating ~2500 generated lines:
//
//   PT / EE :  f = N^2 / D,  N = det[w,u,v] = w.(u x v),  D = |u x v|^2
//              PT: w = p-t0, u = t1-t0, v = t2-t0      EE: w = eb0-ea0, u = ea1-ea0, v = eb1-eb0
//   PE      :  f = |w|^2 - (w.u)^2/|u|^2,  w = p-e0, u = e1-e0
//   cross^2 :  c = D(u,v),  u = ea1-ea0, v = eb1-eb0
//
// Derivatives are formed in the difference variables y = (w,u,v) and pulled back to the vertex
// variables x with the constant +-1 map y = A x (H_x = A^T H_y A).  oracle/_ref (the reference's own
// generated code compiled against a stub Eigen) and finite differences pin these in tests/.
#pragma once
#include "geom.h"

namespace cipc_oracle {

// 3x3 helpers on row-major double[9]
static inline void skew(const V3& a, double* S) // S b = a x b
{
    S[0] = 0; S[1] = -a.z; S[2] = a.y;
    S[3] = a.z; S[4] = 0; S[5] = -a.x;
    S[6] = -a.y; S[7] = a.x; S[8] = 0;
}
static inline void outer(const V3& a, const V3& b, double* M)
{
    M[0] = a.x * b.x; M[1] = a.x * b.y; M[2] = a.x * b.z;
    M[3] = a.y * b.x; M[4] = a.y * b.y; M[5] = a.y * b.z;
    M[6] = a.z * b.x; M[7] = a.z * b.y; M[8] = a.z * b.z;
}

// y-space derivatives of D(u,v) = |u x v|^2 = |u|^2|v|^2 - (u.v)^2
struct DDeriv {
    double D;
    V3 Du, Dv;
    double Duu[9], Dvv[9], Duv[9]; // Duv = d(Du)/dv
};
static inline void d_derivs(const V3& u, const V3& v, DDeriv& o, bool hess)
{
    const double uu = norm2(u), vv = norm2(v), uv = dot(u, v);
    o.D = norm2(cross(u, v));
    o.Du = 2.0 * (vv * u - uv * v);
    o.Dv = 2.0 * (uu * v - uv * u);
    if (!hess) return;
    double vvT[9], uuT[9], uvT[9], vuT[9];
    outer(v, v, vvT); outer(u, u, uuT); outer(u, v, uvT); outer(v, u, vuT);
    for (int i = 0; i < 9; ++i) {
        const double I = (i == 0 || i == 4 || i == 8) ? 1.0 : 0.0;
        o.Duu[i] = 2.0 * (vv * I - vvT[i]);
        o.Dvv[i] = 2.0 * (uu * I - uuT[i]);
        o.Duv[i] = 4.0 * uvT[i] - 2.0 * uv * I - 2.0 * vuT[i];
    }
}

// f(w,u,v) = det[w,u,v]^2 / |u x v|^2 : gradient gy[9] (w,u,v) and Hessian Hy[81] row-major
static inline void wuv_derivs(const V3& w, const V3& u, const V3& v, double* gy, double* Hy)
{
    const V3 n = cross(u, v);
    const double N = dot(w, n);
    DDeriv dd;
    d_derivs(u, v, dd, Hy != nullptr);
    const double D = dd.D;
    const V3 Nw = n, Nu = cross(v, w), Nv = cross(w, u);
    const double gN[9] = {Nw.x, Nw.y, Nw.z, Nu.x, Nu.y, Nu.z, Nv.x, Nv.y, Nv.z};
    const double gD[9] = {0, 0, 0, dd.Du.x, dd.Du.y, dd.Du.z, dd.Dv.x, dd.Dv.y, dd.Dv.z};
    const double c1 = 2.0 * N / D, c2 = N * N / (D * D);
    for (int i = 0; i < 9; ++i) gy[i] = c1 * gN[i] - c2 * gD[i];
    if (!Hy) return;
    // second derivatives of N (multilinear: diagonal blocks vanish)
    double HN[81] = {0}, HD[81] = {0};
    double Su[9], Sv[9], Sw[9];
    skew(u, Su); skew(v, Sv); skew(w, Sw);
    auto setblk = [](double* H, int bi, int bj, const double* B, double s) {
        for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) H[(3 * bi + r) * 9 + 3 * bj + c] = s * B[3 * r + c];
    };
    setblk(HN, 0, 1, Sv, -1.0); setblk(HN, 0, 2, Su, 1.0);   // d(Nw)/du = -[v]x, d(Nw)/dv = [u]x
    setblk(HN, 1, 0, Sv, 1.0);  setblk(HN, 1, 2, Sw, -1.0);  // d(Nu)/dw = [v]x,  d(Nu)/dv = -[w]x
    setblk(HN, 2, 0, Su, -1.0); setblk(HN, 2, 1, Sw, 1.0);   // d(Nv)/dw = -[u]x, d(Nv)/du = [w]x
    double DvuT[9];
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) DvuT[3 * r + c] = dd.Duv[3 * c + r];
    setblk(HD, 1, 1, dd.Duu, 1.0); setblk(HD, 1, 2, dd.Duv, 1.0);
    setblk(HD, 2, 1, DvuT, 1.0);   setblk(HD, 2, 2, dd.Dvv, 1.0);
    const double a = 2.0 / D, b = 2.0 * N / D, c = 2.0 * N / (D * D), d = N * N / (D * D), e = 2.0 * N * N / (D * D * D);
    for (int i = 0; i < 9; ++i)
        for (int j = 0; j < 9; ++j)
            Hy[i * 9 + j] = a * gN[i] * gN[j] + b * HN[i * 9 + j] - c * (gN[i] * gD[j] + gD[i] * gN[j])
                - d * HD[i * 9 + j] + e * gD[i] * gD[j];
}

// pull back: x has nx 3-blocks, y has ny 3-blocks, C[a*nx + I] in {-1,0,1}
static inline void pull_grad(int ny, int nx, const int* C, const double* gy, double* gx)
{
    for (int I = 0; I < nx; ++I)
        for (int r = 0; r < 3; ++r) {
            double s = 0;
            for (int a = 0; a < ny; ++a) s += C[a * nx + I] * gy[3 * a + r];
            gx[3 * I + r] = s;
        }
}
static inline void pull_hess(int ny, int nx, const int* C, const double* Hy, double* Hx)
{
    const int Ny = 3 * ny, Nx = 3 * nx;
    for (int I = 0; I < nx; ++I)
        for (int J = 0; J < nx; ++J)
            for (int r = 0; r < 3; ++r)
                for (int c = 0; c < 3; ++c) {
                    double s = 0;
                    for (int a = 0; a < ny; ++a) {
                        if (!C[a * nx + I]) continue;
                        for (int b = 0; b < ny; ++b) {
                            if (!C[b * nx + J]) continue;
                            s += C[a * nx + I] * C[b * nx + J] * Hy[(3 * a + r) * Ny + 3 * b + c];
                        }
                    }
                    Hx[(3 * I + r) * Nx + 3 * J + c] = s;
                }
}

// ---- point-point (Math/Distance/POINT_POINT.h:19-41)
static inline void pp_grad(const V3& a, const V3& b, double* g)
{
    const V3 d = 2.0 * (a - b);
    g[0] = d.x; g[1] = d.y; g[2] = d.z; g[3] = -d.x; g[4] = -d.y; g[5] = -d.z;
}
static inline void pp_hess(const V3&, const V3&, double* H)
{
    for (int i = 0; i < 36; ++i) H[i] = 0;
    for (int i = 0; i < 6; ++i) H[i * 6 + i] = 2.0;
    for (int i = 0; i < 3; ++i) { H[i * 6 + i + 3] = -2.0; H[(i + 3) * 6 + i] = -2.0; }
}

// ---- point-edge: x = (p,e0,e1), y = (w,u)
static const int C_PE[2 * 3] = {1, -1, 0, /*u*/ 0, -1, 1};
static inline void pe_wu_derivs(const V3& w, const V3& u, double* gy, double* Hy)
{
    const double c = dot(w, u), L = norm2(u);
    const V3 fw = 2.0 * w - (2.0 * c / L) * u;
    const V3 fu = (2.0 * c * c / (L * L)) * u - (2.0 * c / L) * w;
    gy[0] = fw.x; gy[1] = fw.y; gy[2] = fw.z; gy[3] = fu.x; gy[4] = fu.y; gy[5] = fu.z;
    if (!Hy) return;
    double uu[9], uw[9], wu[9], ww[9];
    outer(u, u, uu); outer(u, w, uw); outer(w, u, wu); outer(w, w, ww);
    for (int r = 0; r < 3; ++r)
        for (int cc = 0; cc < 3; ++cc) {
            const int k = 3 * r + cc;
            const double I = (r == cc) ? 1.0 : 0.0;
            const double Hww = 2.0 * I - 2.0 * uu[k] / L;
            const double Hwu = -2.0 * (uw[k] / L + c * I / L - 2.0 * c * uu[k] / (L * L));
            const double Huu = -2.0 * ww[k] / L + 4.0 * c * (wu[k] + uw[k]) / (L * L) + 2.0 * c * c * I / (L * L)
                - 8.0 * c * c * uu[k] / (L * L * L);
            Hy[(r) * 6 + cc] = Hww;
            Hy[(r) * 6 + 3 + cc] = Hwu;
            Hy[(3 + cc) * 6 + r] = Hwu; // symmetric partner (d/dw of f_u) = Hwu^T
            Hy[(3 + r) * 6 + 3 + cc] = Huu;
        }
}
static inline void pe_grad(const V3& p, const V3& e0, const V3& e1, double* g)
{
    double gy[6];
    pe_wu_derivs(p - e0, e1 - e0, gy, nullptr);
    pull_grad(2, 3, C_PE, gy, g);
}
static inline void pe_hess(const V3& p, const V3& e0, const V3& e1, double* H)
{
    double gy[6], Hy[36];
    pe_wu_derivs(p - e0, e1 - e0, gy, Hy);
    pull_hess(2, 3, C_PE, Hy, H);
}

// ---- point-triangle: x = (p,t0,t1,t2), y = (w,u,v) = (p-t0, t1-t0, t2-t0)
static const int C_PT[3 * 4] = {1, -1, 0, 0, /*u*/ 0, -1, 1, 0, /*v*/ 0, -1, 0, 1};
static inline void pt_grad(const V3& p, const V3& t0, const V3& t1, const V3& t2, double* g)
{
    double gy[9];
    wuv_derivs(p - t0, t1 - t0, t2 - t0, gy, nullptr);
    pull_grad(3, 4, C_PT, gy, g);
}
static inline void pt_hess(const V3& p, const V3& t0, const V3& t1, const V3& t2, double* H)
{
    double gy[9], Hy[81];
    wuv_derivs(p - t0, t1 - t0, t2 - t0, gy, Hy);
    pull_hess(3, 4, C_PT, Hy, H);
}

// ---- edge-edge: x = (ea0,ea1,eb0,eb1), y = (w,u,v) = (eb0-ea0, ea1-ea0, eb1-eb0)
static const int C_EE[3 * 4] = {-1, 0, 1, 0, /*u*/ -1, 1, 0, 0, /*v*/ 0, 0, -1, 1};
static inline void ee_grad(const V3& a0, const V3& a1, const V3& b0, const V3& b1, double* g)
{
    double gy[9];
    wuv_derivs(b0 - a0, a1 - a0, b1 - b0, gy, nullptr);
    pull_grad(3, 4, C_EE, gy, g);
}
static inline void ee_hess(const V3& a0, const V3& a1, const V3& b0, const V3& b1, double* H)
{
    double gy[9], Hy[81];
    wuv_derivs(b0 - a0, a1 - a0, b1 - b0, gy, Hy);
    pull_hess(3, 4, C_EE, Hy, H);
}

// ---- edge-edge cross-norm^2: y = (u,v)
static const int C_CN[2 * 4] = {-1, 1, 0, 0, /*v*/ 0, 0, -1, 1};
static inline void eecn2_grad(const V3& a0, const V3& a1, const V3& b0, const V3& b1, double* g)
{
    DDeriv dd;
    d_derivs(a1 - a0, b1 - b0, dd, false);
    const double gy[6] = {dd.Du.x, dd.Du.y, dd.Du.z, dd.Dv.x, dd.Dv.y, dd.Dv.z};
    pull_grad(2, 4, C_CN, gy, g);
}
static inline void eecn2_hess(const V3& a0, const V3& a1, const V3& b0, const V3& b1, double* H)
{
    DDeriv dd;
    d_derivs(a1 - a0, b1 - b0, dd, true);
    double Hy[36];
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) {
            Hy[r * 6 + c] = dd.Duu[3 * r + c];
            Hy[r * 6 + 3 + c] = dd.Duv[3 * r + c];
            Hy[(3 + c) * 6 + r] = dd.Duv[3 * r + c];
            Hy[(3 + r) * 6 + 3 + c] = dd.Dvv[3 * r + c];
        }
    pull_hess(2, 4, C_CN, Hy, H);
}

// ---- mollifier e(x) and derivatives (Math/Distance/EDGE_EDGE_MOLLIFIER.h:461-524)
static inline double ee_mollifier(const V3& a0, const V3& a1, const V3& b0, const V3& b1, double eps_x)
{
    const double s = ee_cross_norm2(a0, a1, b0, b1);
    return (s < eps_x) ? eem(s, eps_x) : 1.0;
}
static inline void ee_mollifier_grad(const V3& a0, const V3& a1, const V3& b0, const V3& b1, double eps_x, double* g)
{
    const double s = ee_cross_norm2(a0, a1, b0, b1);
    if (s < eps_x) {
        const double q = eem_g(s, eps_x);
        eecn2_grad(a0, a1, b0, b1, g);
        for (int i = 0; i < 12; ++i) g[i] *= q;
    }
    else for (int i = 0; i < 12; ++i) g[i] = 0;
}
static inline void ee_mollifier_hess(const V3& a0, const V3& a1, const V3& b0, const V3& b1, double eps_x, double* H)
{
    const double s = ee_cross_norm2(a0, a1, b0, b1);
    if (s < eps_x) {
        const double qg = eem_g(s, eps_x), qH = eem_H(s, eps_x);
        double g[12];
        eecn2_grad(a0, a1, b0, b1, g);
        eecn2_hess(a0, a1, b0, b1, H);
        for (int i = 0; i < 12; ++i)
            for (int j = 0; j < 12; ++j) H[i * 12 + j] = H[i * 12 + j] * qg + (qH * g[i]) * g[j];
    }
    else for (int i = 0; i < 144; ++i) H[i] = 0;
}

} // namespace cipc_oracle
