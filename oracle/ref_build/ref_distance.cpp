// TEST INFRASTRUCTURE: builds oracle/_ref/libcipc_refdist.so -- the reference's OWN
// closed-form distance, classifier, derivative (MATLAB-generated), mollifier, barrier and ACCD
// code (Library/Math/Distance/*.h, Library/Math/BARRIER.h), included from where it lies under
// /root/reference and compiled against the stub Eigen in ref_build/stub.  Used by tests/ to pin
// oracle/geom.h + oracle/derivs.h (and through them the CUDA path) to the reference.
// No reference source is copied: this file only #includes it and exports C probes.
#include <Eigen/Core>
#include <cmath>
using std::log;
#include <Math/Distance/DISTANCE_TYPE.h>
#include <Math/Distance/POINT_POINT.h>
#include <Math/Distance/POINT_EDGE.h>
#include <Math/Distance/POINT_TRIANGLE.h>
#include <Math/Distance/EDGE_EDGE.h>
#include <Math/Distance/EDGE_EDGE_MOLLIFIER.h>
#include <Math/Distance/CCD.h>
#include <Math/BARRIER.h>
#include <FEM/FRICTION_UTILS.h>

using namespace JGSL;
typedef Eigen::Matrix<double, 3, 1> V3;

// symmetric Hessians are written through .data(); read them back the same way
extern "C" {

void ref_dist_derivs(int kind, const double* x, double* d, double* g, double* H)
{
    const V3 a(x), b(x + 3), c(x + 6), e(x + 9);
    switch (kind) {
    case 0: {
        Eigen::Matrix<double, 6, 1> G; Eigen::Matrix<double, 6, 6> HH;
        Point_Point_Distance(a, b, *d); Point_Point_Distance_Gradient(a, b, G); Point_Point_Distance_Hessian(a, b, HH);
        for (int i = 0; i < 6; ++i) g[i] = G[i];
        for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) H[i * 6 + j] = HH(i, j);
        break; }
    case 1: {
        Eigen::Matrix<double, 9, 1> G; Eigen::Matrix<double, 9, 9> HH;
        Point_Edge_Distance(a, b, c, *d); Point_Edge_Distance_Gradient(a, b, c, G); Point_Edge_Distance_Hessian(a, b, c, HH);
        for (int i = 0; i < 9; ++i) g[i] = G[i];
        for (int i = 0; i < 9; ++i) for (int j = 0; j < 9; ++j) H[i * 9 + j] = HH(i, j);
        break; }
    case 2: {
        Eigen::Matrix<double, 12, 1> G; Eigen::Matrix<double, 12, 12> HH;
        Point_Triangle_Distance(a, b, c, e, *d); Point_Triangle_Distance_Gradient(a, b, c, e, G); Point_Triangle_Distance_Hessian(a, b, c, e, HH);
        for (int i = 0; i < 12; ++i) g[i] = G[i];
        for (int i = 0; i < 12; ++i) for (int j = 0; j < 12; ++j) H[i * 12 + j] = HH(i, j);
        break; }
    case 3: {
        Eigen::Matrix<double, 12, 1> G; Eigen::Matrix<double, 12, 12> HH;
        Edge_Edge_Distance(a, b, c, e, *d); Edge_Edge_Distance_Gradient(a, b, c, e, G); Edge_Edge_Distance_Hessian(a, b, c, e, HH);
        for (int i = 0; i < 12; ++i) g[i] = G[i];
        for (int i = 0; i < 12; ++i) for (int j = 0; j < 12; ++j) H[i * 12 + j] = HH(i, j);
        break; }
    default: {
        Eigen::Matrix<double, 12, 1> G; Eigen::Matrix<double, 12, 12> HH;
        Edge_Edge_Cross_Norm2(a, b, c, e, *d); Edge_Edge_Cross_Norm2_Gradient(a, b, c, e, G); Edge_Edge_Cross_Norm2_Hessian(a, b, c, e, HH);
        for (int i = 0; i < 12; ++i) g[i] = G[i];
        for (int i = 0; i < 12; ++i) for (int j = 0; j < 12; ++j) H[i * 12 + j] = HH(i, j);
        break; }
    }
}
void ref_mollifier(const double* x, double eps_x, double* e, double* g, double* H)
{
    const V3 a(x), b(x + 3), c(x + 6), d(x + 9);
    Eigen::Matrix<double, 12, 1> G; Eigen::Matrix<double, 12, 12> HH;
    Edge_Edge_Mollifier(a, b, c, d, eps_x, *e);
    Edge_Edge_Mollifier_Gradient(a, b, c, d, eps_x, G);
    Edge_Edge_Mollifier_Hessian(a, b, c, d, eps_x, HH);
    for (int i = 0; i < 12; ++i) g[i] = G[i];
    for (int i = 0; i < 12; ++i) for (int j = 0; j < 12; ++j) H[i * 12 + j] = HH(i, j);
}
double ref_mollifier_threshold(const double* x0)
{
    double eps_x;
    Edge_Edge_Mollifier_Threshold(V3(x0), V3(x0 + 3), V3(x0 + 6), V3(x0 + 9), eps_x);
    return eps_x;
}
int ref_pt_type(const double* x) { return Point_Triangle_Distance_Type(V3(x), V3(x + 3), V3(x + 6), V3(x + 9)); }
int ref_ee_type(const double* x) { return Edge_Edge_Distance_Type(V3(x), V3(x + 3), V3(x + 6), V3(x + 9)); }
int ref_pe_type(const double* x) { double r; return Point_Edge_Distance_Type(V3(x), V3(x + 3), V3(x + 6), r); }
double ref_dist2_unclassified(int kind, const double* x)
{
    double d = 0;
    if (kind == 1) Point_Edge_Distance_Unclassified(V3(x), V3(x + 3), V3(x + 6), d);
    else if (kind == 2) Point_Triangle_Distance_Unclassified(V3(x), V3(x + 3), V3(x + 6), V3(x + 9), d);
    else Edge_Edge_Distance_Unclassified(V3(x), V3(x + 3), V3(x + 6), V3(x + 9), d);
    return d;
}
int ref_accd(int kind, const double* x, const double* dx, double eta, double thickness, double* toc)
{
    const V3 a(x), b(x + 3), c(x + 6), d(x + 9), da(dx), db(dx + 3), dc(dx + 6), dd(dx + 9);
    switch (kind) {
    case 0: return Point_Point_CCD(a, b, da, db, eta, thickness, *toc);
    case 1: return Point_Edge_CCD(a, b, c, da, db, dc, eta, thickness, *toc);
    case 2: return Point_Triangle_CCD(a, b, c, d, da, db, dc, dd, eta, thickness, *toc);
    default: return Edge_Edge_CCD(a, b, c, d, da, db, dc, dd, eta, thickness, *toc);
    }
}
// kind: 0 PT-CD, 1 EE-CD, 2 PE-CD, 3 PT-CCD, 4 EE-CCD, 5 PE-CCD, 6 PP-CCD
int ref_broadphase(int kind, const double* x, const double* dx, double dist)
{
    const V3 a(x), b(x + 3), c(x + 6), d(x + 9), da(dx), db(dx + 3), dc(dx + 6), dd(dx + 9);
    switch (kind) {
    case 0: return Point_Triangle_CD_Broadphase(a, b, c, d, dist);
    case 1: return Edge_Edge_CD_Broadphase(a, b, c, d, dist);
    case 2: return Point_Edge_CD_Broadphase(a, b, c, dist);
    case 3: return Point_Triangle_CCD_Broadphase(a, b, c, d, da, db, dc, dd, dist);
    case 4: return Edge_Edge_CCD_Broadphase(a, b, c, d, da, db, dc, dd, dist);
    case 5: return Point_Edge_CCD_Broadphase(a, b, c, da, db, dc, dist);
    default: return Point_Point_CCD_Broadphase(a, b, da, db, dist);
    }
}
// FEM/FRICTION_UTILS.h probes: kind 0 PP, 1 PE, 2 PT, 3 EE -> tangent basis (column-major 3x2), closest point, TT (2 x 3nb row-major)
void ref_friction_utils(int kind, const double* x, double* basis, double* closest, double* TTout)
{
    const V3 a(x), b(x + 3), c(x + 6), d(x + 9);
    Eigen::Matrix<double, 3, 2> B;
    Eigen::Matrix<double, 2, 1> cp; cp[0] = cp[1] = 0;
    if (kind == 0) {
        Point_Point_Tangent_Basis(a, b, B);
        Eigen::Matrix<double, 2, 6> TT; Point_Point_TT(B, TT);
        for (int i = 0; i < 2; ++i) for (int j = 0; j < 6; ++j) TTout[i * 6 + j] = TT(i, j);
    }
    else if (kind == 1) {
        Point_Edge_Tangent_Basis(a, b, c, B); Point_Edge_Closest_Point(a, b, c, cp[0]);
        Eigen::Matrix<double, 2, 9> TT; Point_Edge_TT(B, cp[0], TT);
        for (int i = 0; i < 2; ++i) for (int j = 0; j < 9; ++j) TTout[i * 9 + j] = TT(i, j);
    }
    else if (kind == 2) {
        Point_Triangle_Tangent_Basis(a, b, c, d, B); Point_Triangle_Closest_Point(a, b, c, d, cp);
        Eigen::Matrix<double, 2, 12> TT; Point_Triangle_TT(B, cp[0], cp[1], TT);
        for (int i = 0; i < 2; ++i) for (int j = 0; j < 12; ++j) TTout[i * 12 + j] = TT(i, j);
    }
    else {
        Edge_Edge_Tangent_Basis(a, b, c, d, B); Edge_Edge_Closest_Point(a, b, c, d, cp);
        Eigen::Matrix<double, 2, 12> TT; Edge_Edge_TT(B, cp[0], cp[1], TT);
        for (int i = 0; i < 2; ++i) for (int j = 0; j < 12; ++j) TTout[i * 12 + j] = TT(i, j);
    }
    for (int j = 0; j < 2; ++j) for (int i = 0; i < 3; ++i) basis[3 * j + i] = B(i, j);
    closest[0] = cp[0]; closest[1] = cp[1];
}
void ref_friction_f(double x2, double epsvh, double* out3)
{
    f0_SF(x2, epsvh, out3[0]); f1_SF_Div_RelDXNorm(x2, epsvh, out3[1]); f2_SF_Term(x2, epsvh, out3[2]);
}
void ref_barrier_fn(int elastic, double d, double dHat, const double* kappa_in, double* out3)
{
    double kappa[3] = {kappa_in[0], kappa_in[1], kappa_in[2]};
    if (elastic) { Barrier<true>(d, dHat, kappa, out3[0]); Barrier_Gradient<true>(d, dHat, kappa, out3[1]); Barrier_Hessian<true>(d, dHat, kappa, out3[2]); }
    else { Barrier<false>(d, dHat, kappa, out3[0]); Barrier_Gradient<false>(d, dHat, kappa, out3[1]); Barrier_Hessian<false>(d, dHat, kappa, out3[2]); }
}

} // extern "C"
