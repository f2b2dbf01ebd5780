// TEST INFRASTRUCTURE -- CPU oracle for the C-IPC contact hot path.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
// use anything under oracle/.  The product path (codim-ipc_b200/csrc) never includes this.
//
// geom.h: scalar restatement of the reference's closed-form squared distances, closest-feature
// classifiers, AABB broad-phase tests and barrier functions.  Operation ORDER follows the
// reference expression by expression (Eigen fixed-size 3-vector ops evaluate component 0,1,2
// and reduce left to right), because constraint-set membership and the ACCD iterates depend on
// the exact rounding of these values.  Compile with -ffp-contract=off.
//
// All citations are relative to /root/reference/Library.
#pragma once
#include <cmath>
#include <algorithm>
#include <limits>

namespace cipc_oracle {

struct V3 {
    double x, y, z;
    V3() : x(0), y(0), z(0) {}
    V3(double a, double b, double c) : x(a), y(b), z(c) {}
    explicit V3(const double* p) : x(p[0]), y(p[1]), z(p[2]) {}
    double operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
};
static inline V3 operator+(const V3& a, const V3& b) { return V3(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline V3 operator-(const V3& a, const V3& b) { return V3(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline V3 operator*(double s, const V3& a) { return V3(s * a.x, s * a.y, s * a.z); }
static inline V3 operator/(const V3& a, double s) { return V3(a.x / s, a.y / s, a.z / s); }
static inline double dot(const V3& a, const V3& b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
static inline double norm2(const V3& a) { return (a.x * a.x + a.y * a.y) + a.z * a.z; }
static inline V3 cross(const V3& a, const V3& b)
{
    return V3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
static inline V3 vmin(const V3& a, const V3& b) { return V3(std::min(a.x, b.x), std::min(a.y, b.y), std::min(a.z, b.z)); }
static inline V3 vmax(const V3& a, const V3& b) { return V3(std::max(a.x, b.x), std::max(a.y, b.y), std::max(a.z, b.z)); }

// ---------------------------------------------------------------- squared distances
// Math/Distance/POINT_POINT.h:11-17
static inline double pp_dist2(const V3& a, const V3& b) { return norm2(a - b); }
// Math/Distance/POINT_EDGE.h:11-25 (3-D branch)
static inline double pe_dist2(const V3& p, const V3& e0, const V3& e1)
{
    return norm2(cross(e0 - p, e1 - p)) / norm2(e1 - e0);
}
// Math/Distance/POINT_TRIANGLE.h:11-22
static inline double pt_dist2(const V3& p, const V3& t0, const V3& t1, const V3& t2)
{
    const V3 b = cross(t1 - t0, t2 - t0);
    const double aTb = dot(p - t0, b);
    return aTb * aTb / norm2(b);
}
// Math/Distance/EDGE_EDGE.h:11-22
static inline double ee_dist2(const V3& ea0, const V3& ea1, const V3& eb0, const V3& eb1)
{
    const V3 b = cross(ea1 - ea0, eb1 - eb0);
    const double aTb = dot(eb0 - ea0, b);
    return aTb * aTb / norm2(b);
}
// Math/Distance/EDGE_EDGE_MOLLIFIER.h:9-18
static inline double ee_cross_norm2(const V3& ea0, const V3& ea1, const V3& eb0, const V3& eb1)
{
    return norm2(cross(ea1 - ea0, eb1 - eb0));
}
// Math/Distance/EDGE_EDGE_MOLLIFIER.h:582-592
static inline double ee_mollifier_threshold(const V3& a0r, const V3& a1r, const V3& b0r, const V3& b1r)
{
    return 1.0e-3 * norm2(a0r - a1r) * norm2(b0r - b1r);
}

// ---------------------------------------------------------------- 2x2 pivoted LDLT solve
// Restates what Eigen 3.3's LDLT<Matrix2d>::compute + solve do for the call
// `(basis * basis.transpose()).ldlt().solve(rhs)` at Math/Distance/DISTANCE_TYPE.h:46,54,62.
// Eigen is not vendored in the reference (CMakeLists.txt:37 find_package), so this is a
// restatement of its published algorithm: symmetric pivot on the largest |diagonal| (first max),
// unit-lower L, D; solve = P^T L^-T D^-1 L^-1 P b with D entries below 1/highest() zeroed.
static inline void ldlt2_solve(double a00, double a10, double a11, double b0, double b1,
    double& x0, double& x1)
{
    bool swap = !(std::fabs(a00) >= std::fabs(a11)); // maxCoeff returns the first maximum
    if (swap) { std::swap(a00, a11); std::swap(b0, b1); }
    double d0 = a00, l10 = a10, d1 = a11;
    if (std::fabs(d0) > 0.0) {
        l10 = a10 / d0;
        const double temp = d0 * l10;
        d1 = a11 - l10 * temp;
    }
    else {
        // whole diagonal is zero: Eigen stops factorising; L stays identity-like with l10 = a10
        // only if a10 == 0 is it a valid factorisation.  Degenerate triangle, never hit by tests.
        l10 = 0.0;
    }
    // forward substitution (unit lower)
    double y0 = b0;
    double y1 = b1 - l10 * y0;
    // D^-1 with Eigen's tolerance
    const double tol = 1.0 / std::numeric_limits<double>::max();
    y0 = (std::fabs(d0) > tol) ? y0 / d0 : 0.0;
    y1 = (std::fabs(d1) > tol) ? y1 / d1 : 0.0;
    // backward substitution with L^T
    y0 = y0 - l10 * y1;
    if (swap) { x0 = y1; x1 = y0; }
    else { x0 = y0; x1 = y1; }
}

// ---------------------------------------------------------------- closest-feature classifiers
// Math/Distance/DISTANCE_TYPE.h:12-28
static inline int pe_type(const V3& p, const V3& e0, const V3& e1, double& ratio)
{
    const V3 e = e1 - e0;
    ratio = dot(e, p - e0) / norm2(e);
    if (ratio < 0) return 0;
    if (ratio > 1) return 1;
    return 2;
}

// one of the three edge tests of DISTANCE_TYPE.h:40-62
static inline void pt_edge_param(const V3& r0, const V3& nVec, const V3& rel, double& s, double& t)
{
    const V3 r1 = cross(r0, nVec);
    const double m00 = dot(r0, r0), m01 = dot(r0, r1), m10 = dot(r1, r0), m11 = dot(r1, r1);
    (void)m01; // LDLT reads the lower triangle only
    const double b0 = dot(r0, rel), b1 = dot(r1, rel);
    ldlt2_solve(m00, m10, m11, b0, b1, s, t);
}

// Math/Distance/DISTANCE_TYPE.h:30-81
static inline int pt_type(const V3& p, const V3& t0, const V3& t1, const V3& t2)
{
    const V3 nVec = cross(t1 - t0, t2 - t0);
    double p00, p10, p01, p11, p02, p12;
    pt_edge_param(t1 - t0, nVec, p - t0, p00, p10);
    if (p00 > 0.0 && p00 < 1.0 && p10 >= 0.0) return 3;
    pt_edge_param(t2 - t1, nVec, p - t1, p01, p11);
    if (p01 > 0.0 && p01 < 1.0 && p11 >= 0.0) return 4;
    pt_edge_param(t0 - t2, nVec, p - t2, p02, p12);
    if (p02 > 0.0 && p02 < 1.0 && p12 >= 0.0) return 5;
    if (p00 <= 0.0 && p02 >= 1.0) return 0;
    if (p01 <= 0.0 && p00 >= 1.0) return 1;
    if (p02 <= 0.0 && p01 >= 1.0) return 2;
    return 6;
}

// Math/Distance/DISTANCE_TYPE.h:84-163
static inline int ee_type(const V3& ea0, const V3& ea1, const V3& eb0, const V3& eb1)
{
    const V3 u = ea1 - ea0, v = eb1 - eb0, w = ea0 - eb0;
    const double a = norm2(u), b = dot(u, v), c = norm2(v), d = dot(u, w), e = dot(v, w);
    const double D = a * c - b * b;
    double tD = D, sN, tN;
    int defaultCase = 8;
    sN = (b * e - c * d);
    if (sN <= 0.0) { tN = e; tD = c; defaultCase = 2; }
    else if (sN >= D) { tN = e + b; tD = c; defaultCase = 5; }
    else {
        tN = (a * e - b * d);
        if (tN > 0.0 && tN < tD) {
            const V3 uxv = cross(u, v);
            if (dot(uxv, w) == 0.0 || norm2(uxv) < 1.0e-20 * a * c) {
                if (sN < D / 2) { tN = e; tD = c; defaultCase = 2; }
                else { tN = e + b; tD = c; defaultCase = 5; }
            }
        }
    }
    if (tN <= 0.0) {
        if (-d <= 0.0) return 0;
        else if (-d >= a) return 3;
        else return 6;
    }
    else if (tN >= tD) {
        if ((-d + b) <= 0.0) return 1;
        else if ((-d + b) >= a) return 4;
        else return 7;
    }
    return defaultCase;
}

// ---------------------------------------------------------------- unclassified distances
// Math/Distance/DISTANCE_UNCLASSIFIED.h:15-59
static inline double pt_dist2_unclassified(const V3& p, const V3& t0, const V3& t1, const V3& t2)
{
    switch (pt_type(p, t0, t1, t2)) {
    case 0: return pp_dist2(p, t0);
    case 1: return pp_dist2(p, t1);
    case 2: return pp_dist2(p, t2);
    case 3: return pe_dist2(p, t0, t1);
    case 4: return pe_dist2(p, t1, t2);
    case 5: return pe_dist2(p, t2, t0);
    default: return pt_dist2(p, t0, t1, t2);
    }
}
// Math/Distance/DISTANCE_UNCLASSIFIED.h:61-120
static inline double ee_dist2_unclassified(const V3& ea0, const V3& ea1, const V3& eb0, const V3& eb1)
{
    switch (ee_type(ea0, ea1, eb0, eb1)) {
    case 0: return pp_dist2(ea0, eb0);
    case 1: return pp_dist2(ea0, eb1);
    case 2: return pe_dist2(ea0, eb0, eb1);
    case 3: return pp_dist2(ea1, eb0);
    case 4: return pp_dist2(ea1, eb1);
    case 5: return pe_dist2(ea1, eb0, eb1);
    case 6: return pe_dist2(eb0, ea0, ea1);
    case 7: return pe_dist2(eb1, ea0, ea1);
    default: return ee_dist2(ea0, ea1, eb0, eb1);
    }
}
// Math/Distance/DISTANCE_UNCLASSIFIED.h:122-147
static inline double pe_dist2_unclassified(const V3& p, const V3& e0, const V3& e1)
{
    const V3 v = e1 - e0, w = p - e0;
    const double c1 = dot(w, v);
    if (c1 <= 0.0) return pp_dist2(p, e0);
    const double c2 = norm2(v);
    if (c2 <= c1) return pp_dist2(p, e1);
    const double b = c1 / c2;
    return pp_dist2(p, e0 + b * v);
}

// ---------------------------------------------------------------- AABB tests
// Math/Distance/CCD.h:15-29 (3-D instantiation)
static inline bool pe_cd_broadphase(const V3& x0, const V3& x1, const V3& x2, double dist)
{
    const V3 mx = vmax(x1, x2), mn = vmin(x1, x2);
    return !((x0.x - mx.x > dist) || (x0.y - mx.y > dist) || (x0.z - mx.z > dist) ||
             (mn.x - x0.x > dist) || (mn.y - x0.y > dist) || (mn.z - x0.z > dist));
}
// Math/Distance/CCD.h:149-165
static inline bool pt_cd_broadphase(const V3& p, const V3& t0, const V3& t1, const V3& t2, double dist)
{
    const V3 mx = vmax(vmax(t0, t1), t2), mn = vmin(vmin(t0, t1), t2);
    return !((p.x - mx.x > dist) || (p.y - mx.y > dist) || (p.z - mx.z > dist) ||
             (mn.x - p.x > dist) || (mn.y - p.y > dist) || (mn.z - p.z > dist));
}
static inline bool box_gap_ok(const V3& mna, const V3& mxa, const V3& mnb, const V3& mxb, double dist)
{
    return !((mna.x - mxb.x > dist) || (mna.y - mxb.y > dist) || (mna.z - mxb.z > dist) ||
             (mnb.x - mxa.x > dist) || (mnb.y - mxa.y > dist) || (mnb.z - mxa.z > dist));
}
// Math/Distance/CCD.h:167-185
static inline bool ee_cd_broadphase(const V3& a0, const V3& a1, const V3& b0, const V3& b1, double dist)
{
    return box_gap_ok(vmin(a0, a1), vmax(a0, a1), vmin(b0, b1), vmax(b0, b1), dist);
}
// Math/Distance/CCD.h:187-211
static inline bool pt_ccd_broadphase(const V3& p, const V3& t0, const V3& t1, const V3& t2,
    const V3& dp, const V3& dt0, const V3& dt1, const V3& dt2, double dist)
{
    const V3 pe = p + dp;
    const V3 mxp = vmax(p, pe), mnp = vmin(p, pe);
    const V3 t0e = t0 + dt0, t1e = t1 + dt1, t2e = t2 + dt2;
    const V3 mxt = vmax(vmax(vmax(vmax(vmax(t0, t1), t2), t0e), t1e), t2e);
    const V3 mnt = vmin(vmin(vmin(vmin(vmin(t0, t1), t2), t0e), t1e), t2e);
    return box_gap_ok(mnp, mxp, mnt, mxt, dist);
}
// Math/Distance/CCD.h:213-235
static inline bool ee_ccd_broadphase(const V3& a0, const V3& a1, const V3& b0, const V3& b1,
    const V3& da0, const V3& da1, const V3& db0, const V3& db1, double dist)
{
    const V3 a0e = a0 + da0, a1e = a1 + da1, b0e = b0 + db0, b1e = b1 + db1;
    const V3 mxa = vmax(vmax(vmax(a0, a1), a0e), a1e), mna = vmin(vmin(vmin(a0, a1), a0e), a1e);
    const V3 mxb = vmax(vmax(vmax(b0, b1), b0e), b1e), mnb = vmin(vmin(vmin(b0, b1), b0e), b1e);
    return box_gap_ok(mna, mxa, mnb, mxb, dist);
}
// Math/Distance/CCD.h:237-257
static inline bool pe_ccd_broadphase(const V3& p, const V3& e0, const V3& e1,
    const V3& dp, const V3& de0, const V3& de1, double dist)
{
    const V3 pe = p + dp, e0e = e0 + de0, e1e = e1 + de1;
    const V3 mxp = vmax(p, pe), mnp = vmin(p, pe);
    const V3 mxe = vmax(vmax(vmax(e0, e1), e0e), e1e), mne = vmin(vmin(vmin(e0, e1), e0e), e1e);
    return box_gap_ok(mnp, mxp, mne, mxe, dist);
}
// Math/Distance/CCD.h:259-277
static inline bool pp_ccd_broadphase(const V3& p0, const V3& p1, const V3& dp0, const V3& dp1, double dist)
{
    const V3 p0e = p0 + dp0, p1e = p1 + dp1;
    return box_gap_ok(vmin(p0, p0e), vmax(p0, p0e), vmin(p1, p1e), vmax(p1, p1e), dist);
}

// ---------------------------------------------------------------- barrier (non-elastic and elastic)
// Math/BARRIER.h:9-23
static inline double barrier(bool elastic, double d, double dHat, const double* kappa)
{
    if (!elastic) return -kappa[0] * (d - dHat) * (d - dHat) * std::log(d / dHat);
    return -kappa[0] * std::pow(d / dHat - 1, 2) * std::log(d / dHat);
}
// Math/BARRIER.h:25-43
static inline double barrier_gradient(bool elastic, double d1, double dHat1, const double* kappa)
{
    if (!elastic) {
        const double t2 = d1 - dHat1;
        return kappa[0] * (t2 * std::log(d1 / dHat1) * -2.0 - (t2 * t2) / d1);
    }
    const double one_over_dHat = 1 / dHat1;
    const double t2 = d1 * one_over_dHat - 1;
    return kappa[0] * (t2 * one_over_dHat * std::log(d1 * one_over_dHat) * -2.0 - (t2 * t2) / d1);
}
// Math/BARRIER.h:45-62
static inline double barrier_hessian(bool elastic, double d1, double dHat1, const double* kappa)
{
    const double t2 = d1 - dHat1;
    const double H = kappa[0] * ((std::log(d1 / dHat1) * -2.0 - t2 * 4.0 / d1) + 1.0 / (d1 * d1) * (t2 * t2));
    return elastic ? H / (dHat1 * dHat1) : H;
}

// ---------------------------------------------------------------- mollifier scalar parts
// Math/Distance/EDGE_EDGE_MOLLIFIER.h:440-459
static inline double eem(double x, double eps) { const double r = x / eps; return (-r + 2.0) * r; }
static inline double eem_g(double x, double eps) { const double o = 1.0 / eps; return 2.0 * o * (-o * x + 1.0); }
static inline double eem_H(double, double eps) { return -2.0 / (eps * eps); }

} // namespace cipc_oracle
