// geom.cuh -- per-pair geometry of the C-IPC contact path for sm_100a (fp64).
//
// Two families of functions live here:
//
//  (1) EXACT functions (suffix-free names taking `xv3`): squared distances, closest-feature
//      classifiers, AABB tests and the ACCD iterates.  Their results feed predicates
//      (d < dHat^2, toc comparisons), so they evaluate the reference's expressions in the
//      reference's order with every multiply/add rounded separately (`xd` wraps
//      __dmul_rn/__dadd_rn so that nvcc never contracts them into FMAs).  Reference:
//      Library/Math/Distance/{POINT_POINT,POINT_EDGE,POINT_TRIANGLE,EDGE_EDGE}.h distance heads,
//      DISTANCE_TYPE.h:12-163, DISTANCE_UNCLASSIFIED.h:15-148, CCD.h:149-277.
//
//  (2) DERIVATIVE functions (plain double, FMA allowed): gradients / Hessians of the squared
//      distances and of the edge-edge mollifier.  The reference ships MATLAB-generated code for
//      these (POINT_TRIANGLE.h:24-549, EDGE_EDGE.h:24-732, POINT_EDGE.h:60-587,
//      EDGE_EDGE_MOLLIFIER.h:20-366); here they are evaluated from a compact derivation in the
//      difference variables y=(w,u,v):  f = det[w,u,v]^2 / |u x v|^2  (PT and EE share it),
//      f = |w|^2 - (w.u)^2/|u|^2 (PE), c = |u x v|^2 (mollifier), pulled back to the vertices by
//      the constant +-1 map.  Tolerance on these is 1e-9 relative (north star), measured ~1e-13.
//
// The file is also compiled for the host by tests (CIPC_HOST_TEST) to check it against the oracle
// without a GPU; the product only ever uses the device instantiation.
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define CIPC_HD __host__ __device__ __forceinline__
#else
#define CIPC_HD inline
#endif

namespace cipc {

// ------------------------------------------------------------------ exact scalar
struct xd {
    double v;
    CIPC_HD xd() : v(0) {}
    CIPC_HD xd(double a) : v(a) {}
    CIPC_HD operator double() const { return v; }
};
#if defined(__CUDA_ARCH__)
CIPC_HD xd operator+(xd a, xd b) { return xd(__dadd_rn(a.v, b.v)); }
CIPC_HD xd operator-(xd a, xd b) { return xd(__dsub_rn(a.v, b.v)); }
CIPC_HD xd operator*(xd a, xd b) { return xd(__dmul_rn(a.v, b.v)); }
CIPC_HD xd operator/(xd a, xd b) { return xd(__ddiv_rn(a.v, b.v)); }
#else
CIPC_HD xd operator+(xd a, xd b) { return xd(a.v + b.v); }
CIPC_HD xd operator-(xd a, xd b) { return xd(a.v - b.v); }
CIPC_HD xd operator*(xd a, xd b) { return xd(a.v * b.v); }
CIPC_HD xd operator/(xd a, xd b) { return xd(a.v / b.v); }
#endif
CIPC_HD xd operator-(xd a) { return xd(-a.v); }
CIPC_HD bool operator<(xd a, xd b) { return a.v < b.v; }
CIPC_HD bool operator>(xd a, xd b) { return a.v > b.v; }
CIPC_HD bool operator<=(xd a, xd b) { return a.v <= b.v; }
CIPC_HD bool operator>=(xd a, xd b) { return a.v >= b.v; }
CIPC_HD bool operator==(xd a, xd b) { return a.v == b.v; }
CIPC_HD xd xsqrt(xd a) { return xd(sqrt(a.v)); }
CIPC_HD xd xmin(xd a, xd b) { return a.v < b.v ? a : b; }
CIPC_HD xd xmax(xd a, xd b) { return a.v > b.v ? a : b; }
CIPC_HD xd xabs(xd a) { return xd(fabs(a.v)); }
CIPC_HD double cipc_rsqrt(double x)
{
#if defined(__CUDA_ARCH__)
    return rsqrt(x);
#else
    return 1.0 / sqrt(x);
#endif
}

template <class S>
struct vec3 {
    S x, y, z;
    CIPC_HD vec3() {}
    CIPC_HD vec3(S a, S b, S c) : x(a), y(b), z(c) {}
};
typedef vec3<xd> xv3;     // exact
typedef vec3<double> dv3; // free

template <class S> CIPC_HD vec3<S> operator+(const vec3<S>& a, const vec3<S>& b) { return vec3<S>(a.x + b.x, a.y + b.y, a.z + b.z); }
template <class S> CIPC_HD vec3<S> operator-(const vec3<S>& a, const vec3<S>& b) { return vec3<S>(a.x - b.x, a.y - b.y, a.z - b.z); }
template <class S> CIPC_HD vec3<S> operator*(S s, const vec3<S>& a) { return vec3<S>(s * a.x, s * a.y, s * a.z); }
template <class S> CIPC_HD vec3<S> operator/(const vec3<S>& a, S s) { return vec3<S>(a.x / s, a.y / s, a.z / s); }
template <class S> CIPC_HD S dot(const vec3<S>& a, const vec3<S>& b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
template <class S> CIPC_HD S norm2(const vec3<S>& a) { return (a.x * a.x + a.y * a.y) + a.z * a.z; }
template <class S> CIPC_HD vec3<S> cross(const vec3<S>& a, const vec3<S>& b)
{
    return vec3<S>(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
CIPC_HD xv3 vmin(const xv3& a, const xv3& b) { return xv3(xmin(a.x, b.x), xmin(a.y, b.y), xmin(a.z, b.z)); }
CIPC_HD xv3 vmax(const xv3& a, const xv3& b) { return xv3(xmax(a.x, b.x), xmax(a.y, b.y), xmax(a.z, b.z)); }
CIPC_HD dv3 to_d(const xv3& a) { return dv3(a.x.v, a.y.v, a.z.v); }

// ------------------------------------------------------------------ exact squared distances
CIPC_HD xd pp_dist2(const xv3& a, const xv3& b) { return norm2(a - b); }                     // POINT_POINT.h:11-17
CIPC_HD xd pe_dist2(const xv3& p, const xv3& e0, const xv3& e1)                              // POINT_EDGE.h:22-24
{
    return norm2(cross(e0 - p, e1 - p)) / norm2(e1 - e0);
}
CIPC_HD xd pt_dist2(const xv3& p, const xv3& t0, const xv3& t1, const xv3& t2)               // POINT_TRIANGLE.h:11-22
{
    const xv3 b = cross(t1 - t0, t2 - t0);
    const xd aTb = dot(p - t0, b);
    return aTb * aTb / norm2(b);
}
CIPC_HD xd ee_dist2(const xv3& a0, const xv3& a1, const xv3& b0, const xv3& b1)              // EDGE_EDGE.h:11-22
{
    const xv3 b = cross(a1 - a0, b1 - b0);
    const xd aTb = dot(b0 - a0, b);
    return aTb * aTb / norm2(b);
}
CIPC_HD xd ee_cross_norm2(const xv3& a0, const xv3& a1, const xv3& b0, const xv3& b1)        // EDGE_EDGE_MOLLIFIER.h:9-18
{
    return norm2(cross(a1 - a0, b1 - b0));
}
CIPC_HD xd ee_mollifier_threshold(const xv3& a0r, const xv3& a1r, const xv3& b0r, const xv3& b1r) // EDGE_EDGE_MOLLIFIER.h:582-592
{
    return xd(1.0e-3) * norm2(a0r - a1r) * norm2(b0r - b1r);
}

// ------------------------------------------------------------------ 2x2 pivoted LDLT solve
// What Eigen's `(B B^T).ldlt().solve(rhs)` computes at DISTANCE_TYPE.h:46,54,62 (Eigen is an
// un-vendored dependency of the reference; algorithm restated: pivot on the larger |diagonal|
// (first on ties), unit-lower L, D, solve P^T L^-T D^-1 L^-1 P b, |D_ii| <= 1/DBL_MAX treated as 0).
CIPC_HD void ldlt2_solve(xd a00, xd a10, xd a11, xd b0, xd b1, xd& x0, xd& x1)
{
    const bool swp = !(xabs(a00) >= xabs(a11));
    if (swp) { xd t = a00; a00 = a11; a11 = t; t = b0; b0 = b1; b1 = t; }
    xd d0 = a00, l10 = a10, d1 = a11;
    if (xabs(d0) > xd(0.0)) {
        l10 = a10 / d0;
        const xd temp = d0 * l10;
        d1 = a11 - l10 * temp;
    }
    else l10 = xd(0.0);
    xd y0 = b0;
    xd y1 = b1 - l10 * y0;
    const xd tol(1.0 / 1.7976931348623157e308);
    y0 = (xabs(d0) > tol) ? y0 / d0 : xd(0.0);
    y1 = (xabs(d1) > tol) ? y1 / d1 : xd(0.0);
    y0 = y0 - l10 * y1;
    if (swp) { x0 = y1; x1 = y0; }
    else { x0 = y0; x1 = y1; }
}

// ------------------------------------------------------------------ classifiers
CIPC_HD int pe_type(const xv3& p, const xv3& e0, const xv3& e1)                              // DISTANCE_TYPE.h:12-28
{
    const xv3 e = e1 - e0;
    const xd ratio = dot(e, p - e0) / norm2(e);
    if (ratio < xd(0.0)) return 0;
    if (ratio > xd(1.0)) return 1;
    return 2;
}
CIPC_HD void pt_edge_param(const xv3& r0, const xv3& nVec, const xv3& rel, xd& s, xd& t)
{
    const xv3 r1 = cross(r0, nVec);
    ldlt2_solve(dot(r0, r0), dot(r1, r0), dot(r1, r1), dot(r0, rel), dot(r1, rel), s, t);
}
CIPC_HD int pt_type(const xv3& p, const xv3& t0, const xv3& t1, const xv3& t2)               // DISTANCE_TYPE.h:30-81
{
    const xv3 nVec = cross(t1 - t0, t2 - t0);
    xd p00, p10, p01, p11, p02, p12;
    pt_edge_param(t1 - t0, nVec, p - t0, p00, p10);
    if (p00 > xd(0.0) && p00 < xd(1.0) && p10 >= xd(0.0)) return 3;
    pt_edge_param(t2 - t1, nVec, p - t1, p01, p11);
    if (p01 > xd(0.0) && p01 < xd(1.0) && p11 >= xd(0.0)) return 4;
    pt_edge_param(t0 - t2, nVec, p - t2, p02, p12);
    if (p02 > xd(0.0) && p02 < xd(1.0) && p12 >= xd(0.0)) return 5;
    if (p00 <= xd(0.0) && p02 >= xd(1.0)) return 0;
    if (p01 <= xd(0.0) && p00 >= xd(1.0)) return 1;
    if (p02 <= xd(0.0) && p01 >= xd(1.0)) return 2;
    return 6;
}
CIPC_HD int ee_type(const xv3& ea0, const xv3& ea1, const xv3& eb0, const xv3& eb1)          // DISTANCE_TYPE.h:84-163
{
    const xv3 u = ea1 - ea0, v = eb1 - eb0, w = ea0 - eb0;
    const xd a = norm2(u), b = dot(u, v), c = norm2(v), d = dot(u, w), e = dot(v, w);
    const xd D = a * c - b * b;
    xd tD = D, sN, tN;
    int defaultCase = 8;
    sN = (b * e - c * d);
    if (sN <= xd(0.0)) { tN = e; tD = c; defaultCase = 2; }
    else if (sN >= D) { tN = e + b; tD = c; defaultCase = 5; }
    else {
        tN = (a * e - b * d);
        if (tN > xd(0.0) && tN < tD) {
            const xv3 uxv = cross(u, v);
            if (dot(uxv, w) == xd(0.0) || norm2(uxv) < xd(1.0e-20) * a * c) {
                if (sN < D / xd(2.0)) { tN = e; tD = c; defaultCase = 2; }
                else { tN = e + b; tD = c; defaultCase = 5; }
            }
        }
    }
    if (tN <= xd(0.0)) {
        if (-d <= xd(0.0)) return 0;
        else if (-d >= a) return 3;
        else return 6;
    }
    else if (tN >= tD) {
        if ((-d + b) <= xd(0.0)) return 1;
        else if ((-d + b) >= a) return 4;
        else return 7;
    }
    return defaultCase;
}

// ------------------------------------------------------------------ unclassified distances (ACCD)
CIPC_HD xd pt_dist2_unclassified(const xv3& p, const xv3& t0, const xv3& t1, const xv3& t2)  // DISTANCE_UNCLASSIFIED.h:15-59
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
CIPC_HD xd ee_dist2_unclassified(const xv3& a0, const xv3& a1, const xv3& b0, const xv3& b1) // DISTANCE_UNCLASSIFIED.h:61-120
{
    switch (ee_type(a0, a1, b0, b1)) {
    case 0: return pp_dist2(a0, b0);
    case 1: return pp_dist2(a0, b1);
    case 2: return pe_dist2(a0, b0, b1);
    case 3: return pp_dist2(a1, b0);
    case 4: return pp_dist2(a1, b1);
    case 5: return pe_dist2(a1, b0, b1);
    case 6: return pe_dist2(b0, a0, a1);
    case 7: return pe_dist2(b1, a0, a1);
    default: return ee_dist2(a0, a1, b0, b1);
    }
}
CIPC_HD xd pe_dist2_unclassified(const xv3& p, const xv3& e0, const xv3& e1)                 // DISTANCE_UNCLASSIFIED.h:122-147
{
    const xv3 v = e1 - e0, w = p - e0;
    const xd c1 = dot(w, v);
    if (c1 <= xd(0.0)) return pp_dist2(p, e0);
    const xd c2 = norm2(v);
    if (c2 <= c1) return pp_dist2(p, e1);
    const xd b = c1 / c2;
    return pp_dist2(p, e0 + b * v);
}

// ------------------------------------------------------------------ AABB gap tests (CCD.h:15-29,149-277)
CIPC_HD bool box_gap_ok(const xv3& mna, const xv3& mxa, const xv3& mnb, const xv3& mxb, xd dist)
{
    return !((mna.x - mxb.x > dist) || (mna.y - mxb.y > dist) || (mna.z - mxb.z > dist) ||
             (mnb.x - mxa.x > dist) || (mnb.y - mxa.y > dist) || (mnb.z - mxa.z > dist));
}
CIPC_HD bool pt_cd_broadphase(const xv3& p, const xv3& t0, const xv3& t1, const xv3& t2, xd dist)
{
    return box_gap_ok(p, p, vmin(vmin(t0, t1), t2), vmax(vmax(t0, t1), t2), dist);
}
CIPC_HD bool pe_cd_broadphase(const xv3& p, const xv3& e0, const xv3& e1, xd dist)
{
    return box_gap_ok(p, p, vmin(e0, e1), vmax(e0, e1), dist);
}
CIPC_HD bool ee_cd_broadphase(const xv3& a0, const xv3& a1, const xv3& b0, const xv3& b1, xd dist)
{
    return box_gap_ok(vmin(a0, a1), vmax(a0, a1), vmin(b0, b1), vmax(b0, b1), dist);
}
CIPC_HD bool pt_ccd_broadphase(const xv3& p, const xv3& t0, const xv3& t1, const xv3& t2,
    const xv3& dp, const xv3& dt0, const xv3& dt1, const xv3& dt2, xd dist)
{
    const xv3 pe = p + dp, t0e = t0 + dt0, t1e = t1 + dt1, t2e = t2 + dt2;
    return box_gap_ok(vmin(p, pe), vmax(p, pe),
        vmin(vmin(vmin(vmin(vmin(t0, t1), t2), t0e), t1e), t2e),
        vmax(vmax(vmax(vmax(vmax(t0, t1), t2), t0e), t1e), t2e), dist);
}
CIPC_HD bool ee_ccd_broadphase(const xv3& a0, const xv3& a1, const xv3& b0, const xv3& b1,
    const xv3& da0, const xv3& da1, const xv3& db0, const xv3& db1, xd dist)
{
    const xv3 a0e = a0 + da0, a1e = a1 + da1, b0e = b0 + db0, b1e = b1 + db1;
    return box_gap_ok(vmin(vmin(vmin(a0, a1), a0e), a1e), vmax(vmax(vmax(a0, a1), a0e), a1e),
        vmin(vmin(vmin(b0, b1), b0e), b1e), vmax(vmax(vmax(b0, b1), b0e), b1e), dist);
}
CIPC_HD bool pe_ccd_broadphase(const xv3& p, const xv3& e0, const xv3& e1, const xv3& dp, const xv3& de0, const xv3& de1, xd dist)
{
    const xv3 pe = p + dp, e0e = e0 + de0, e1e = e1 + de1;
    return box_gap_ok(vmin(p, pe), vmax(p, pe), vmin(vmin(vmin(e0, e1), e0e), e1e), vmax(vmax(vmax(e0, e1), e0e), e1e), dist);
}
CIPC_HD bool pp_ccd_broadphase(const xv3& p0, const xv3& p1, const xv3& dp0, const xv3& dp1, xd dist)
{
    const xv3 p0e = p0 + dp0, p1e = p1 + dp1;
    return box_gap_ok(vmin(p0, p0e), vmax(p0, p0e), vmin(p1, p1e), vmax(p1, p1e), dist);
}

// ------------------------------------------------------------------ ACCD (CCD.h:279-483)
// Each returns true when a time of impact toc <= toc_bound was found.  `live_bound` (optional) points
// at the running global minimum: reading it only tightens the bail-out test `toc > bound`, which
// never changes the returned toc of a pair that does return true below the final minimum.
#define CIPC_ACCD_LOOP(ADVANCE, DIST2, FIXUP)                                                       \
    xd dist_cur = xsqrt(dist2_cur);                                                                   \
    const xd gap = eta * dFunc / (dist_cur + thickness);                                              \
    toc = xd(0.0);                                                                                    \
    for (int it = 0; it < 1000000; ++it) {                                                            \
        const xd tocLowerBound = (xd(1.0) - eta) * dFunc / ((dist_cur + thickness) * maxDispMag);     \
        ADVANCE;                                                                                      \
        dist2_cur = DIST2;                                                                            \
        dFunc = dist2_cur - thickness * thickness;                                                    \
        FIXUP;                                                                                        \
        dist_cur = xsqrt(dist2_cur);                                                                  \
        if (toc.v != 0.0 && (dFunc / (dist_cur + thickness) < gap)) return true;                      \
        toc = toc + tocLowerBound;                                                                    \
        if (toc > toc_bound) return false;                                                            \
        if (live_bound && toc.v > *live_bound) return false;                                          \
    }                                                                                                 \
    return false;

CIPC_HD bool pt_accd(xv3 p, xv3 t0, xv3 t1, xv3 t2, xv3 dp, xv3 dt0, xv3 dt1, xv3 dt2, xd eta, xd thickness,
    xd toc_bound, xd& toc, const volatile double* live_bound)
{
    const xv3 mov = (((dt0 + dt1) + dt2) + dp) / xd(4.0);
    dt0 = dt0 - mov; dt1 = dt1 - mov; dt2 = dt2 - mov; dp = dp - mov;
    const xd maxDispMag = xsqrt(norm2(dp)) + xsqrt(xmax(xmax(norm2(dt0), norm2(dt1)), norm2(dt2)));
    if (maxDispMag == xd(0.0)) return false;
    xd dist2_cur = pt_dist2_unclassified(p, t0, t1, t2);
    xd dFunc = dist2_cur - thickness * thickness;
    CIPC_ACCD_LOOP((p = p + tocLowerBound * dp, t0 = t0 + tocLowerBound * dt0, t1 = t1 + tocLowerBound * dt1, t2 = t2 + tocLowerBound * dt2),
        pt_dist2_unclassified(p, t0, t1, t2), (void)0)
}
CIPC_HD xd ee_min_endpoint(const xv3& a0, const xv3& a1, const xv3& b0, const xv3& b1)
{
    return xmin(xmin(norm2(a0 - b0), norm2(a0 - b1)), xmin(norm2(a1 - b0), norm2(a1 - b1)));
}
CIPC_HD bool ee_accd(xv3 a0, xv3 a1, xv3 b0, xv3 b1, xv3 da0, xv3 da1, xv3 db0, xv3 db1, xd eta, xd thickness,
    xd toc_bound, xd& toc, const volatile double* live_bound)
{
    const xv3 mov = (((da0 + da1) + db0) + db1) / xd(4.0);
    da0 = da0 - mov; da1 = da1 - mov; db0 = db0 - mov; db1 = db1 - mov;
    const xd maxDispMag = xsqrt(xmax(norm2(da0), norm2(da1))) + xsqrt(xmax(norm2(db0), norm2(db1)));
    if (maxDispMag == xd(0.0)) return false;
    xd dist2_cur = ee_dist2_unclassified(a0, a1, b0, b1);
    xd dFunc = dist2_cur - thickness * thickness;
    if (dFunc <= xd(0.0)) { dist2_cur = ee_min_endpoint(a0, a1, b0, b1); dFunc = dist2_cur - thickness * thickness; }
    CIPC_ACCD_LOOP((a0 = a0 + tocLowerBound * da0, a1 = a1 + tocLowerBound * da1, b0 = b0 + tocLowerBound * db0, b1 = b1 + tocLowerBound * db1),
        ee_dist2_unclassified(a0, a1, b0, b1),
        if (dFunc <= xd(0.0)) { dist2_cur = ee_min_endpoint(a0, a1, b0, b1); dFunc = dist2_cur - thickness * thickness; })
}
CIPC_HD bool pe_accd(xv3 p, xv3 e0, xv3 e1, xv3 dp, xv3 de0, xv3 de1, xd eta, xd thickness, xd toc_bound, xd& toc,
    const volatile double* live_bound)
{
    const xv3 mov = ((dp + de0) + de1) / xd(3.0);
    de0 = de0 - mov; de1 = de1 - mov; dp = dp - mov;
    const xd maxDispMag = xsqrt(norm2(dp)) + xsqrt(xmax(norm2(de0), norm2(de1)));
    if (maxDispMag == xd(0.0)) return false;
    xd dist2_cur = pe_dist2_unclassified(p, e0, e1);
    xd dFunc = dist2_cur - thickness * thickness;
    CIPC_ACCD_LOOP((p = p + tocLowerBound * dp, e0 = e0 + tocLowerBound * de0, e1 = e1 + tocLowerBound * de1),
        pe_dist2_unclassified(p, e0, e1), (void)0)
}
CIPC_HD bool pp_accd(xv3 p0, xv3 p1, xv3 dp0, xv3 dp1, xd eta, xd thickness, xd toc_bound, xd& toc,
    const volatile double* live_bound)
{
    const xv3 mov = (dp0 + dp1) / xd(2.0);
    dp1 = dp1 - mov; dp0 = dp0 - mov;
    const xd maxDispMag = xsqrt(norm2(dp0)) + xsqrt(norm2(dp1));
    if (maxDispMag == xd(0.0)) return false;
    xd dist2_cur = pp_dist2(p0, p1);
    xd dFunc = dist2_cur - thickness * thickness;
    CIPC_ACCD_LOOP((p0 = p0 + tocLowerBound * dp0, p1 = p1 + tocLowerBound * dp1), pp_dist2(p0, p1), (void)0)
}

// ------------------------------------------------------------------ barrier (Math/BARRIER.h:9-62)
CIPC_HD double barrier_b(bool elastic, double d, double dHat, double k0)
{
    if (!elastic) return -k0 * (d - dHat) * (d - dHat) * log(d / dHat);
    const double r = d / dHat - 1;
    return -k0 * (r * r) * log(d / dHat);
}
CIPC_HD double barrier_g(bool elastic, double d, double dHat, double k0)
{
    if (!elastic) {
        const double t2 = d - dHat;
        return k0 * (t2 * log(d / dHat) * -2.0 - (t2 * t2) / d);
    }
    const double o = 1 / dHat, t2 = d * o - 1;
    return k0 * (t2 * o * log(d * o) * -2.0 - (t2 * t2) / d);
}
CIPC_HD double barrier_H(bool elastic, double d, double dHat, double k0)
{
    const double t2 = d - dHat;
    const double H = k0 * ((log(d / dHat) * -2.0 - t2 * 4.0 / d) + 1.0 / (d * d) * (t2 * t2));
    return elastic ? H / (dHat * dHat) : H;
}

// ------------------------------------------------------------------ derivatives (free arithmetic)
struct m3 { // row-major 3x3
    double a[9];
};
CIPC_HD m3 m3_skew(const dv3& v, double s) // s * [v]x
{
    m3 r;
    r.a[0] = 0; r.a[1] = -s * v.z; r.a[2] = s * v.y;
    r.a[3] = s * v.z; r.a[4] = 0; r.a[5] = -s * v.x;
    r.a[6] = -s * v.y; r.a[7] = s * v.x; r.a[8] = 0;
    return r;
}
CIPC_HD void m3_add_outer(m3& r, double s, const dv3& a, const dv3& b)
{
    r.a[0] += s * a.x * b.x; r.a[1] += s * a.x * b.y; r.a[2] += s * a.x * b.z;
    r.a[3] += s * a.y * b.x; r.a[4] += s * a.y * b.y; r.a[5] += s * a.y * b.z;
    r.a[6] += s * a.z * b.x; r.a[7] += s * a.z * b.y; r.a[8] += s * a.z * b.z;
}
CIPC_HD void m3_add_diag(m3& r, double s) { r.a[0] += s; r.a[4] += s; r.a[8] += s; }
CIPC_HD m3 m3_zero() { m3 r; for (int i = 0; i < 9; ++i) r.a[i] = 0; return r; }

// Writes block (I,J) (and its transpose at (J,I) when I != J) of an n x n row-major matrix, +=.
CIPC_HD void blk_add(double* H, int n, int I, int J, const m3& B, double s)
{
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) {
            const double v = s * B.a[3 * r + c];
            H[(3 * I + r) * n + 3 * J + c] += v;
            if (I != J) H[(3 * J + c) * n + 3 * I + r] += v;
        }
}

// y-space derivative data of D(u,v) = |u x v|^2
struct ddata {
    double D;
    dv3 Du, Dv;
    m3 Duu, Dvv, Duv;
};
CIPC_HD void d_derivs(const dv3& u, const dv3& v, ddata& o, bool hess)
{
    const double uu = norm2(u), vv = norm2(v), uv = dot(u, v);
    o.D = norm2(cross(u, v));
    o.Du = dv3(2.0 * (vv * u.x - uv * v.x), 2.0 * (vv * u.y - uv * v.y), 2.0 * (vv * u.z - uv * v.z));
    o.Dv = dv3(2.0 * (uu * v.x - uv * u.x), 2.0 * (uu * v.y - uv * u.y), 2.0 * (uu * v.z - uv * u.z));
    if (!hess) return;
    o.Duu = m3_zero(); m3_add_diag(o.Duu, 2.0 * vv); m3_add_outer(o.Duu, -2.0, v, v);
    o.Dvv = m3_zero(); m3_add_diag(o.Dvv, 2.0 * uu); m3_add_outer(o.Dvv, -2.0, u, u);
    o.Duv = m3_zero(); m3_add_outer(o.Duv, 4.0, u, v); m3_add_diag(o.Duv, -2.0 * uv); m3_add_outer(o.Duv, -2.0, v, u);
}

// f(w,u,v) = det[w,u,v]^2/|u x v|^2: gradient blocks g[3] and the six upper Hessian blocks
// Hb = {ww, wu, wv, uu, uv, vv}
CIPC_HD void wuv_derivs(const dv3& w, const dv3& u, const dv3& v, dv3* g, m3* Hb)
{
    const dv3 n = cross(u, v);
    const double N = dot(w, n);
    ddata dd;
    d_derivs(u, v, dd, Hb != nullptr);
    const double iD = 1.0 / dd.D;
    const dv3 gN[3] = {n, cross(v, w), cross(w, u)};
    const dv3 zero(0.0, 0.0, 0.0);
    const dv3 gD[3] = {zero, dd.Du, dd.Dv};
    const double c1 = 2.0 * N * iD, c2 = N * N * iD * iD;
    for (int a = 0; a < 3; ++a)
        g[a] = dv3(c1 * gN[a].x - c2 * gD[a].x, c1 * gN[a].y - c2 * gD[a].y, c1 * gN[a].z - c2 * gD[a].z);
    if (!Hb) return;
    const double A1 = 2.0 * iD, A2 = 2.0 * N * iD, A3 = 2.0 * N * iD * iD, A4 = c2, A5 = 2.0 * N * N * iD * iD * iD;
    const int ia[6] = {0, 0, 0, 1, 1, 2}, ib[6] = {0, 1, 2, 1, 2, 2};
    for (int k = 0; k < 6; ++k) {
        const int a = ia[k], b = ib[k];
        m3 B = m3_zero();
        m3_add_outer(B, A1, gN[a], gN[b]);
        m3_add_outer(B, -A3, gN[a], gD[b]);
        m3_add_outer(B, -A3, gD[a], gN[b]);
        m3_add_outer(B, A5, gD[a], gD[b]);
        Hb[k] = B;
    }
    // A2 * d2N: (w,u) = -[v]x, (w,v) = [u]x, (u,v) = -[w]x
    { const m3 S = m3_skew(v, -A2); for (int i = 0; i < 9; ++i) Hb[1].a[i] += S.a[i]; }
    { const m3 S = m3_skew(u, A2); for (int i = 0; i < 9; ++i) Hb[2].a[i] += S.a[i]; }
    { const m3 S = m3_skew(w, -A2); for (int i = 0; i < 9; ++i) Hb[4].a[i] += S.a[i]; }
    // -A4 * d2D
    for (int i = 0; i < 9; ++i) { Hb[3].a[i] -= A4 * dd.Duu.a[i]; Hb[4].a[i] -= A4 * dd.Duv.a[i]; Hb[5].a[i] -= A4 * dd.Dvv.a[i]; }
}

// pull-back coefficient tables: C[a][I] for y-block a, x-block I
//   PT: x=(p,t0,t1,t2), y=(p-t0, t1-t0, t2-t0)      EE: x=(a0,a1,b0,b1), y=(b0-a0, a1-a0, b1-b0)
#define CIPC_C_PT {{1, -1, 0, 0}, {0, -1, 1, 0}, {0, -1, 0, 1}}
#define CIPC_C_EE {{-1, 0, 1, 0}, {-1, 1, 0, 0}, {0, 0, -1, 1}}

// gradient (12) and optional Hessian (n=12 row-major, ZEROED by the caller or accumulated with
// scale `sH`) of the PT (ee=false) or EE (ee=true) squared distance; x = 4 points.
CIPC_HD void d4_derivs(bool ee, const dv3* x, double* g, double* H, double sH)
{
    const int Cpt[3][4] = CIPC_C_PT, Cee[3][4] = CIPC_C_EE;
    dv3 w, u, v;
    if (!ee) { w = x[0] - x[1]; u = x[2] - x[1]; v = x[3] - x[1]; }
    else { w = x[2] - x[0]; u = x[1] - x[0]; v = x[3] - x[2]; }
    dv3 gy[3];
    m3 Hb[6];
    wuv_derivs(w, u, v, gy, H ? Hb : nullptr);
    for (int I = 0; I < 4; ++I) {
        double sx = 0, sy = 0, sz = 0;
        for (int a = 0; a < 3; ++a) {
            const int c = ee ? Cee[a][I] : Cpt[a][I];
            sx += c * gy[a].x; sy += c * gy[a].y; sz += c * gy[a].z;
        }
        g[3 * I] = sx; g[3 * I + 1] = sy; g[3 * I + 2] = sz;
    }
    if (!H) return;
    const int idx[3][3] = {{0, 1, 2}, {1, 3, 4}, {2, 4, 5}};
    for (int I = 0; I < 4; ++I)
        for (int J = I; J < 4; ++J) {
            m3 B = m3_zero();
            for (int a = 0; a < 3; ++a) {
                const int ca = ee ? Cee[a][I] : Cpt[a][I];
                if (!ca) continue;
                for (int b = 0; b < 3; ++b) {
                    const int cb = ee ? Cee[b][J] : Cpt[b][J];
                    if (!cb) continue;
                    const m3& S = Hb[idx[a][b]];
                    const double s = (double)(ca * cb);
                    if (a <= b) for (int i = 0; i < 9; ++i) B.a[i] += s * S.a[i];
                    else for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) B.a[3 * r + c] += s * S.a[3 * c + r];
                }
            }
            if (I == J) { // diagonal block: write once (blk_add mirrors only off-diagonal blocks)
                for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) H[(3 * I + r) * 12 + 3 * I + c] += sH * B.a[3 * r + c];
            }
            else blk_add(H, 12, I, J, B, sH);
        }
}

// PE: x=(p,e0,e1); rows[] gives the destination block of each point in an n x n frame
CIPC_HD void pe_derivs(const dv3& p, const dv3& e0, const dv3& e1, double* g9, double* H, int n, const int* rows, double sH)
{
    const dv3 w = p - e0, u = e1 - e0;
    const double c = dot(w, u), L = norm2(u), iL = 1.0 / L;
    const dv3 fw(2.0 * w.x - 2.0 * c * iL * u.x, 2.0 * w.y - 2.0 * c * iL * u.y, 2.0 * w.z - 2.0 * c * iL * u.z);
    const double k = 2.0 * c * c * iL * iL;
    const dv3 fu(k * u.x - 2.0 * c * iL * w.x, k * u.y - 2.0 * c * iL * w.y, k * u.z - 2.0 * c * iL * w.z);
    g9[0] = fw.x; g9[1] = fw.y; g9[2] = fw.z;
    g9[3] = -fw.x - fu.x; g9[4] = -fw.y - fu.y; g9[5] = -fw.z - fu.z;
    g9[6] = fu.x; g9[7] = fu.y; g9[8] = fu.z;
    if (!H) return;
    m3 Hww = m3_zero(), Hwu = m3_zero(), Huu = m3_zero();
    m3_add_diag(Hww, 2.0); m3_add_outer(Hww, -2.0 * iL, u, u);
    m3_add_outer(Hwu, -2.0 * iL, u, w); m3_add_diag(Hwu, -2.0 * c * iL); m3_add_outer(Hwu, 4.0 * c * iL * iL, u, u);
    m3_add_outer(Huu, -2.0 * iL, w, w); m3_add_outer(Huu, 4.0 * c * iL * iL, w, u); m3_add_outer(Huu, 4.0 * c * iL * iL, u, w);
    m3_add_diag(Huu, 2.0 * c * c * iL * iL); m3_add_outer(Huu, -8.0 * c * c * iL * iL * iL, u, u);
    // x-blocks: p = w ; e0 = -w-u ; e1 = u
    // x-blocks with C = [[1,-1,0],[0,-1,1]]:  (p,p)=Hww (p,e0)=-(Hww+Hwu) (p,e1)=Hwu
    //   (e0,e0)=Hww+Hwu+Hwu^T+Huu (e0,e1)=-(Hwu+Huu) (e1,e1)=Huu
    m3 Hp0, H01, H00;
    for (int r = 0; r < 3; ++r)
        for (int cc = 0; cc < 3; ++cc) {
            const int i = 3 * r + cc, t = 3 * cc + r;
            Hp0.a[i] = -(Hww.a[i] + Hwu.a[i]);
            H01.a[i] = -(Hwu.a[i] + Huu.a[i]);
            H00.a[i] = Hww.a[i] + Hwu.a[i] + Hwu.a[t] + Huu.a[i];
        }
    auto put = [&](int I, int J, const m3& B) {
        const int bi = rows[I], bj = rows[J];
        for (int r = 0; r < 3; ++r)
            for (int cc = 0; cc < 3; ++cc) {
                const double v = sH * B.a[3 * r + cc];
                H[(3 * bi + r) * n + 3 * bj + cc] += v;
                if (I != J) H[(3 * bj + cc) * n + 3 * bi + r] += v;
            }
    };
    put(0, 0, Hww); put(0, 1, Hp0); put(0, 2, Hwu); put(1, 1, H00); put(1, 2, H01); put(2, 2, Huu);
}

// PP: gradient (6); Hessian adds sH * 2 [I -I; -I I] into blocks rows[0], rows[1]
CIPC_HD void pp_derivs(const dv3& a, const dv3& b, double* g6, double* H, int n, const int* rows, double sH)
{
    g6[0] = 2.0 * (a.x - b.x); g6[1] = 2.0 * (a.y - b.y); g6[2] = 2.0 * (a.z - b.z);
    g6[3] = -g6[0]; g6[4] = -g6[1]; g6[5] = -g6[2];
    if (!H) return;
    const int b0 = 3 * rows[0], b1 = 3 * rows[1];
    for (int r = 0; r < 3; ++r) {
        H[(b0 + r) * n + b0 + r] += 2.0 * sH;
        H[(b1 + r) * n + b1 + r] += 2.0 * sH;
        H[(b0 + r) * n + b1 + r] -= 2.0 * sH;
        H[(b1 + r) * n + b0 + r] -= 2.0 * sH;
    }
}

// cross-norm^2 c = |u x v|^2, u = a1-a0, v = b1-b0: gradient (12) and Hessian (12x12, +=, scale sH)
CIPC_HD void eecn2_derivs(const dv3* x, double* g, double* H, double sH)
{
    ddata dd;
    d_derivs(x[1] - x[0], x[3] - x[2], dd, H != nullptr);
    g[0] = -dd.Du.x; g[1] = -dd.Du.y; g[2] = -dd.Du.z; g[3] = dd.Du.x; g[4] = dd.Du.y; g[5] = dd.Du.z;
    g[6] = -dd.Dv.x; g[7] = -dd.Dv.y; g[8] = -dd.Dv.z; g[9] = dd.Dv.x; g[10] = dd.Dv.y; g[11] = dd.Dv.z;
    if (!H) return;
    // blocks: (a0,a0)=Duu (a0,a1)=-Duu (a1,a1)=Duu ; (b*,b*) likewise with Dvv ; (a_i,b_j) = +-Duv
    const double sa[2] = {-1.0, 1.0};
    for (int i = 0; i < 2; ++i)
        for (int j = 0; j < 2; ++j)
            for (int r = 0; r < 3; ++r)
                for (int c = 0; c < 3; ++c) {
                    const double s = sH * sa[i] * sa[j];
                    H[(3 * i + r) * 12 + 3 * j + c] += s * dd.Duu.a[3 * r + c];
                    H[(6 + 3 * i + r) * 12 + 6 + 3 * j + c] += s * dd.Dvv.a[3 * r + c];
                    H[(3 * i + r) * 12 + 6 + 3 * j + c] += s * dd.Duv.a[3 * r + c];
                    H[(6 + 3 * j + c) * 12 + 3 * i + r] += s * dd.Duv.a[3 * r + c];
                }
}

// mollifier scalar parts (EDGE_EDGE_MOLLIFIER.h:440-459)
CIPC_HD double eem(double x, double eps) { const double r = x / eps; return (-r + 2.0) * r; }
CIPC_HD double eem_g(double x, double eps) { const double o = 1.0 / eps; return 2.0 * o * (-o * x + 1.0); }
CIPC_HD double eem_H(double eps) { return -2.0 / (eps * eps); }

} // namespace cipc
