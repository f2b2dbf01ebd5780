"""TEST INFRASTRUCTURE -- ctypes front-end of the CPU oracle (oracle/cipc_oracle.cpp) and of
oracle/_ref (the reference's own distance headers compiled against a stub Eigen).

Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference) may import this.
The product path (codim-ipc_b200) never does.
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_REF = None
_REFDRV = None
_REFDRV_FAST = None

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int)


def _dp(a):
    return a.ctypes.data_as(c_dp)


def _ip(a):
    return a.ctypes.data_as(c_ip)


def build(force=False):
    """Compile the oracle (always possible: g++ only) and, if /root/reference exists, oracle/_ref."""
    so = os.path.join(_HERE, "libcipc_oracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("cipc_oracle.cpp", "geom.h", "derivs.h", "eig.h")]
    stale = force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs)
    if stale:
        subprocess.check_call(["make", "-C", _HERE, "libcipc_oracle.so"], stdout=subprocess.DEVNULL)
    ref_sos = [os.path.join(_HERE, "_ref", f) for f in ("libcipc_refdist.so", "libcipc_refdrv.so", "libcipc_refdrv_fast.so")]
    ref_srcs = [os.path.join(_HERE, "ref_build", f) for f in ("ref_drivers.cpp", "ref_distance.cpp")] + [os.path.join(_HERE, "Makefile")]
    if os.path.isdir("/root/reference/Library/Math/Distance") and (
            force or not all(os.path.exists(f) for f in ref_sos)
            or max(os.path.getmtime(f) for f in ref_srcs) > min(os.path.getmtime(f) for f in ref_sos)):
        subprocess.check_call(["make", "-C", _HERE, "ref"], stdout=subprocess.DEVNULL)


def lib():
    global _LIB
    if _LIB is None:
        build()
        L = C.CDLL(os.path.join(_HERE, "libcipc_oracle.so"))
        L.oracle_scene_create.restype = C.c_void_p
        L.oracle_barrier_hessian.restype = C.c_long
        L.oracle_dist2_unclassified.restype = C.c_double
        L.oracle_friction_coef.restype = C.c_double
        L.oracle_friction_hessian.restype = C.c_long
        L.oracle_triplets_data.restype = C.c_void_p
        _LIB = L
    return _LIB


def ref():
    """The reference-compiled probes, or None when oracle/_ref was never built."""
    global _REF
    if _REF is None:
        p = os.path.join(_HERE, "_ref", "libcipc_refdist.so")
        if not os.path.exists(p):
            build()
        if not os.path.exists(p):
            return None
        R = C.CDLL(p)
        R.ref_dist2_unclassified.restype = C.c_double
        R.ref_mollifier_threshold.restype = C.c_double
        _REF = R
    return _REF


def _load_refdrv(name):
    p = os.path.join(_HERE, "_ref", name)
    if not os.path.exists(p):
        build()
    if not os.path.exists(p):
        return None
    R = C.CDLL(p)
    R.ref_scene_create.restype = C.c_void_p
    R.ref_barrier_hessian.restype = C.c_long
    R.ref_friction_hessian.restype = C.c_long
    R.ref_friction_coef.restype = C.c_double
    R.ref_timer.restype = C.c_double
    R.ref_triplets_data.restype = C.c_void_p
    return R


def refdrv():
    """The reference's own drivers (see RefScene), PARITY build (-ffp-contract=off), or None when oracle/_ref was never built."""
    global _REFDRV
    if _REFDRV is None:
        _REFDRV = _load_refdrv("libcipc_refdrv.so")
    return _REFDRV


def refdrv_fast():
    """The same drivers compiled with exactly the reference's flags (default FMA contraction): the TIMING build."""
    global _REFDRV_FAST
    if _REFDRV_FAST is None:
        _REFDRV_FAST = _load_refdrv("libcipc_refdrv_fast.so")
    return _REFDRV_FAST


def num_threads():
    return lib().oracle_num_threads()


def set_num_threads(n):
    lib().oracle_set_num_threads(int(n))


class OracleScene:
    """Holds contiguous copies of a scene dict (see codim-ipc_b200 scenes) for the oracle."""
    _prefix = "oracle_"

    def _lib(self):
        return lib()

    def _fn(self, name):
        return getattr(self._lib(), self._prefix + name)

    def __init__(self, sc):
        self.X = np.ascontiguousarray(sc["X"], dtype=np.float64)
        self.X0 = np.ascontiguousarray(sc["X0"], dtype=np.float64)
        self.BN = np.ascontiguousarray(sc["BN"], dtype=np.int32)
        self.BE = np.ascontiguousarray(sc["BE"], dtype=np.int32).reshape(-1, 2)
        self.BT = np.ascontiguousarray(sc["BT"], dtype=np.int32).reshape(-1, 3)
        self.DBC = np.ascontiguousarray(sc["DBC"], dtype=np.uint8)
        self.nnx = np.ascontiguousarray(sc.get("NNX", np.zeros((0, 2), np.int32)), dtype=np.int32).reshape(-1, 2)
        self.nV = self.X.shape[0]
        self.areas = [np.ascontiguousarray(sc[k], dtype=np.float64) if sc.get(k) is not None else None
                      for k in ("BNArea", "BEArea", "BTArea")]
        ap = [(_dp(a) if a is not None else None) for a in self.areas]
        cd = sc.get("codim", (len(self.BN), len(self.BN)))
        self._fn("scene_create").restype = C.c_void_p
        self.h = C.c_void_p(self._fn("scene_create")(
            self.nV, _dp(self.X), _dp(self.X0), len(self.BN), _ip(self.BN), len(self.BE), _ip(self.BE),
            len(self.BT), _ip(self.BT), int(sc.get("nRod", 0)), int(cd[0]), int(cd[1]),
            self.DBC.ctypes.data_as(C.POINTER(C.c_uint8)), len(self.nnx), _ip(self.nnx), ap[0], ap[1], ap[2]))

    def set_X(self, X):
        self.X = np.ascontiguousarray(X, dtype=np.float64)
        self._fn("scene_set_X")(self.h, _dp(self.X))

    def __del__(self):
        try:
            self._fn("scene_destroy")(self.h)
        except Exception:
            pass

    # ---- Compute_Constraint_Set
    def constraint_set(self, dHat2, thickness, elastic=False, use_hash=True, timers=None):
        tm = np.zeros(4)
        n = self._fn("constraint_set")(self.h, int(elastic), C.c_double(dHat2), C.c_double(thickness), int(use_hash), _dp(tm))
        cs = np.zeros((n, 4), np.int32)
        info = np.zeros((n, 2), np.float64)
        if n:
            self._fn("fetch_constraints")(_ip(cs), _dp(info))
        if timers is not None:
            timers[:] = tm
        return cs, info

    # ---- Compute_Barrier (returns E added to E0)
    def barrier(self, cs, info, dHat2, kappa, thickness, elastic=False, E0=0.0):
        cs = np.ascontiguousarray(cs, np.int32); info = np.ascontiguousarray(info, np.float64)
        kappa = np.ascontiguousarray(kappa, np.float64)
        E = C.c_double(E0)
        err = self._fn("barrier")(self.h, int(elastic), _ip(cs), _dp(info), len(cs), C.c_double(dHat2), _dp(kappa),
                                   C.c_double(thickness), C.byref(E))
        if err:
            raise FloatingPointError("non-positive distance during barrier evaluation")
        return E.value

    def barrier_gradient(self, cs, info, dHat2, kappa, thickness, elastic=False, g=None):
        cs = np.ascontiguousarray(cs, np.int32); info = np.ascontiguousarray(info, np.float64)
        kappa = np.ascontiguousarray(kappa, np.float64)
        if g is None:
            g = np.zeros((self.nV, 3))
        self._fn("barrier_gradient")(self.h, int(elastic), _ip(cs), _dp(info), len(cs), C.c_double(dHat2), _dp(kappa),
                                      C.c_double(thickness), _dp(g))
        return g

    def barrier_hessian(self, cs, info, dHat2, kappa, thickness, projectSPD=True, elastic=False):
        cs = np.ascontiguousarray(cs, np.int32); info = np.ascontiguousarray(info, np.float64)
        kappa = np.ascontiguousarray(kappa, np.float64)
        n = self._fn("barrier_hessian")(self.h, int(elastic), _ip(cs), _dp(info), len(cs), C.c_double(dHat2), _dp(kappa),
                                     C.c_double(thickness), int(projectSPD))
        rows = np.zeros(n, np.int32); cols = np.zeros(n, np.int32); vals = np.zeros(n)
        if n:
            self._fn("fetch_triplets")(_ip(rows), _ip(cols), _dp(vals))
        return rows, cols, vals

    def triplets_data(self):
        """(address, count) of the 16-byte {int row; int col; double val} records of the last Hessian, in place"""
        n = C.c_long(0)
        p = self._fn("triplets_data")(C.byref(n))
        return p, n.value

    def barrier_hessian_notfetch(self, cs, info, dHat2, kappa, thickness, projectSPD=True, elastic=False):
        """computes the triplets inside the oracle without copying them out (CPU-baseline timing); returns their number"""
        cs = np.ascontiguousarray(cs, np.int32); info = np.ascontiguousarray(info, np.float64)
        kappa = np.ascontiguousarray(kappa, np.float64)
        return self._fn("barrier_hessian")(self.h, int(elastic), _ip(cs), _dp(info), len(cs), C.c_double(dHat2), _dp(kappa),
                                            C.c_double(thickness), int(projectSPD))

    def step_size(self, searchDir, thickness, stepSize=1.0, elastic=False, use_hash=True, timers=None):
        p = np.ascontiguousarray(searchDir, np.float64)
        a = C.c_double(stepSize)
        tm = np.zeros(3)
        npairs = C.c_long(0)
        err = self._fn("step_size")(self.h, int(elastic), _dp(p), C.c_double(thickness), int(use_hash), C.byref(a), _dp(tm),
                                     C.byref(npairs))
        if err:
            raise FloatingPointError("ACCD returned a zero step (reference would exit(-1))")
        if timers is not None:
            timers[:] = tm
        self.last_pairs = npairs.value
        return a.value

    # ---- friction (FEM/FRICTION.h); the friction set lives inside the oracle library between calls
    def friction_basis(self, cs, info, dHat2, kappa, thickness, elastic=False):
        cs = np.ascontiguousarray(cs, np.int32); info = np.ascontiguousarray(info, np.float64)
        kappa = np.ascontiguousarray(kappa, np.float64)
        n = self._fn("friction_basis")(self.h, int(elastic), _ip(cs), _dp(info), len(cs), C.c_double(dHat2), _dp(kappa), C.c_double(thickness))
        return self.fetch_friction(n)

    def fetch_friction(self, n):
        fcs = np.zeros((n, 4), np.int32); cp = np.zeros((n, 2)); B = np.zeros((n, 6)); nf = np.zeros(n)
        if n:
            self._fn("fetch_friction")(_ip(fcs), _dp(cp), _dp(B), _dp(nf))
        return fcs, cp, B, nf

    def set_friction(self, fcs, cp, B, nf):
        fcs = np.ascontiguousarray(fcs, np.int32); cp = np.ascontiguousarray(cp, np.float64)
        B = np.ascontiguousarray(B, np.float64); nf = np.ascontiguousarray(nf, np.float64)
        self._fn("set_friction")(_ip(fcs), _dp(cp), _dp(B), _dp(nf), len(fcs))

    def friction_coef(self, compNodeRange, muComp):
        r = np.ascontiguousarray(compNodeRange, np.int32); m = np.ascontiguousarray(muComp, np.float64)
        return self._fn("friction_coef")(len(r), _ip(r), _dp(m))

    def friction_potential(self, Xn, epsvh2, mu, E0=0.0):
        Xn = np.ascontiguousarray(Xn, np.float64)
        E = C.c_double(E0)
        self._fn("friction_potential")(self.h, _dp(Xn), C.c_double(epsvh2), C.c_double(mu), C.byref(E))
        return E.value

    def friction_gradient(self, Xn, epsvh2, mu, g=None):
        Xn = np.ascontiguousarray(Xn, np.float64)
        if g is None:
            g = np.zeros((self.nV, 3))
        self._fn("friction_gradient")(self.h, _dp(Xn), C.c_double(epsvh2), C.c_double(mu), _dp(g))
        return g

    def friction_hessian(self, Xn, epsvh2, mu, projectSPD=True):
        Xn = np.ascontiguousarray(Xn, np.float64)
        n = self._fn("friction_hessian")(self.h, _dp(Xn), C.c_double(epsvh2), C.c_double(mu), int(projectSPD))
        rows = np.zeros(n, np.int32); cols = np.zeros(n, np.int32); vals = np.zeros(n)
        if n:
            self._fn("fetch_triplets")(_ip(rows), _ip(cols), _dp(vals))
        return rows, cols, vals

    def min_dist2(self, cs, thickness):
        cs = np.ascontiguousarray(cs, np.int32)
        d = np.zeros(len(cs))
        m = C.c_double(0)
        self._fn("min_dist2")(self.h, _ip(cs), len(cs), C.c_double(thickness), _dp(d), C.byref(m))
        return d, m.value


class RefScene(OracleScene):
    """Same interface, executed by the reference's OWN drivers (oracle/_ref/libcipc_refdrv.so: FEM/IPC.h,
    Grid/SPATIAL_HASH.h, FEM/FRICTION.h compiled from /root/reference against the stubs in ref_build/stub)."""
    _prefix = "ref_"

    def _lib(self):
        L = refdrv()
        if L is None:
            raise RuntimeError("oracle/_ref/libcipc_refdrv.so is not built (needs /root/reference)")
        return L

    def timer(self, name):
        return self._lib().ref_timer(name.encode())

    def release_triplets(self):
        self._lib().ref_triplets_release()

    def contact_stage(self, sc):
        """one contact stage in the reference's own call pattern (ref_drivers.cpp ref_contact_stage): -> (times dict [s], results dict)"""
        tm = np.zeros(7); res = np.zeros(6)
        k = np.ascontiguousarray(sc["kappa"], np.float64); p = np.ascontiguousarray(sc["p"], np.float64)
        self._lib().ref_contact_stage(self.h, C.c_double(sc["dHat2"]), _dp(k), C.c_double(sc["xi"]), _dp(p), _dp(tm), _dp(res))
        names = ("Compute_Constraint_Set", "Compute_Barrier", "Compute_Barrier_Gradient", "Compute_Barrier_Hessian",
                 "Compute_Intersection_Free_StepSize", "Compute_Min_Dist2_a", "Compute_Min_Dist2_b")
        return dict(zip(names, tm.tolist())), dict(E=res[0], step=res[1], minDist2=res[2], nC=int(res[3]), nTriplets=int(res[4]), tripletSum=res[5])

    def timer_reset(self):
        self._lib().ref_timer_reset()


class RefSceneFast(RefScene):
    """RefScene on the TIMING build (the reference's own compiler flags); never used for parity."""

    def _lib(self):
        L = refdrv_fast()
        if L is None:
            raise RuntimeError("oracle/_ref/libcipc_refdrv_fast.so is not built (needs /root/reference)")
        return L


def compare_triplet_blocks(ptr_a, ptr_b, cs):
    """block-by-block comparison of two triplet streams (addresses of 16-byte records, blocks in the order of `cs`):
    -> dict(max_block_rel_err, index_mismatches, max_quad_rel_err, blocks, triplets); b is the reference side"""
    cs = np.ascontiguousarray(cs, np.int32).reshape(-1, 4)
    out = np.zeros(5)
    lib().oracle_compare_triplet_blocks(C.c_void_p(ptr_a), C.c_void_p(ptr_b), _ip(cs), len(cs), _dp(out))
    return dict(max_block_rel_err=float(out[0]), index_mismatches=int(out[1]), max_quad_rel_err=float(out[2]), blocks=int(out[3]),
                triplets=int(out[4]))


def build_boundary(X, tri, seg=None, rod=None, rodRadius=None, particle=None):
    """literal restatement of Find_Surface_Primitives_And_Compute_Area (Utils/MESHIO.h:768-834) + the seg / rod / particle
    appends of Shell/IMPLICIT_EULER.h:245-277 -> dict(BN, BE (n,2), BT (n,3), BNArea, BEArea, BTArea, codim (2,))"""
    X = np.ascontiguousarray(X, np.float64)
    ia = lambda a, k: np.ascontiguousarray(a if a is not None else np.zeros((0, k)), np.int32).reshape(-1, k) if k > 1 else \
        np.ascontiguousarray(a if a is not None else np.zeros(0), np.int32).reshape(-1)
    tri, seg, rod, particle = ia(tri, 3), ia(seg, 2), ia(rod, 2), ia(particle, 1)
    rr = np.ascontiguousarray(rodRadius if rodRadius is not None else np.zeros(len(rod)), np.float64)
    cnt = np.zeros(6, np.int32)
    lib().oracle_build_boundary(len(X), _dp(X), len(tri), _ip(tri), len(seg), _ip(seg), len(rod), _ip(rod), _dp(rr), len(particle), _ip(particle), _ip(cnt))
    BN = np.zeros(cnt[0], np.int32); BE = np.zeros((cnt[1], 2), np.int32); BT = np.zeros((cnt[2], 3), np.int32)
    BNA = np.zeros(cnt[5]); BEA = np.zeros(cnt[1] - len(seg)); BTA = np.zeros(cnt[2])  # seg edges carry no BEArea entry (IMPLICIT_EULER.h:245)
    lib().oracle_fetch_boundary(_ip(BN), _ip(BE), _ip(BT), _dp(BNA), _dp(BEA), _dp(BTA))
    return dict(BN=BN, BE=BE, BT=BT, BNArea=BNA, BEArea=BEA, BTArea=BTA, codim=cnt[3:5].copy())


# ---- per-stencil probes (kind: 0 PP, 1 PE, 2 PT, 3 EE, 4 EE cross-norm^2)
_NDOF = {0: 6, 1: 9, 2: 12, 3: 12, 4: 12}


def _x12(x):
    v = np.zeros(12)
    x = np.asarray(x, np.float64).ravel()
    v[:len(x)] = x
    return v


def dist_derivs(kind, x, which="oracle"):
    n = _NDOF[kind]
    x = _x12(x)
    d = C.c_double(0); g = np.zeros(12); H = np.zeros(144)
    fn = lib().oracle_dist_derivs if which == "oracle" else ref().ref_dist_derivs
    fn(kind, _dp(x), C.byref(d), _dp(g), _dp(H))
    return d.value, g[:n].copy(), H[:n * n].reshape(n, n).copy()


def mollifier(x, eps_x, which="oracle"):
    x = _x12(x)
    e = C.c_double(0); g = np.zeros(12); H = np.zeros(144)
    fn = lib().oracle_mollifier if which == "oracle" else ref().ref_mollifier
    fn(_dp(x), C.c_double(eps_x), C.byref(e), _dp(g), _dp(H))
    return e.value, g, H.reshape(12, 12)


def dist_type(kind, x, which="oracle"):
    x = _x12(x)
    name = {1: "pe_type", 2: "pt_type", 3: "ee_type"}[kind]
    fn = getattr(lib(), "oracle_" + name) if which == "oracle" else getattr(ref(), "ref_" + name)
    return fn(_dp(x))


def dist2_unclassified(kind, x, which="oracle"):
    x = _x12(x)
    fn = lib().oracle_dist2_unclassified if which == "oracle" else ref().ref_dist2_unclassified
    return fn(kind, _dp(x))


def accd(kind, x, dx, eta, thickness, toc, which="oracle"):
    x = _x12(x); dx = _x12(dx)
    t = C.c_double(toc)
    fn = lib().oracle_accd if which == "oracle" else ref().ref_accd
    ok = fn(kind, _dp(x), _dp(dx), C.c_double(eta), C.c_double(thickness), C.byref(t))
    return bool(ok), t.value


def barrier_fn(d, dHat, kappa, elastic=False, which="oracle"):
    k = np.ascontiguousarray(kappa, np.float64)
    out = np.zeros(3)
    fn = lib().oracle_barrier_fn if which == "oracle" else ref().ref_barrier_fn
    fn(int(elastic), C.c_double(d), C.c_double(dHat), _dp(k), _dp(out))
    return out


def friction_utils(kind, x, which="oracle"):
    """tangent basis (6, column-major 3x2) and closest-point parameters (2) of FEM/FRICTION_UTILS.h; kind 0 PP, 1 PE, 2 PT, 3 EE"""
    x = _x12(x)
    B = np.zeros(6); cp = np.zeros(2); TT = np.zeros(24)
    if which == "oracle":
        lib().oracle_friction_utils(kind, _dp(x), _dp(B), _dp(cp))
        return B, cp, None
    ref().ref_friction_utils(kind, _dp(x), _dp(B), _dp(cp), _dp(TT))
    return B, cp, TT[:2 * _NDOF[kind]].reshape(2, -1)


def friction_f(x2, epsvh, which="oracle"):
    out = np.zeros(3)
    (lib().oracle_friction_f if which == "oracle" else ref().ref_friction_f)(C.c_double(x2), C.c_double(epsvh), _dp(out))
    return out


def make_pd(H):
    H = np.ascontiguousarray(H, np.float64).copy()
    lib().oracle_make_pd(H.shape[0], _dp(H))
    return H


def sym_eig(A):
    A = np.ascontiguousarray(A, np.float64)
    n = A.shape[0]
    V = np.zeros((n, n)); d = np.zeros(n)
    lib().oracle_sym_eig(n, _dp(A), _dp(V), _dp(d))
    return d, V
