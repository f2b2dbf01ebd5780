"""Host-side mirror of the reference's contact entry points (Library/FEM/IPC.h) over the C ABI of
libcipc_b200.so (include/cipc_b200.h).

The six module-level functions keep the reference's names, argument order and semantics
(accumulate / append / in-out exactly as the C++ templates do); containers are numpy arrays instead
of BASE_STORAGE / std::vector.  There is NO CPU fallback: if the CUDA library is missing or no
device is present, every call raises.

    Compute_Constraint_Set              IPC.h:19-36
    Compute_Barrier                     IPC.h:742-748
    Compute_Barrier_Gradient            IPC.h:943-948
    Compute_Barrier_Hessian             IPC.h:1258-1265
    Compute_Intersection_Free_StepSize  IPC.h:1879-1890
    Compute_Min_Dist2                   IPC.h:2246-2249
and the lagged-friction entry points of Library/FEM/FRICTION.h:
    Compute_Friction_Basis              FRICTION.h:16-25
    Compute_Friction_Coef               FRICTION.h:126-130
    Compute_Friction_Potential          FRICTION.h:172-180
    Compute_Friction_Gradient           FRICTION.h:254-262
    Compute_Friction_Hessian            FRICTION.h:381-390
"""
import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CIPC_LIB") or os.path.join(_HERE, "libcipc_b200.so")  # CIPC_LIB: an experimental build of the same library

CIPC_OK, CIPC_ERR_CUDA, CIPC_ERR_NONPOSITIVE_DIST, CIPC_ERR_ZERO_STEP, CIPC_ERR_ARG, CIPC_ERR_GRID, CIPC_ERR_UNSUPPORTED = range(7)

TRIPLET_DTYPE = np.dtype([("row", np.int32), ("col", np.int32), ("val", np.float64)])  # Eigen::Triplet<double,int>

_lib = None


class CipcError(RuntimeError):
    def __init__(self, status, msg):
        super().__init__("cipc_b200 status %d: %s" % (status, msg))
        self.status = status


class NonPositiveDistance(CipcError):
    """reference: printf("%le distance detected during barrier evaluation!") + exit(-1), IPC.h:773-776"""


class ZeroStep(CipcError):
    """reference: coordinate dump + exit(-1), IPC.h:2014-2032"""


def load_library():
    """dlopen libcipc_b200.so (built by __graft_entry__.build()).  Raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("libcipc_b200.so is not built (run `python -c 'import __graft_entry__ as g; g.build()'`); "
                           "this package has no CPU fallback")
    L = C.CDLL(LIB_PATH)
    L.cipc_last_error.restype = C.c_char_p
    L.cipc_version.restype = C.c_char_p
    L.cipc_stage_ms.restype = C.c_double
    L.cipc_event_elapsed_ms.restype = C.c_double
    L.cipc_counter.restype = C.c_int64
    L.cipc_kernel_launches.restype = C.c_int64
    L.cipc_hash_bytes.restype = C.c_uint64
    L.cipc_hash_bytes.argtypes = [C.c_void_p, C.c_size_t]
    for f in ("cipc_dev_positions", "cipc_dev_gradient", "cipc_dev_scalars", "cipc_dev_triplets"):
        getattr(L, f).restype = C.c_void_p
    _lib = L
    return L


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def _nnx_pairs(NNExclusion):
    if NNExclusion is None:
        return np.zeros((0, 2), np.int32)
    if isinstance(NNExclusion, dict):
        pairs = [(k, m) for k, ms in NNExclusion.items() for m in ms]
        return np.asarray(pairs, np.int32).reshape(-1, 2)
    return np.ascontiguousarray(NNExclusion, np.int32).reshape(-1, 2)


class ContactContext:
    """One device context (one per process / GPU).  rank/world select this process' share of the
    candidate pairs (multi-GPU)."""

    def __init__(self, device=0, rank=0, world=1, devices=None):
        """devices: list of CUDA ordinals -> ONE context over several GPUs of this box driven by this thread (cipc_create_multi:
        slab-partitioned pairs, NVLink peer exchanges inside the library); otherwise a single-device context (rank/world select
        this process' share when one process per GPU is used)."""
        L = load_library()
        self.L = L
        h = C.c_void_p()
        if devices is not None:
            d = (C.c_int * len(devices))(*[int(x) for x in devices])
            st = L.cipc_create_multi(len(devices), d, C.byref(h))
        else:
            st = L.cipc_create(int(device), int(rank), int(world), C.byref(h))
        if st != CIPC_OK:
            raise CipcError(st, "cipc_create failed (no CUDA device?) -- there is no CPU fallback")
        self.h = h
        self.rank, self.world = rank, world
        self.nV = 0
        self.nC = 0
        self.nF = 0
        self._keep = []

    def close(self):
        if getattr(self, "h", None):
            self.L.cipc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, st):
        if st == CIPC_OK:
            return
        msg = self.L.cipc_last_error(self.h).decode()
        if st == CIPC_ERR_NONPOSITIVE_DIST:
            raise NonPositiveDistance(st, "non-positive distance detected during barrier evaluation")
        if st == CIPC_ERR_ZERO_STEP:
            raise ZeroStep(st, "ACCD produced a zero step")
        raise CipcError(st, msg)

    # ---- marshalling
    def set_topology(self, nV, boundaryNode, boundaryEdge, boundaryTri, nRod, codimBNStartInd, DBCb, NNExclusion=None,
                     BNArea=None, BEArea=None, BTArea=None):
        BN = np.ascontiguousarray(boundaryNode, np.int32)
        BE = np.ascontiguousarray(boundaryEdge, np.int32).reshape(-1, np.shape(boundaryEdge)[-1] if np.ndim(boundaryEdge) == 2 else 2)
        BT = np.ascontiguousarray(boundaryTri, np.int32).reshape(-1, np.shape(boundaryTri)[-1] if np.ndim(boundaryTri) == 2 else 3)
        dbc = np.ascontiguousarray(DBCb, np.uint8)
        if len(dbc) != nV:
            raise ValueError("DBCb must have nV entries")
        nn = _nnx_pairs(NNExclusion)
        cd = (C.c_int32 * 2)(int(codimBNStartInd[0]), int(codimBNStartInd[1]))
        ar = [np.ascontiguousarray(a, np.float64) if a is not None else None for a in (BNArea, BEArea, BTArea)]
        ap = [(_p(a, C.c_double) if a is not None else None) for a in ar]
        self._ck(self.L.cipc_set_topology(self.h, int(nV), len(BN), _p(BN, C.c_int32), len(BE), _p(BE, C.c_int32), BE.shape[1] if len(BE) else 2,
                                          len(BT), _p(BT, C.c_int32), BT.shape[1] if len(BT) else 3, int(nRod), cd,
                                          _p(dbc, C.c_uint8), len(nn), _p(nn, C.c_int32), ap[0], ap[1], ap[2]))
        self.nV = int(nV)

    def _vec3(self, A):
        A = np.ascontiguousarray(A, np.float64)
        if A.ndim != 2 or A.shape[0] != self.nV or A.shape[1] not in (3, 4):
            raise ValueError("expected (nV,3) or (nV,4) float64")
        return A, A.shape[1] * 8

    def set_positions(self, X):
        A, s = self._vec3(X)
        self._ck(self.L.cipc_set_positions(self.h, _p(A, C.c_double), s))

    def set_rest_positions(self, X0):
        A, s = self._vec3(X0)
        self._ck(self.L.cipc_set_rest_positions(self.h, _p(A, C.c_double), s))

    def set_search_dir(self, p):
        A = np.ascontiguousarray(p, np.float64).reshape(-1)
        if A.size != 3 * self.nV:
            raise ValueError("searchDir must have 3*nV entries")
        self._ck(self.L.cipc_set_search_dir(self.h, _p(A, C.c_double)))

    def set_scene(self, sc):
        """Upload a scene dict (codim_ipc_b200.scenes)."""
        self.set_topology(len(sc["X"]), sc["BN"], sc["BE"], sc["BT"], sc.get("nRod", 0), sc.get("codim", (len(sc["BN"]),) * 2),
                          sc["DBC"], sc.get("NNX"), sc.get("BNArea"), sc.get("BEArea"), sc.get("BTArea"))
        self.set_positions(sc["X"])
        self.set_rest_positions(sc["X0"])
        if sc.get("p") is not None:
            self.set_search_dir(sc["p"])

    # ---- stages
    def constraint_set(self, dHat2, thickness, elasticIPC=False, fetch=True):
        n = C.c_int(0)
        self._ck(self.L.cipc_constraint_set(self.h, int(elasticIPC), C.c_double(dHat2), C.c_double(thickness), C.byref(n)))
        self.nC = n.value
        if not fetch:
            return n.value
        return self.get_constraints()

    def get_constraints(self):
        cs = np.zeros((self.nC, 4), np.int32)
        info = np.zeros((self.nC, 2), np.float64)
        self._ck(self.L.cipc_get_constraints(self.h, _p(cs, C.c_int32), _p(info, C.c_double)))
        return cs, info

    def set_constraints(self, cs, info):
        cs = np.ascontiguousarray(cs, np.int32).reshape(-1, 4)
        info = np.ascontiguousarray(info, np.float64).reshape(-1, 2)
        self._ck(self.L.cipc_set_constraints(self.h, _p(cs, C.c_int32), _p(info, C.c_double), len(cs)))
        self.nC = len(cs)

    def barrier_energy(self, dHat2, kappa, thickness, E=0.0, elasticIPC=False):
        k = (C.c_double * 3)(*[float(x) for x in kappa])
        e = C.c_double(E)
        self._ck(self.L.cipc_barrier_energy(self.h, int(elasticIPC), C.c_double(dHat2), k, C.c_double(thickness), C.byref(e)))
        return e.value

    def barrier_gradient(self, dHat2, kappa, thickness, g=None, elasticIPC=False):
        k = (C.c_double * 3)(*[float(x) for x in kappa])
        if g is None:
            g = np.zeros((self.nV, 3))
        assert g.flags.c_contiguous and g.dtype == np.float64 and g.shape[0] == self.nV
        self._ck(self.L.cipc_barrier_gradient(self.h, int(elasticIPC), C.c_double(dHat2), k, C.c_double(thickness), _p(g, C.c_double),
                                              g.shape[1] * 8))
        return g

    def barrier_hessian(self, dHat2, kappa, thickness, projectSPD=True, elasticIPC=False, fetch=True, out=None):
        k = (C.c_double * 3)(*[float(x) for x in kappa])
        n = C.c_int64(0)
        self._ck(self.L.cipc_barrier_hessian(self.h, int(elasticIPC), C.c_double(dHat2), k, C.c_double(thickness), int(projectSPD), C.byref(n)))
        if not fetch:
            return n.value
        trip = out if out is not None else np.zeros(n.value, TRIPLET_DTYPE)
        assert len(trip) >= n.value and trip.dtype == TRIPLET_DTYPE
        if n.value:
            self._ck(self.L.cipc_get_triplets(self.h, trip.ctypes.data_as(C.c_void_p)))
        return trip[:n.value]

    def barrier_hessian_merged(self, dHat2, kappa, thickness, projectSPD=True, elasticIPC=False, fetch=True, out=None):
        """Compute_Barrier_Hessian delivered as MERGED triplets: one triplet per distinct (row, col) of the contact matrix
        (what setFromTriplets would produce from the raw stream), summed on the device at 3x3-block granularity"""
        k = (C.c_double * 3)(*[float(x) for x in kappa])
        n = C.c_int64(0)
        self._ck(self.L.cipc_barrier_hessian_merged(self.h, int(elasticIPC), C.c_double(dHat2), k, C.c_double(thickness), int(projectSPD), C.byref(n)))
        if not fetch:
            return n.value
        trip = out if out is not None else np.zeros(n.value, TRIPLET_DTYPE)
        assert len(trip) >= n.value and trip.dtype == TRIPLET_DTYPE
        if n.value:
            self._ck(self.L.cipc_get_triplets(self.h, trip.ctypes.data_as(C.c_void_p)))
        return trip[:n.value]

    def step_size(self, thickness, stepSize=1.0, elasticIPC=False):
        a = C.c_double(stepSize)
        self._ck(self.L.cipc_step_size(self.h, int(elasticIPC), C.c_double(thickness), C.byref(a)))
        return a.value

    def min_dist2(self, thickness, want_dist2=True):
        d = np.zeros(self.nC) if want_dist2 else None
        m = C.c_double(0)
        self._ck(self.L.cipc_min_dist2(self.h, C.c_double(thickness), _p(d, C.c_double) if want_dist2 else None, C.byref(m)))
        return d, m.value

    # ---- lagged friction (FEM/FRICTION.h) on the resident contact constraint set
    def set_prev_positions(self, Xn):
        A, s = self._vec3(Xn)
        self._ck(self.L.cipc_set_prev_positions(self.h, _p(A, C.c_double), s))

    def friction_basis(self, dHat2, kappa, thickness, elasticIPC=False, fetch=True):
        k = (C.c_double * 3)(*[float(x) for x in kappa])
        n = C.c_int(0)
        self._ck(self.L.cipc_friction_basis(self.h, int(elasticIPC), C.c_double(dHat2), k, C.c_double(thickness), C.byref(n)))
        self.nF = n.value
        return self.get_friction_basis() if fetch else n.value

    def get_friction_basis(self):
        """-> (fricConstraintSet (n,4) int32, closestPoint (n,2), tanBasis (n,6) column-major 3x2, normalForce (n,))"""
        n = self.nF
        fcs = np.zeros((n, 4), np.int32); cp = np.zeros((n, 2)); B = np.zeros((n, 6)); nf = np.zeros(n)
        self._ck(self.L.cipc_get_friction_basis(self.h, _p(fcs, C.c_int32), _p(cp, C.c_double), _p(B, C.c_double), _p(nf, C.c_double)))
        return fcs, cp, B, nf

    def set_friction_basis(self, fcs, closestPoint, tanBasis, normalForce):
        fcs = np.ascontiguousarray(fcs, np.int32).reshape(-1, 4)
        cp = np.ascontiguousarray(closestPoint, np.float64).reshape(-1, 2)
        B = np.ascontiguousarray(tanBasis, np.float64).reshape(-1, 6)
        nf = np.ascontiguousarray(normalForce, np.float64).reshape(-1)
        if not (len(fcs) == len(cp) == len(B) == len(nf)):
            raise ValueError("friction containers differ in length")
        self._ck(self.L.cipc_set_friction_basis(self.h, _p(fcs, C.c_int32), _p(cp, C.c_double), _p(B, C.c_double), _p(nf, C.c_double), len(fcs)))
        self.nF = len(fcs)

    def friction_coef(self, compNodeRange, muComp):
        r = np.ascontiguousarray(compNodeRange, np.int32).reshape(-1)
        m = np.ascontiguousarray(muComp, np.float64).reshape(-1)
        if m.size != r.size * r.size:
            raise ValueError("muComp must have nComp x nComp entries")
        mu = C.c_double(0)
        self._ck(self.L.cipc_friction_coef(self.h, len(r), _p(r, C.c_int32), _p(m, C.c_double), C.byref(mu)))
        return mu.value

    def friction_energy(self, epsvh2, mu, E=0.0):
        e = C.c_double(E)
        self._ck(self.L.cipc_friction_energy(self.h, C.c_double(epsvh2), C.c_double(mu), C.byref(e)))
        return e.value

    def friction_gradient(self, epsvh2, mu, g=None):
        if g is None:
            g = np.zeros((self.nV, 3))
        assert g.flags.c_contiguous and g.dtype == np.float64 and g.shape[0] == self.nV
        self._ck(self.L.cipc_friction_gradient(self.h, C.c_double(epsvh2), C.c_double(mu), _p(g, C.c_double), g.shape[1] * 8))
        return g

    def friction_hessian(self, epsvh2, mu, projectSPD=True, fetch=True, out=None):
        n = C.c_int64(0)
        self._ck(self.L.cipc_friction_hessian(self.h, C.c_double(epsvh2), C.c_double(mu), int(projectSPD), C.byref(n)))
        if not fetch:
            return n.value
        trip = out if out is not None else np.zeros(n.value, TRIPLET_DTYPE)
        assert len(trip) >= n.value and trip.dtype == TRIPLET_DTYPE
        if n.value:
            self._ck(self.L.cipc_get_triplets(self.h, trip.ctypes.data_as(C.c_void_p)))
        return trip[:n.value]

    def friction_hessian_merged(self, epsvh2, mu, projectSPD=True, fetch=True, out=None):
        """Compute_Friction_Hessian delivered as merged triplets (see barrier_hessian_merged)"""
        n = C.c_int64(0)
        self._ck(self.L.cipc_friction_hessian_merged(self.h, C.c_double(epsvh2), C.c_double(mu), int(projectSPD), C.byref(n)))
        if not fetch:
            return n.value
        trip = out if out is not None else np.zeros(n.value, TRIPLET_DTYPE)
        assert len(trip) >= n.value and trip.dtype == TRIPLET_DTYPE
        if n.value:
            self._ck(self.L.cipc_get_triplets(self.h, trip.ctypes.data_as(C.c_void_p)))
        return trip[:n.value]

    def friction_hessian_dev(self, epsvh2, mu, projectSPD=True):
        """friction blocks computed and expanded on the device in one fused pass; the stream stays in HBM"""
        n = C.c_int64(0)
        self._ck(self.L.cipc_friction_hessian_dev(self.h, C.c_double(epsvh2), C.c_double(mu), int(projectSPD), C.byref(n)))
        return n.value

    def friction_energy_dev(self, epsvh2, mu):
        self._ck(self.L.cipc_friction_energy_dev(self.h, C.c_double(epsvh2), C.c_double(mu)))

    def friction_gradient_dev(self, epsvh2, mu, accumulate=True):
        self._ck(self.L.cipc_friction_gradient_dev(self.h, C.c_double(epsvh2), C.c_double(mu), int(accumulate)))

    # ---- boundary-primitive construction: Utils/MESHIO.h:768-834 + Shell/IMPLICIT_EULER.h:245-277
    def build_boundary(self, X, tri, seg=None, rod=None, rodRadius=None, particle=None):
        """-> dict(BN, BE (n,2), BT (n,3), BNArea, BEArea, BTArea, codim (2,)) in the reference's exact order"""
        X = np.ascontiguousarray(X, np.float64)
        if X.ndim != 2 or X.shape[1] not in (3, 4):
            raise ValueError("X must be (nV,3) or (nV,4) float64")

        def ia(a, k):
            a = np.zeros((0, k), np.int32) if a is None else np.ascontiguousarray(a, np.int32)
            return a.reshape(-1, a.shape[-1] if a.ndim == 2 else k) if k > 1 else a.reshape(-1)
        tri, seg, rod, particle = ia(tri, 3), ia(seg, 2), ia(rod, 2), ia(particle, 1)
        rr = np.ascontiguousarray(np.zeros(len(rod)) if rodRadius is None else rodRadius, np.float64)
        cnt = (C.c_int32 * 6)()
        self._ck(self.L.cipc_build_boundary(self.h, len(X), _p(X, C.c_double), X.shape[1] * 8, len(tri), _p(tri, C.c_int32), tri.shape[1] if len(tri) else 3,
                                            len(seg), _p(seg, C.c_int32), seg.shape[1] if len(seg) else 2, len(rod), _p(rod, C.c_int32),
                                            rod.shape[1] if len(rod) else 2, _p(rr, C.c_double), len(particle), _p(particle, C.c_int32), cnt))
        BN = np.zeros(cnt[0], np.int32); BE = np.zeros((cnt[1], 2), np.int32); BT = np.zeros((cnt[2], 3), np.int32)
        BNA = np.zeros(cnt[5]); BEA = np.zeros(cnt[1] - len(seg)); BTA = np.zeros(cnt[2])  # seg edges carry no BEArea entry (IMPLICIT_EULER.h:245)
        self._ck(self.L.cipc_get_boundary(self.h, _p(BN, C.c_int32), _p(BE, C.c_int32), 2, _p(BT, C.c_int32), 3, _p(BNA, C.c_double), _p(BEA, C.c_double),
                                          _p(BTA, C.c_double)))
        return dict(BN=BN, BE=BE, BT=BT, BNArea=BNA, BEArea=BEA, BTArea=BTA, codim=np.array([cnt[3], cnt[4]], np.int32))

    # ---- Hessian triplets -> CSR on the device
    def csr_begin(self):
        self._ck(self.L.cipc_csr_begin(self.h))

    def csr_add(self):
        """append the blocks of the Hessian computed last (barrier or friction)"""
        self._ck(self.L.cipc_csr_add(self.h))

    def csr_finish(self, fetch=True):
        """-> (rowPtr (3nV+1,), colIdx (nnz,), val (nnz,)) like Eigen's row-major SparseMatrix after setFromTriplets"""
        n = C.c_int64(0)
        self._ck(self.L.cipc_csr_finish(self.h, C.byref(n)))
        if not fetch:
            return n.value
        rp = np.zeros(3 * self.nV + 1, np.int32); ci = np.zeros(n.value, np.int32); v = np.zeros(n.value)
        self._ck(self.L.cipc_get_csr(self.h, _p(rp, C.c_int32), _p(ci, C.c_int32), _p(v, C.c_double)))
        return rp, ci, v

    # ---- device-resident line search
    def save_positions(self):
        self._ck(self.L.cipc_save_positions(self.h))

    def step_positions(self, alpha):
        self._ck(self.L.cipc_step_positions(self.h, C.c_double(alpha)))

    def get_positions(self):
        X = np.zeros((self.nV, 3))
        self._ck(self.L.cipc_get_positions(self.h, _p(X, C.c_double), 24))
        return X

    # ---- device-resident variants (results stay in HBM)
    def barrier_energy_dev(self, dHat2, kappa, thickness, elasticIPC=False):
        k = (C.c_double * 3)(*[float(x) for x in kappa])
        self._ck(self.L.cipc_barrier_energy_dev(self.h, int(elasticIPC), C.c_double(dHat2), k, C.c_double(thickness)))

    def barrier_gradient_dev(self, dHat2, kappa, thickness, elasticIPC=False):
        k = (C.c_double * 3)(*[float(x) for x in kappa])
        self._ck(self.L.cipc_barrier_gradient_dev(self.h, int(elasticIPC), C.c_double(dHat2), k, C.c_double(thickness)))

    def barrier_hessian_dev(self, dHat2, kappa, thickness, projectSPD=True, elasticIPC=False):
        """blocks computed and expanded on the device in one fused pass; the stream stays in HBM (dev_triplets)"""
        k = (C.c_double * 3)(*[float(x) for x in kappa])
        n = C.c_int64(0)
        self._ck(self.L.cipc_barrier_hessian_dev(self.h, int(elasticIPC), C.c_double(dHat2), k, C.c_double(thickness), int(projectSPD), C.byref(n)))
        return n.value

    def barrier_gradient_hessian_dev(self, dHat2, kappa, thickness, elasticIPC=False):
        """barrier gradient (dev_ptrs()['g'], overwritten) and projected Hessian (dev_triplets) in one pass over the stencils"""
        k = (C.c_double * 3)(*[float(x) for x in kappa])
        n = C.c_int64(0)
        self._ck(self.L.cipc_barrier_gradient_hessian_dev(self.h, int(elasticIPC), C.c_double(dHat2), k, C.c_double(thickness), C.byref(n)))
        return n.value

    def get_triplets(self, n):
        trip = np.zeros(n, TRIPLET_DTYPE)
        if n:
            self._ck(self.L.cipc_get_triplets(self.h, trip.ctypes.data_as(C.c_void_p)))
        return trip

    def step_size_dev(self, thickness, stepSize=1.0, elasticIPC=False):
        self._ck(self.L.cipc_step_size_dev(self.h, int(elasticIPC), C.c_double(thickness), C.c_double(stepSize)))

    def min_dist2_dev(self, thickness):
        self._ck(self.L.cipc_min_dist2_dev(self.h, C.c_double(thickness)))

    def dev_triplets(self):
        """device pointer of the (row, col, value) stream of the last barrier_hessian (expanded in HBM on first use)"""
        p = self.L.cipc_dev_triplets(self.h)
        if not p:
            raise CipcError(CIPC_ERR_CUDA, self.L.cipc_last_error(self.h).decode())
        return p

    def sync(self):
        self._ck(self.L.cipc_sync(self.h))

    def set_stream(self, cuda_stream_handle):
        """run on the caller's CUDA stream (int handle, e.g. torch.cuda.current_stream().cuda_stream); 0/None = own stream"""
        self._ck(self.L.cipc_set_stream(self.h, C.c_void_p(cuda_stream_handle or None)))

    def set_timing(self, on):
        """stage timers (stage_ms) on / off: two CUDA event records per stage scope"""
        self._ck(self.L.cipc_set_timing(self.h, int(bool(on))))

    def event_record(self, slot):
        self._ck(self.L.cipc_event_record(self.h, int(slot)))

    def event_elapsed_ms(self, a, b):
        return self.L.cipc_event_elapsed_ms(self.h, int(a), int(b))

    def dev_ptrs(self):
        return dict(X=self.L.cipc_dev_positions(self.h), g=self.L.cipc_dev_gradient(self.h), scalars=self.L.cipc_dev_scalars(self.h))

    def stage_ms(self, name):
        return self.L.cipc_stage_ms(self.h, name.encode())

    def counter(self, name):
        return self.L.cipc_counter(self.h, name.encode())


def kernel_launches():
    return load_library().cipc_kernel_launches()


# ------------------------------------------------------------------------------------------------
# reference-named entry points (module-level default context, like the header-only templates)
_default = None


def default_context():
    global _default
    if _default is None:
        _default = ContactContext(device=int(os.environ.get("LOCAL_RANK", "0")))
    return _default


def _ensure_nodes(ctx, X):
    """barrier / min-dist calls only need node storage; give the context an empty topology if none was set"""
    if ctx.nV != len(X):
        ctx.set_topology(len(X), np.zeros(0, np.int32), np.zeros((0, 2), np.int32), np.zeros((0, 3), np.int32), 0, (0, 0),
                         np.zeros(len(X), np.uint8))


def Compute_Constraint_Set(X, nodeAttr_x0, boundaryNode, boundaryEdge, boundaryTri, particle, rod, NNExclusion, BNArea, BEArea, BTArea,
                           codimBNStartInd, DBCb, dHat2, thickness, getPTEE=False, elasticIPC=False, ctx=None):
    """-> (constraintSet (n,4) int32, cs_PTEE (always empty: getPTEE is false at every reference call site), stencilInfo (n,2))"""
    if getPTEE:
        raise CipcError(CIPC_ERR_UNSUPPORTED, "getPTEE is false at every call site of the reference (SURVEY 8a3)")
    ctx = ctx or default_context()
    ctx.set_topology(len(X), boundaryNode, boundaryEdge, boundaryTri, len(rod), codimBNStartInd, DBCb, NNExclusion,
                     BNArea if elasticIPC else None, BEArea if elasticIPC else None, BTArea if elasticIPC else None)
    ctx.set_positions(X)
    ctx.set_rest_positions(nodeAttr_x0)
    cs, info = ctx.constraint_set(dHat2, thickness, elasticIPC)
    return cs, np.zeros((0, 2), np.int32), info


def Compute_Barrier(X, nodeAttr_x0, constraintSet, stencilInfo, dHat2, kappa, thickness, E, elasticIPC=False, ctx=None):
    """-> E + barrier energy (the reference accumulates into its T& E)"""
    ctx = ctx or default_context()
    _ensure_nodes(ctx, X)
    ctx.set_positions(X)
    ctx.set_rest_positions(nodeAttr_x0)
    ctx.set_constraints(constraintSet, stencilInfo)
    return ctx.barrier_energy(dHat2, kappa, thickness, E, elasticIPC)


def Compute_Barrier_Gradient(X, constraintSet, stencilInfo, dHat2, kappa, thickness, nodeAttr_x0, g, elasticIPC=False, ctx=None):
    """g (nV,3|4) float64 is accumulated in place (nodeAttr.g +=) and returned"""
    ctx = ctx or default_context()
    _ensure_nodes(ctx, X)
    ctx.set_positions(X)
    ctx.set_rest_positions(nodeAttr_x0)
    ctx.set_constraints(constraintSet, stencilInfo)
    return ctx.barrier_gradient(dHat2, kappa, thickness, g, elasticIPC)


def Compute_Barrier_Hessian(X, nodeAttr_x0, constraintSet, stencilInfo, dHat2, kappa, thickness, projectSPD, triplets=None,
                            elasticIPC=False, ctx=None):
    """-> triplets with the new blocks appended (structured array row/col/val = Eigen::Triplet layout)"""
    ctx = ctx or default_context()
    _ensure_nodes(ctx, X)
    ctx.set_positions(X)
    ctx.set_rest_positions(nodeAttr_x0)
    ctx.set_constraints(constraintSet, stencilInfo)
    new = ctx.barrier_hessian(dHat2, kappa, thickness, projectSPD, elasticIPC)
    if triplets is None or len(triplets) == 0:
        return new
    return np.concatenate([triplets, new])


def Compute_Intersection_Free_StepSize(X, boundaryNode, boundaryEdge, boundaryTri, particle, rod, NNExclusion, codimBNStartInd, DBCb,
                                       searchDir, thickness, stepSize, elasticIPC=False, ctx=None):
    """-> new stepSize (in/out parameter of the reference)"""
    ctx = ctx or default_context()
    ctx.set_topology(len(X), boundaryNode, boundaryEdge, boundaryTri, len(rod), codimBNStartInd, DBCb, NNExclusion)
    ctx.set_positions(X)
    ctx.set_search_dir(searchDir)
    return ctx.step_size(thickness, stepSize, elasticIPC)


def Compute_Min_Dist2(X, constraintSet, thickness, ctx=None):
    """-> (dist2 per constraint, minDist2 = min - thickness^2)"""
    ctx = ctx or default_context()
    _ensure_nodes(ctx, X)
    ctx.set_positions(X)
    ctx.set_constraints(constraintSet, np.ones((len(constraintSet), 2)))
    return ctx.min_dist2(thickness)


# ---- FEM/FRICTION.h
def Compute_Friction_Basis(X, contactConstraintSet, stencilInfo, dHat2, kappa, thickness, elasticIPC=False, ctx=None):
    """-> (constraintSet, closestPoint, tanBasis, normalForce): the four output containers of FRICTION.h:16-25"""
    ctx = ctx or default_context()
    _ensure_nodes(ctx, X)
    ctx.set_positions(X)
    ctx.set_constraints(contactConstraintSet, stencilInfo)
    return ctx.friction_basis(dHat2, kappa, thickness, elasticIPC)


def Compute_Friction_Coef(constraintSet, closestPoint, tanBasis, compNodeRange, muComp, normalForce, ctx=None):
    """-> (normalForce scaled per component pair, mu = 1).  The reference takes only constraintSet and normalForce; the
    other two containers make the friction set resident when it is not the one the context produced last."""
    ctx = ctx or default_context()
    ctx.set_friction_basis(constraintSet, closestPoint, tanBasis, normalForce)
    mu = ctx.friction_coef(compNodeRange, muComp)
    return ctx.get_friction_basis()[3], mu


def _friction_inputs(ctx, X, Xn, constraintSet, closestPoint, tanBasis, normalForce):
    _ensure_nodes(ctx, X)
    ctx.set_positions(X)
    ctx.set_prev_positions(Xn)
    ctx.set_friction_basis(constraintSet, closestPoint, tanBasis, normalForce)


def Compute_Friction_Potential(X, Xn, constraintSet, closestPoint, tanBasis, normalForce, epsvh2, mu, E, ctx=None):
    """-> E + friction potential (the reference accumulates into its T& E)"""
    ctx = ctx or default_context()
    _friction_inputs(ctx, X, Xn, constraintSet, closestPoint, tanBasis, normalForce)
    return ctx.friction_energy(epsvh2, mu, E)


def Compute_Friction_Gradient(X, Xn, constraintSet, closestPoint, tanBasis, normalForce, epsvh2, mu, g, ctx=None):
    """g (nV,3|4) float64 is accumulated in place (nodeAttr.g +=) and returned"""
    ctx = ctx or default_context()
    _friction_inputs(ctx, X, Xn, constraintSet, closestPoint, tanBasis, normalForce)
    return ctx.friction_gradient(epsvh2, mu, g)


def Compute_Friction_Hessian(X, Xn, constraintSet, closestPoint, tanBasis, normalForce, epsvh2, mu, projectSPD, triplets=None, ctx=None):
    """-> triplets with the friction blocks appended"""
    ctx = ctx or default_context()
    _friction_inputs(ctx, X, Xn, constraintSet, closestPoint, tanBasis, normalForce)
    new = ctx.friction_hessian(epsvh2, mu, projectSPD)
    if triplets is None or len(triplets) == 0:
        return new
    return np.concatenate([triplets, new])
