"""Host-side costs of the shim's building blocks on pageable memory (cfg5_1m sizes): content hash, staged uploads / downloads."""
import os, sys, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import codim_ipc_b200 as cipc
from codim_ipc_b200 import scenes
L = cipc.load_library()
def T(f, n=5):
    ts = []
    for _ in range(n):
        t = time.perf_counter(); f(); ts.append(1e3 * (time.perf_counter() - t))
    return "min %.2f ms  median %.2f ms" % (min(ts), sorted(ts)[len(ts) // 2])
a = np.random.default_rng(0).integers(0, 1 << 30, size=(7237373, 4), dtype=np.int32)
print("hash 116 MB:", T(lambda: L.cipc_hash_bytes(a.ctypes.data, a.nbytes)))
sc = scenes.cloth_stack(224, 10)
ctx = cipc.ContactContext(0)
ctx.set_scene(sc)
nV = len(sc["X"])
X4 = np.zeros((nV, 4)); X4[:, :3] = sc["X"]
print("set_positions 16 MB pageable (+sync):", T(lambda: (ctx.set_positions(X4), ctx.sync())))
nC = ctx.constraint_set(sc["dHat2"], sc["xi"], fetch=False)
print("constraint_set device only:", T(lambda: ctx.constraint_set(sc["dHat2"], sc["xi"], fetch=False)))
cs = np.zeros((nC, 4), np.int32); info = np.zeros((nC, 4))
ip = lambda x, t: x.ctypes.data_as(C.POINTER(t))
print("get_constraints cs only 116 MB:", T(lambda: L.cipc_get_constraints_strided(ctx.h, ip(cs, C.c_int32), None, 32)))
print("get_constraints info fill 232 MB:", T(lambda: L.cipc_get_constraints_strided(ctx.h, None, ip(info, C.c_double), 32)))
BE4 = np.zeros((len(sc["BE"]), 4), np.int32); BE4[:, :2] = sc["BE"]
BT4 = np.zeros((len(sc["BT"]), 4), np.int32); BT4[:, :3] = sc["BT"]
ctx.set_topology(nV, sc["BN"], BE4, BT4, 0, sc["codim"], sc["DBC"]); ctx.set_positions(X4); ctx.set_rest_positions(X4)
ctx.constraint_set(sc["dHat2"], sc["xi"], fetch=False)
print("set_topology (unchanged, hash only):", T(lambda: ctx.set_topology(nV, sc["BN"], BE4, BT4, 0, sc["codim"], sc["DBC"])))
d = np.zeros(nC)
m = C.c_double(0)
print("min_dist2 with dist2 58 MB:", T(lambda: L.cipc_min_dist2(ctx.h, C.c_double(sc["xi"]), ip(d, C.c_double), C.byref(m))))
def fresh():
    dd = np.empty(nC)
    L.cipc_min_dist2(ctx.h, C.c_double(sc["xi"]), ip(dd, C.c_double), C.byref(m))
print("min_dist2 into a fresh 58 MB array:", T(fresh))
g = np.zeros((nV, 4))
print("barrier_gradient into stride-32 g:", T(lambda: ctx.barrier_gradient(sc["dHat2"], sc["kappa"], sc["xi"], g)))
