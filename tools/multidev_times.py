"""Contact stage through ONE multi-device context (cipc_create_multi) on 1..N GPUs of this box: host wall time per call
(pinned host buffers, merged Hessian delivery) and the slowest rank's device stage times.  Usage: multidev_times.py [n layers]"""
import os, sys, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import codim_ipc_b200 as cipc
from codim_ipc_b200 import scenes
n, L = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (224, 10)
sc = scenes.cloth_stack(n, L)
nV = len(sc["X"])
a = (sc["dHat2"], sc["kappa"], sc["xi"])
pin = lambda shape, dt: torch.empty(shape, dtype=dt).pin_memory().numpy()
X4 = pin((nV, 4), torch.float64); X4[:, :3] = sc["X"]; X4[:, 3] = 0
ndev = torch.cuda.device_count()
res = {}
for N in [k for k in (1, 2, 4, 8) if k <= ndev]:
    M = cipc.ContactContext(devices=list(range(N)))
    M.set_scene(sc)
    nC = M.constraint_set(sc["dHat2"], sc["xi"], fetch=False)
    nT = M.barrier_hessian_merged(*a, True, fetch=False)
    cs = pin((nC + 1024, 4), torch.int32); info = pin((nC + 1024, 2), torch.float64); d = pin((nC + 1024,), torch.float64)
    trip = torch.empty((int(nT * 1.2) + 1024, 2), dtype=torch.float64).pin_memory().numpy().view(cipc.TRIPLET_DTYPE).reshape(-1)
    g = pin((nV, 4), torch.float64)
    ip = lambda x, t: x.ctypes.data_as(C.POINTER(t))
    best = None
    for rep in range(5):
        t = [time.perf_counter()]
        M.set_positions(X4); M.constraint_set(sc["dHat2"], sc["xi"], fetch=False)
        M._ck(M.L.cipc_get_constraints(M.h, ip(cs, C.c_int32), ip(info, C.c_double))); t.append(time.perf_counter())
        st = {k: M.stage_ms(k) for k in ("ccs_hash_build", "ccs_pairs", "ccs_narrow", "ccs_merge")}
        E = M.barrier_energy(*a); t.append(time.perf_counter())
        g[:] = 0; M.barrier_gradient(*a, g); t.append(time.perf_counter())
        tr = M.barrier_hessian_merged(*a, True, out=trip); t.append(time.perf_counter())
        st.update({k: M.stage_ms(k) for k in ("barrier_H", "hessian_merge")})
        al = M.step_size(sc["xi"], 1.0); t.append(time.perf_counter())
        st.update({k: M.stage_ms(k) for k in ("ccd_hash_build", "ccd_pairs", "ccd_accd")})
        mm = C.c_double(0)
        M._ck(M.L.cipc_min_dist2(M.h, C.c_double(sc["xi"]), ip(d, C.c_double), C.byref(mm))); t.append(time.perf_counter())
        tot = 1e3 * (t[-1] - t[0])
        if best is None or tot < best[0]:
            best = (tot, [round(1e3 * (y - x), 2) for x, y in zip(t, t[1:])], {k: round(v, 3) for k, v in st.items()}, E, al, mm.value, len(tr))
    res[N] = best
    print("N=%d: stage %.2f ms  calls [CCS, E, g, H, step, minDist] = %s  device(max over ranks) %s  E=%.12e step=%.15f minDist2=%.6e nTrip=%d" % (
        N, best[0], best[1], best[2], best[3], best[4], best[5], best[6]), flush=True)
    M.close()
