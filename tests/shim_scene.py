"""TEST INFRASTRUCTURE: ctypes front-end of tests/shim_harness/_build/libcipc_shimdrv.so -- the reference's own drivers
(oracle/ref_build/ref_drivers.cpp) compiled THROUGH the drop-in shim against the reference's real Storage / VECTOR / FEM
headers, so the same C API (ref_*) runs on the CUDA path.  ShimScene has the interface of oracle.cipc_oracle.RefScene."""
import ctypes as C
import os

import numpy as np

from oracle import cipc_oracle as O

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "shim_harness", "_build", "libcipc_shimdrv.so")
_LIB = None


def present():
    return os.path.exists(LIB_PATH)


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(LIB_PATH)
        L.ref_scene_create.restype = C.c_void_p
        L.ref_barrier_hessian.restype = C.c_long
        L.ref_friction_hessian.restype = C.c_long
        L.ref_friction_coef.restype = C.c_double
        L.ref_timer.restype = C.c_double
        L.ref_timer_parent.restype = C.c_char_p
        L.ref_triplets_data.restype = C.c_void_p
        _LIB = L
    return _LIB


class ShimScene(O.RefScene):
    def _lib(self):
        return lib()

    def timer_parent(self, name):
        return lib().ref_timer_parent(name.encode()).decode()

    def selfcheck(self, sc, Xn=None, epsvh2=1e-10, mu=0.4, raw=True):
        """GPU templates vs the reference's *_CPU templates inside the one binary -> array of 19 error figures (shim_selfcheck)"""
        out = np.zeros(19)
        k = np.ascontiguousarray(sc["kappa"], np.float64); p = np.ascontiguousarray(sc["p"], np.float64)
        xn = np.ascontiguousarray(Xn, np.float64) if Xn is not None else None
        lib().shim_selfcheck(self.h, C.c_double(sc["dHat2"]), O._dp(k), C.c_double(sc["xi"]), O._dp(p), O._dp(xn) if xn is not None else None,
                             C.c_double(epsvh2), C.c_double(mu), int(raw), O._dp(out))
        return out
