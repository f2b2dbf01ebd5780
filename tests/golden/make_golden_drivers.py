"""Generates tests/golden/ref_drivers.npz from oracle/_ref/libcipc_refdrv.so -- the reference's OWN
drivers (FEM/IPC.h, Grid/SPATIAL_HASH.h, FEM/FRICTION.h compiled from /root/reference by
oracle/Makefile `ref`) run on the seeded scenes of codim-ipc_b200/scenes.py.  Run in the container that
has /root/reference; the .npz is committed so that machines without the reference (and the GPU box,
should oracle/_ref not have travelled) can still check the oracle and the CUDA path against outputs
of the reference itself.

    python tests/golden/make_golden_drivers.py

Hessians are stored per constraint block as (Frobenius norm, v^T H v) with v = quad_vector(n) -- 16 B
instead of up to 1152 B per block -- which is enough to catch any wrong entry, sign or ordering.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import cipc_oracle as O  # noqa: E402
from golden_cases import CASES, FRICTION, block_summaries, used_closest  # noqa: E402
from helpers import sort_cs  # noqa: E402

assert O.refdrv() is not None, "oracle/_ref/libcipc_refdrv.so is not built"
out = {}
for name, mk in CASES.items():
    sc = mk()
    R = O.RefScene(sc)
    cs, info = sort_cs(*R.constraint_set(sc["dHat2"], sc["xi"]))
    out[name + "/cs"], out[name + "/info"] = cs, info
    out[name + "/E"] = np.array(R.barrier(cs, info, sc["dHat2"], sc["kappa"], sc["xi"], E0=0.25))
    out[name + "/g"] = R.barrier_gradient(cs, info, sc["dHat2"], sc["kappa"], sc["xi"])
    for spd in (1, 0):
        r, c, v = R.barrier_hessian(cs, info, sc["dHat2"], sc["kappa"], sc["xi"], projectSPD=bool(spd))
        out[name + "/H%d" % spd] = block_summaries(cs, r, c, v)
    d, m = R.min_dist2(cs, sc["xi"])
    out[name + "/dist2"], out[name + "/minDist2"] = d, np.array(m)
    out[name + "/step"] = np.array([R.step_size(sc["p"] * s, sc["xi"]) for s in (1.0, 40.0)])
    # friction (FEM/FRICTION.h)
    fcs, cp, B, nf = R.friction_basis(cs, info, sc["dHat2"], sc["kappa"], sc["xi"])
    out[name + "/f_cs"], out[name + "/f_cp"], out[name + "/f_B"], out[name + "/f_nf"] = fcs, used_closest(fcs, cp), B, nf
    rng = np.random.default_rng(FRICTION["seed"])
    for k, mag in enumerate(FRICTION["mags"]):
        Xn = sc["X"] - rng.normal(size=sc["X"].shape) * mag
        out[name + "/f_E%d" % k] = np.array(R.friction_potential(Xn, FRICTION["epsvh2"], FRICTION["mu"], E0=0.1))
        out[name + "/f_g%d" % k] = R.friction_gradient(Xn, FRICTION["epsvh2"], FRICTION["mu"])
        for spd in (1, 0):
            r, c, v = R.friction_hessian(Xn, FRICTION["epsvh2"], FRICTION["mu"], bool(spd))
            out[name + "/f_H%d%d" % (k, spd)] = block_summaries(fcs, r, c, v)
    print(name, len(cs), "constraints,", len(fcs), "friction stencils")
path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_drivers.npz")
np.savez_compressed(path, **out)
print("wrote", path, os.path.getsize(path), "bytes")
