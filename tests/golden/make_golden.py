"""Generates tests/golden/ref_probes.npz from oracle/_ref (the reference's own distance headers,
compiled from /root/reference/Library/Math by oracle/Makefile `ref`).  Run in the container that has
/root/reference; the .npz is committed so that machines without the reference can still pin the oracle.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import cipc_oracle as O  # noqa: E402

assert O.ref() is not None, "oracle/_ref is not built"
rng = np.random.default_rng(20261017)
out = {}
for kind in range(5):
    X = rng.normal(size=(40, 12)) * (10.0 ** rng.integers(-3, 2, size=(40, 1)))
    D, G, H = [], [], []
    for x in X:
        d, g, h = O.dist_derivs(kind, x, "ref")
        D.append(d); G.append(np.pad(g, (0, 12 - len(g)))); H.append(np.pad(h.ravel(), (0, 144 - h.size)))
    out["x_%d" % kind] = X; out["d_%d" % kind] = np.array(D); out["g_%d" % kind] = np.array(G); out["H_%d" % kind] = np.array(H)
X = rng.normal(size=(300, 12)); DX = rng.normal(size=(300, 12))
out["x_cls"], out["dx_cls"] = X, DX
for k in (1, 2, 3):
    out["type_%d" % k] = np.array([O.dist_type(k, x, "ref") for x in X])
    out["unc_%d" % k] = np.array([O.dist2_unclassified(k, x, "ref") for x in X])
for k in range(4):
    r = [O.accd(k, x, dx, 0.1, 0.01, 1.0, "ref") for x, dx in zip(X, DX)]
    out["accd_ok_%d" % k] = np.array([a for a, _ in r]); out["accd_toc_%d" % k] = np.array([b for _, b in r])
XM = rng.normal(size=(20, 12)); XM[:, 9:12] = XM[:, 6:9] + (XM[:, 3:6] - XM[:, 0:3]) * 1.3 + rng.normal(size=(20, 3)) * 0.05
EPS = np.array([O.dist_derivs(4, x, "ref")[0] * s for x, s in zip(XM, rng.choice([0.5, 3.0], 20))])
out["x_mol"], out["eps_mol"] = XM, EPS
r = [O.mollifier(x, e, "ref") for x, e in zip(XM, EPS)]
out["e_mol"] = np.array([a for a, _, _ in r]); out["g_mol"] = np.array([b for _, b, _ in r]); out["H_mol"] = np.array([c for _, _, c in r])
BI = np.array([[1e-8, 1e-6], [5e-7, 1e-6], [9.9e-7, 1e-6], [2e-9, 2.5e-7], [1e-16, 1e-6]])
out["bar_in"] = BI
out["bar_out"] = np.array([[O.barrier_fn(d, dh, [3e4, 0, 0], bool(el), "ref") for el in (0, 1)] for d, dh in BI])
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_probes.npz"), **out)
print("wrote ref_probes.npz")
