"""Pins the CPU oracle: (1) against oracle/_ref -- the reference's OWN distance / classifier /
derivative / mollifier / barrier / ACCD code compiled against a stub Eigen -- when it was built
(this container), and (2) against tests/golden/ref_probes.npz, vectors generated from oracle/_ref by
tests/golden/make_golden.py, which travel to machines without /root/reference."""
import os

import numpy as np
import pytest

from oracle import cipc_oracle as O

GOLD = os.path.join(os.path.dirname(__file__), "golden", "ref_probes.npz")
NDOF = {0: 6, 1: 9, 2: 12, 3: 12, 4: 12}


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


def test_golden_distances_and_derivatives(gold):
    for kind in range(5):
        X = gold["x_%d" % kind]
        n = NDOF[kind]
        for i, x in enumerate(X):
            d, g, H = O.dist_derivs(kind, x)
            assert d == gold["d_%d" % kind][i]  # same expression order: bit-exact
            gr = gold["g_%d" % kind][i][:n]; Hr = gold["H_%d" % kind][i][:n * n].reshape(n, n)
            assert np.abs(g - gr).max() <= 1e-11 * np.abs(gr).max()
            assert np.linalg.norm(H - Hr) <= 1e-11 * np.linalg.norm(Hr)


def test_golden_types_unclassified_accd(gold):
    X, DX = gold["x_cls"], gold["dx_cls"]
    for i, x in enumerate(X):
        for k in (1, 2, 3):
            assert O.dist_type(k, x) == gold["type_%d" % k][i]
            assert O.dist2_unclassified(k, x) == gold["unc_%d" % k][i]
        for k in range(4):
            ok, toc = O.accd(k, x, DX[i], 0.1, 0.01, 1.0)
            assert ok == bool(gold["accd_ok_%d" % k][i])
            if ok:
                assert toc == gold["accd_toc_%d" % k][i]


def test_golden_mollifier_and_barrier(gold):
    for i, x in enumerate(gold["x_mol"]):
        eps = gold["eps_mol"][i]
        e, g, H = O.mollifier(x, eps)
        assert e == gold["e_mol"][i]
        assert np.abs(g - gold["g_mol"][i]).max() <= 1e-11 * max(np.abs(gold["g_mol"][i]).max(), 1e-300)
        assert np.abs(H - gold["H_mol"][i]).max() <= 1e-11 * max(np.abs(gold["H_mol"][i]).max(), 1e-300)
    for i, (d, dh) in enumerate(gold["bar_in"]):
        for el in (0, 1):
            assert np.allclose(O.barrier_fn(d, dh, [3e4, 0, 0], bool(el)), gold["bar_out"][i, el], rtol=1e-13, atol=0)


def test_known_answer_derivtest_inputs():
    """the fixed inputs of derivTest_EECross / derivTest_e (Math/Distance/EDGE_EDGE_MOLLIFIER.h:387-391):
    analytic gradient / Hessian of the cross-norm and of the mollifier against central differences"""
    v = np.array([0, 0, 0, 1, 0.1, 0, 0, 1.1, -0.1, 0, 0.1, -1.1], float)
    for fn in (lambda x: O.dist_derivs(4, x), lambda x: O.mollifier(x, 10.0)):
        f0, g, H = fn(v)
        eps = 1e-6
        gfd = np.zeros(12); Hfd = np.zeros((12, 12))
        for i in range(12):
            e = np.zeros(12); e[i] = eps
            fp, gp, _ = fn(v + e); fm, gm, _ = fn(v - e)
            gfd[i] = (fp - fm) / (2 * eps); Hfd[:, i] = (gp - gm) / (2 * eps)
        assert np.linalg.norm(g - gfd) <= 1e-8 * np.linalg.norm(gfd)
        assert np.linalg.norm(H - Hfd) <= 1e-7 * np.linalg.norm(Hfd)


@pytest.mark.skipif(O.ref() is None, reason="oracle/_ref not built (no /root/reference on this machine)")
def test_live_against_reference_compiled():
    rng = np.random.default_rng(11)
    for it in range(400):
        x = rng.normal(size=12) * (10.0 ** rng.integers(-3, 2))
        dx = rng.normal(size=12)
        for kind in range(5):
            d, g, H = O.dist_derivs(kind, x); d2, g2, H2 = O.dist_derivs(kind, x, "ref")
            assert d == d2
            assert np.abs(g - g2).max() <= 1e-10 * np.abs(g2).max()
            assert np.linalg.norm(H - H2) <= 1e-10 * np.linalg.norm(H2)
        for k in (1, 2, 3):
            assert O.dist_type(k, x) == O.dist_type(k, x, "ref")
            assert O.dist2_unclassified(k, x) == O.dist2_unclassified(k, x, "ref")
        # ACCD is only defined for separated pairs (d > thickness); the reference loops forever otherwise
        xi = 1e-3
        for k in range(4):
            d0 = O.dist2_unclassified(k, x) if k else float(np.sum((x[:3] - x[3:6]) ** 2))
            if d0 > (2 * xi) ** 2:
                assert O.accd(k, x, dx, 0.1, xi, 1.0) == O.accd(k, x, dx, 0.1, xi, 1.0, "ref")
    # near-degenerate classification inputs: point above a vertex / an edge of the triangle
    tri = np.array([0, 0, 0, 1, 0, 0, 0, 1, 0], float)
    for p in ([0, 0, 0.1], [0.5, 0, 0.1], [1, 0, 0.1], [0.5, 0.5, 0.1], [-1, -1, 0.3], [2, -1, 0.1], [0.25, 0.25, -0.2]):
        x = np.concatenate([p, tri])
        assert O.dist_type(2, x) == O.dist_type(2, x, "ref")
