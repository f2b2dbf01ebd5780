"""The product's device geometry header (codim-ipc_b200/csrc/geom.cuh, hess.cuh, eig.cuh) compiled for
the HOST and compared with the oracle: lets the CPU suite check the CUDA math without a GPU."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import cipc_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def probe(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("probe") / "geom_probe.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-x", "c++", "-ffp-contract=off", "-fPIC", "-shared", "-o", so,
                           os.path.join(ROOT, "tests", "host", "geom_host_probe.cpp")])
    P = C.CDLL(so)
    P.probe_dist2_unclassified.restype = C.c_double
    return P


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


NDOF = {0: 6, 1: 9, 2: 12, 3: 12, 4: 12}


def test_distances_and_derivatives(probe):
    rng = np.random.default_rng(1)
    for kind in range(5):
        n = NDOF[kind]
        for _ in range(100):
            x = rng.normal(size=12) * 10.0 ** rng.integers(-3, 2)
            d = C.c_double(); g = np.zeros(12); H = np.zeros(144)
            probe.probe_dist_derivs(kind, _dp(x), C.byref(d), _dp(g), _dp(H))
            d2, g2, H2 = O.dist_derivs(kind, x)
            assert d.value == d2
            assert np.abs(g[:n] - g2).max() <= 1e-11 * np.abs(g2).max()
            assert np.linalg.norm(H[:n * n].reshape(n, n) - H2) <= 1e-11 * np.linalg.norm(H2)


def test_classifiers_accd_broadphase_bit_exact(probe):
    rng = np.random.default_rng(2)
    for it in range(3000):
        x = rng.normal(size=12); dx = rng.normal(size=12)
        for k in (1, 2, 3):
            assert probe.probe_type(k, _dp(x)) == O.dist_type(k, x)
            assert probe.probe_dist2_unclassified(k, _dp(x)) == O.dist2_unclassified(k, x)
        if it < 600:
            for k in range(4):
                d0 = O.dist2_unclassified(k, x) if k else float(np.sum((x[:3] - x[3:6]) ** 2))
                if d0 <= 4e-4:
                    continue
                t = C.c_double(1.0)
                ok = probe.probe_accd(k, _dp(x), _dp(dx), C.c_double(0.1), C.c_double(0.01), C.byref(t))
                ok2, t2 = O.accd(k, x, dx, 0.1, 0.01, 1.0)
                assert bool(ok) == ok2 and (not ok2 or t.value == t2)


def test_lowrank_hessian_equals_dense_projection(probe):
    """hess.cuh: H = alpha g g^T + beta K assembled from 5 (PT/EE) / 4 (PE) / 1 (PP) vectors and projected
    through the small eigenproblem == dense eigen projection of the oracle's block"""
    rng = np.random.default_rng(3)
    for kind in range(4):
        n = NDOF[kind]
        for _ in range(200):
            x = rng.normal(size=12) * 10.0 ** rng.integers(-3, 1)
            d, g, K = O.dist_derivs(kind, x)
            alpha = rng.uniform(0.1, 5) / np.abs(g).max() ** 2
            beta = -rng.uniform(0.1, 5) / np.abs(K).max()
            H = alpha * np.outer(g, g) + beta * K
            for pr in (0, 1):
                out = np.zeros(n * n)
                probe.probe_hess_lowrank(kind, _dp(x), C.c_double(alpha), C.c_double(beta), pr, _dp(out))
                ref = H
                if pr:
                    w, V = np.linalg.eigh(H)
                    ref = (V * np.maximum(w, 0)) @ V.T
                    assert (np.abs(w) > 1e-9 * np.abs(w).max()).sum() <= 5  # rank <= 5: the structure hess.cuh relies on
                assert np.linalg.norm(out.reshape(n, n) - ref) <= 1e-10 * np.linalg.norm(H)
            # factored form used by the two-phase kernel (valid for a barrier: alpha > 0 > beta)
            out = np.zeros(n * n)
            probe.probe_hess_factor(kind, _dp(x), C.c_double(alpha), C.c_double(beta), _dp(out))
            assert np.linalg.norm(out.reshape(n, n) - ref) <= 1e-10 * np.linalg.norm(H)


def test_dense_jacobi_projection(probe):
    rng = np.random.default_rng(4)
    for n in (6, 9, 12):
        for _ in range(30):
            A = rng.normal(size=(n, n)); A = A + A.T
            B = np.ascontiguousarray(A.copy())
            probe.probe_psd_jacobi(n, _dp(B))
            w, V = np.linalg.eigh(A)
            assert np.linalg.norm(B - (V * np.maximum(w, 0)) @ V.T) <= 1e-12 * np.linalg.norm(A)


def test_ql_eigensolver_matches_lapack(probe):
    """eig.cuh sym_eig_ql (tridiagonalisation + implicit QL, the dense path's eigen-solver): eigenvalues, orthonormality and
    reconstruction against numpy.linalg.eigh, including rank-deficient, repeated-eigenvalue, diagonal and zero matrices"""
    rng = np.random.default_rng(11)
    cases = []
    for n in (3, 6, 9):
        for _ in range(300):
            A = rng.normal(size=(n, n)) * 10.0 ** rng.integers(-6, 6); cases.append(A + A.T)
        for r in range(0, n):  # rank r, indefinite
            B = rng.normal(size=(n, max(r, 1))); sgn = np.diag(rng.choice([-1.0, 1.0], size=max(r, 1)))
            cases.append((B @ sgn @ B.T) * (r > 0))
        Q = np.linalg.qr(rng.normal(size=(n, n)))[0]
        cases.append(Q @ np.diag([2.0] * (n - 1) + [-1.0]) @ Q.T)  # repeated eigenvalue
        cases.append(np.diag(rng.normal(size=n))); cases.append(np.eye(n)); cases.append(np.zeros((n, n)))
        T = np.diag(rng.normal(size=n)) + np.diag(rng.normal(size=n - 1), 1); cases.append(T + T.T)  # already tridiagonal
    for A in cases:
        n = len(A)
        Z = np.ascontiguousarray(A, dtype=np.float64).copy(); d = np.zeros(n)
        assert probe.probe_sym_eig_ql(n, _dp(Z), _dp(d)) == 1
        w = np.linalg.eigvalsh(A)
        scale = max(np.abs(w).max(), 1e-300)
        assert np.abs(np.sort(d) - w).max() <= 1e-13 * scale
        assert np.abs(Z.T @ Z - np.eye(n)).max() <= 1e-13
        assert np.abs(Z @ np.diag(d) @ Z.T - A).max() <= 1e-13 * scale
