"""Self-validation of the CPU oracle (SURVEY 8(c) "how the oracle earns trust"): finite differences
of the whole barrier potential, hash == brute force, ACCD conservativeness, PSD projection."""
import numpy as np
import pytest

from oracle import cipc_oracle as O
from helpers import sort_cs, split_blocks


@pytest.fixture(scope="module")
def mixed():
    from codim_ipc_b200 import scenes
    sc = scenes.mixed_small()
    S = O.OracleScene(sc)
    cs, info = S.constraint_set(sc["dHat2"], sc["xi"])
    return sc, S, cs, info


def test_hash_equals_brute_force(mixed):
    sc, S, cs, info = mixed
    cs_b, info_b = S.constraint_set(sc["dHat2"], sc["xi"], use_hash=False)
    assert np.array_equal(sort_cs(cs), sort_cs(cs_b)) and len(cs) > 1000
    for scale in (1.0, 30.0):
        a_h = S.step_size(sc["p"] * scale, sc["xi"], 1.0, use_hash=True)
        assert 0 < a_h <= 1.0
    # brute force sees every pair; the hash only those whose alpha-swept voxel ranges overlap, so it can only be larger
    assert S.step_size(sc["p"], sc["xi"], 1.0, use_hash=False) <= S.step_size(sc["p"], sc["xi"], 1.0, use_hash=True)


def test_gradient_is_derivative_of_energy(mixed):
    sc, S, cs, info = mixed
    g = S.barrier_gradient(cs, info, sc["dHat2"], sc["kappa"], sc["xi"])
    rng = np.random.default_rng(0)
    dx = rng.normal(size=sc["X"].shape)
    h = 1e-9
    S.set_X(sc["X"] + h * dx); Ep = S.barrier(cs, info, sc["dHat2"], sc["kappa"], sc["xi"])
    S.set_X(sc["X"] - h * dx); Em = S.barrier(cs, info, sc["dHat2"], sc["kappa"], sc["xi"])
    S.set_X(sc["X"])
    fd = (Ep - Em) / (2 * h)
    assert abs(fd - np.sum(g * dx)) <= 2e-5 * abs(fd)


def test_hessian_is_derivative_of_gradient(mixed):
    sc, S, cs, info = mixed
    r, c, v = S.barrier_hessian(cs, info, sc["dHat2"], sc["kappa"], sc["xi"], projectSPD=False)
    n = 3 * len(sc["X"])
    rng = np.random.default_rng(1)
    dx = rng.normal(size=sc["X"].shape)
    Hdx = np.zeros(n)
    np.add.at(Hdx, r, v * dx.ravel()[c])
    h = 1e-9
    S.set_X(sc["X"] + h * dx); gp = S.barrier_gradient(cs, info, sc["dHat2"], sc["kappa"], sc["xi"])
    S.set_X(sc["X"] - h * dx); gm = S.barrier_gradient(cs, info, sc["dHat2"], sc["kappa"], sc["xi"])
    S.set_X(sc["X"])
    fd = ((gp - gm) / (2 * h)).ravel()
    assert np.linalg.norm(fd - Hdx) <= 1e-4 * np.linalg.norm(fd)


def test_projected_blocks_are_psd_projections(mixed):
    sc, S, cs, info = mixed
    _, _, v0 = S.barrier_hessian(cs, info, sc["dHat2"], sc["kappa"], sc["xi"], projectSPD=False)
    _, _, v1 = S.barrier_hessian(cs, info, sc["dHat2"], sc["kappa"], sc["xi"], projectSPD=True)
    for H0, H1 in list(zip(split_blocks(cs, v0), split_blocks(cs, v1)))[::7]:
        w, U = np.linalg.eigh(H0)
        P = (U * np.maximum(w, 0)) @ U.T
        assert np.linalg.norm(H1 - P) <= 1e-10 * np.linalg.norm(H0)
        assert np.linalg.eigvalsh(H1).min() >= -1e-10 * np.abs(w).max()


def test_eigensolver_against_numpy():
    rng = np.random.default_rng(2)
    for n in (6, 9, 12):
        for _ in range(50):
            A = rng.normal(size=(n, n)); A = A + A.T
            d, V = O.sym_eig(A)
            assert np.abs(d - np.linalg.eigvalsh(A)).max() <= 1e-12 * np.abs(d).max()
            assert np.abs(V @ np.diag(d) @ V.T - A).max() <= 1e-12 * np.abs(A).max()
    A = np.diag([3.0, 1.0, 0.0, 0.0, 2.0, 5.0])  # already PSD: untouched
    assert np.array_equal(O.make_pd(A), A)


def test_accd_is_conservative():
    """toc never exceeds the first time the pair gets closer than the thickness (dense time sampling)"""
    rng = np.random.default_rng(3)
    xi = 0.05
    checked = 0
    for _ in range(300):
        x = rng.normal(size=12); dx = rng.normal(size=12) * 2
        for kind in (2, 3):
            if O.dist2_unclassified(kind, x) <= (1.5 * xi) ** 2:
                continue
            ok, toc = O.accd(kind, x, dx, 0.1, xi, 1.0)
            if not ok:
                continue
            ts = np.linspace(0, toc, 200)
            dmin = min(O.dist2_unclassified(kind, x + t * dx) for t in ts)
            assert dmin > xi * xi, (kind, toc, dmin)
            checked += 1
    assert checked > 50


def test_constraint_encoding_and_multiplicity(mixed):
    sc, S, cs, info = mixed
    assert np.all(cs[:, 1] >= 0)  # IPC.h:803 assert
    pp = cs[(cs[:, 0] < 0) & (cs[:, 2] < 0)]
    assert np.all(pp[:, 2] == -1) and np.all(pp[:, 3] <= -1)
    dedup = cs[(cs[:, 0] < 0) & (cs[:, 3] < 0)]
    assert len(np.unique(dedup[:, :3], axis=0)) == len(dedup)  # keys are unique after the merge
    assert (dedup[:, 3] < -1).any()  # some PP/PE were found from several triangles/edges
    assert np.all(info[:, 0] == 1.0)  # !elasticIPC: weights forced to 1 (IPC.h:656-660)
    dHat = np.sqrt(sc["dHat2"]) + sc["xi"]
    assert np.allclose(info[:, 1], dHat * dHat, rtol=0, atol=0)
