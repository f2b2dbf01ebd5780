"""Pins the oracle's DRIVERS (constraint set, barrier E/g/H, step size, min-dist, friction) to the
reference's own code: (1) against tests/golden/ref_drivers.npz -- outputs of FEM/IPC.h,
Grid/SPATIAL_HASH.h and FEM/FRICTION.h themselves (compiled from /root/reference by oracle/Makefile
`ref`, generator tests/golden/make_golden_drivers.py) -- and (2) live against oracle/_ref/libcipc_refdrv.so
on larger scenes wherever that library exists.  Index sets, distances, energies and step sizes are
compared bit for bit; gradients / Hessians to 1e-11 (product sums inside Eigen are the only difference)."""
import os

import numpy as np
import pytest

from oracle import cipc_oracle as O
from golden_cases import CASES, FRICTION, block_summaries, used_closest
from helpers import sort_cs, max_block_rel_err

GOLD = os.path.join(os.path.dirname(__file__), "golden", "ref_drivers.npz")
RTOL = 1e-11


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


def _close(a, b, tol=RTOL):
    a = np.asarray(a); b = np.asarray(b)
    return np.abs(a - b).max() <= tol * max(np.abs(b).max(), 1e-300)


def _summaries_close(a, b, tol=1e-10):
    assert np.array_equal(a[:, 2], b[:, 2])  # (row, col) checksum per block
    assert np.all(np.abs(a[:, 0] - b[:, 0]) <= tol * b[:, 0])
    assert np.all(np.abs(a[:, 1] - b[:, 1]) <= tol * b[:, 0] * 12)


@pytest.mark.parametrize("name", list(CASES))
def test_oracle_matches_reference_golden(gold, name):
    sc = CASES[name]()
    S = O.OracleScene(sc)
    cs, info = sort_cs(*S.constraint_set(sc["dHat2"], sc["xi"]))
    assert np.array_equal(cs, gold[name + "/cs"]) and np.array_equal(info, gold[name + "/info"])
    assert S.barrier(cs, info, sc["dHat2"], sc["kappa"], sc["xi"], E0=0.25) == float(gold[name + "/E"])
    assert _close(S.barrier_gradient(cs, info, sc["dHat2"], sc["kappa"], sc["xi"]), gold[name + "/g"])
    for spd in (1, 0):
        r, c, v = S.barrier_hessian(cs, info, sc["dHat2"], sc["kappa"], sc["xi"], projectSPD=bool(spd))
        _summaries_close(block_summaries(cs, r, c, v), gold[name + "/H%d" % spd])
    d, m = S.min_dist2(cs, sc["xi"])
    assert np.array_equal(d, gold[name + "/dist2"]) and m == float(gold[name + "/minDist2"])
    steps = np.array([S.step_size(sc["p"] * s, sc["xi"]) for s in (1.0, 40.0)])
    assert np.array_equal(steps, gold[name + "/step"])


@pytest.mark.parametrize("name", list(CASES))
def test_oracle_friction_matches_reference_golden(gold, name):
    sc = CASES[name]()
    S = O.OracleScene(sc)
    cs, info = gold[name + "/cs"], gold[name + "/info"]
    fcs, cp, B, nf = S.friction_basis(cs, info, sc["dHat2"], sc["kappa"], sc["xi"])
    assert np.array_equal(fcs, gold[name + "/f_cs"])
    assert _close(used_closest(fcs, cp), gold[name + "/f_cp"]) and _close(B, gold[name + "/f_B"]) and _close(nf, gold[name + "/f_nf"])
    rng = np.random.default_rng(FRICTION["seed"])
    for k, mag in enumerate(FRICTION["mags"]):
        Xn = sc["X"] - rng.normal(size=sc["X"].shape) * mag
        E = S.friction_potential(Xn, FRICTION["epsvh2"], FRICTION["mu"], E0=0.1)
        assert abs(E - float(gold[name + "/f_E%d" % k])) <= RTOL * abs(E)
        assert _close(S.friction_gradient(Xn, FRICTION["epsvh2"], FRICTION["mu"]), gold[name + "/f_g%d" % k])
        for spd in (1, 0):
            r, c, v = S.friction_hessian(Xn, FRICTION["epsvh2"], FRICTION["mu"], bool(spd))
            _summaries_close(block_summaries(fcs, r, c, v), gold[name + "/f_H%d%d" % (k, spd)])


def _live_cases():
    from codim_ipc_b200 import scenes
    return {
        "stack_40x6_xi": lambda: scenes.cloth_stack(40, 6, xi=1e-3),
        "stack_64x4": lambda: scenes.cloth_stack(64, 4),
        "sphere_64": lambda: scenes.cloth_on_sphere(64, draped=True),
        "noodles_8x80": lambda: scenes.noodles(8, 80),
        "granules_6k": lambda: scenes.granules(6000, cloth_n=25),
    }


@pytest.mark.skipif(O.refdrv() is None, reason="oracle/_ref/libcipc_refdrv.so not built (needs /root/reference)")
@pytest.mark.parametrize("name", list(_live_cases()))
def test_oracle_matches_reference_live(name):
    sc = _live_cases()[name]()
    S, R = O.OracleScene(sc), O.RefScene(sc)
    cs, info = sort_cs(*S.constraint_set(sc["dHat2"], sc["xi"]))
    cr, ir = sort_cs(*R.constraint_set(sc["dHat2"], sc["xi"]))
    assert np.array_equal(cs, cr) and np.array_equal(info, ir) and len(cs) > 0
    a = (cs, info, sc["dHat2"], sc["kappa"], sc["xi"])
    assert S.barrier(*a, E0=0.5) == R.barrier(*a, E0=0.5)
    assert _close(S.barrier_gradient(*a), R.barrier_gradient(*a))
    for spd in (True, False):
        ro, co, vo = S.barrier_hessian(*a, projectSPD=spd)
        rr, cc, vr = R.barrier_hessian(*a, projectSPD=spd)
        assert np.array_equal(ro, rr) and np.array_equal(co, cc)
        assert max_block_rel_err(cs, vo, vr) <= 1e-10
    do, mo = S.min_dist2(cs, sc["xi"]); dr, mr = R.min_dist2(cs, sc["xi"])
    assert np.array_equal(do, dr) and mo == mr
    for scale in (1.0, 40.0):  # 40x triggers the span-rule step shrink (SPATIAL_HASH.h:466-482)
        assert S.step_size(sc["p"] * scale, sc["xi"]) == R.step_size(sc["p"] * scale, sc["xi"])
    # friction
    fo = S.friction_basis(*a); fr = R.friction_basis(*a)
    assert np.array_equal(fo[0], fr[0]) and _close(used_closest(fo[0], fo[1]), used_closest(fr[0], fr[1]))
    assert _close(fo[2], fr[2]) and _close(fo[3], fr[3])
    Xn = sc["X"] - np.random.default_rng(3).normal(size=sc["X"].shape) * 3e-6
    assert abs(S.friction_potential(Xn, 1e-10, 0.3) - R.friction_potential(Xn, 1e-10, 0.3)) <= RTOL * abs(R.friction_potential(Xn, 1e-10, 0.3))
    assert _close(S.friction_gradient(Xn, 1e-10, 0.3), R.friction_gradient(Xn, 1e-10, 0.3))
    ro, co, vo = S.friction_hessian(Xn, 1e-10, 0.3, True); rr, cc, vr = R.friction_hessian(Xn, 1e-10, 0.3, True)
    assert np.array_equal(ro, rr) and np.array_equal(co, cc) and max_block_rel_err(fo[0], vo, vr) <= 1e-10
