"""GPU parity of the lagged-friction path (FEM/FRICTION.h) through the C ABI mirror:
(1) against tests/golden/ref_drivers.npz -- outputs of the reference's own FRICTION.h compiled from its sources
(generator tests/golden/make_golden_drivers.py), (2) against the CPU oracle / the live reference library on larger
scenes.  Gates: friction constraint set identical (order preserved), closest points / tangent bases / normal forces,
potential, gradient and per-block Hessians within 1e-9 relative."""
import os

import numpy as np
import pytest

from golden_cases import CASES, FRICTION, block_summaries, used_closest
from helpers import sort_cs, max_block_rel_err

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "ref_drivers.npz")
TOL = 1e-9


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


def _close(a, b, tol=TOL):
    a = np.asarray(a); b = np.asarray(b)
    return np.abs(a - b).max() <= tol * max(np.abs(b).max(), 1e-300)


def _trip_arrays(t):
    return np.ascontiguousarray(t["row"]), np.ascontiguousarray(t["col"]), np.ascontiguousarray(t["val"])


def _summaries_close(s, ref):
    assert np.array_equal(s[:, 2], ref[:, 2])
    assert np.all(np.abs(s[:, 0] - ref[:, 0]) <= TOL * ref[:, 0] + 1e-300)
    assert np.all(np.abs(s[:, 1] - ref[:, 1]) <= TOL * 12 * ref[:, 0] + 1e-300)


@pytest.mark.parametrize("name", list(CASES))
def test_friction_matches_reference_golden(ctx, gold, name):
    sc = CASES[name]()
    ctx.set_scene(sc)
    cs, info = gold[name + "/cs"], gold[name + "/info"]
    ctx.set_constraints(cs, info)
    fcs, cp, B, nf = ctx.friction_basis(sc["dHat2"], sc["kappa"], sc["xi"])
    assert np.array_equal(fcs, gold[name + "/f_cs"])
    assert _close(used_closest(fcs, cp), gold[name + "/f_cp"]) and _close(B, gold[name + "/f_B"]) and _close(nf, gold[name + "/f_nf"])
    rng = np.random.default_rng(FRICTION["seed"])
    for k, mag in enumerate(FRICTION["mags"]):
        Xn = sc["X"] - rng.normal(size=sc["X"].shape) * mag
        ctx.set_prev_positions(Xn)
        E = ctx.friction_energy(FRICTION["epsvh2"], FRICTION["mu"], E=0.1)
        Er = float(gold[name + "/f_E%d" % k])
        assert abs(E - Er) <= TOL * abs(Er)
        assert _close(ctx.friction_gradient(FRICTION["epsvh2"], FRICTION["mu"]), gold[name + "/f_g%d" % k])
        for spd in (1, 0):
            r, c, v = _trip_arrays(ctx.friction_hessian(FRICTION["epsvh2"], FRICTION["mu"], bool(spd)))
            _summaries_close(block_summaries(fcs, r, c, v), gold[name + "/f_H%d%d" % (k, spd)])


def _live_cases():
    from codim_ipc_b200 import scenes
    return {
        "stack_48x6": lambda: scenes.cloth_stack(48, 6),
        "sphere_64": lambda: scenes.cloth_on_sphere(64, draped=True),
        "noodles_8x80": lambda: scenes.noodles(8, 80),
        "granules_6k": lambda: scenes.granules(6000, cloth_n=25),
    }


def _cpu_scene(sc):
    """the reference's own FRICTION.h when oracle/_ref travelled here, else the oracle port"""
    from oracle import cipc_oracle as O
    return O.RefScene(sc) if O.refdrv() is not None else O.OracleScene(sc)


@pytest.mark.parametrize("name", list(_live_cases()))
def test_friction_pipeline_matches_cpu(ctx, name):
    """whole friction stage on the device-produced constraint set, with per-component coefficients and both
    sliding regimes (|u| below / above eps_v h) in one Xn; triplets compared entry by entry"""
    sc = _live_cases()[name]()
    S = _cpu_scene(sc)
    ctx.set_scene(sc)
    cs, info = sort_cs(*ctx.constraint_set(sc["dHat2"], sc["xi"]))
    ctx.set_constraints(cs, info)
    fcs, cp, B, nf = ctx.friction_basis(sc["dHat2"], sc["kappa"], sc["xi"])
    fo, cpo, Bo, nfo = S.friction_basis(cs, info, sc["dHat2"], sc["kappa"], sc["xi"])
    assert len(fcs) > 0 and np.array_equal(fcs, fo)
    assert _close(used_closest(fcs, cp), used_closest(fo, cpo)) and _close(B, Bo) and _close(nf, nfo)
    # two components split at the middle node, asymmetric coefficient table
    nV = len(sc["X"])
    ranges, muComp = np.array([nV // 2, nV], np.int32), np.array([[0.3, 0.7], [0.5, 0.2]])
    assert ctx.friction_coef(ranges, muComp) == 1.0
    S.friction_coef(ranges, muComp)
    nf2 = ctx.get_friction_basis()[3]
    assert _close(nf2, S.fetch_friction(len(fo))[3]) and not np.array_equal(nf2, nf)
    rng = np.random.default_rng(11)
    mag = np.where(rng.random(nV) < 0.5, 1e-7, 1e-4)[:, None]
    Xn = sc["X"] - rng.normal(size=sc["X"].shape) * mag
    Xn[: nV // 50] = sc["X"][: nV // 50]  # some nodes did not move: |u| == 0 branch
    ctx.set_prev_positions(Xn)
    epsvh2, mu = 1e-10, 0.35
    Er = S.friction_potential(Xn, epsvh2, mu, E0=2.0)
    assert abs(ctx.friction_energy(epsvh2, mu, E=2.0) - Er) <= TOL * abs(Er)
    g0 = rng.normal(size=(nV, 3))
    g = ctx.friction_gradient(epsvh2, mu, g0.copy())
    gr = S.friction_gradient(Xn, epsvh2, mu)
    assert np.abs((g - g0) - gr).max() <= TOL * np.abs(gr).max() + 1e-15 * np.abs(g0).max()
    for spd in (True, False):
        rr, cc, vr = S.friction_hessian(Xn, epsvh2, mu, spd)
        r, c, v = _trip_arrays(ctx.friction_hessian(epsvh2, mu, spd))
        assert np.array_equal(r, rr) and np.array_equal(c, cc)
        assert max_block_rel_err(fcs, v, vr) <= TOL


def test_friction_negative_normal_force_and_delivery_paths(ctx, monkeypatch):
    """a stale friction set (normal force < 0) gives negative semi-definite blocks in the reference (no projection on
    that factor); the factor form carries the sign.  Host expansion and device expansion (DMA) agree."""
    from codim_ipc_b200 import scenes
    from oracle import cipc_oracle as O
    sc = scenes.mixed_small()
    S = O.OracleScene(sc)
    ctx.set_scene(sc)
    cs, info = sort_cs(*ctx.constraint_set(sc["dHat2"], sc["xi"]))
    ctx.set_constraints(cs, info)
    fcs, cp, B, nf = ctx.friction_basis(sc["dHat2"], sc["kappa"], sc["xi"])
    nf = nf.copy(); nf[::3] *= -1.0
    ctx.set_friction_basis(fcs, cp, B, nf)
    S.set_friction(fcs, cp, B, nf)
    rng = np.random.default_rng(3)
    Xn = sc["X"] - rng.normal(size=sc["X"].shape) * 3e-6
    ctx.set_prev_positions(Xn)
    rr, cc, vr = S.friction_hessian(Xn, 1e-10, 0.4, True)
    t_host = ctx.friction_hessian(1e-10, 0.4, True)
    assert np.array_equal(t_host["row"], rr) and np.array_equal(t_host["col"], cc)
    assert max_block_rel_err(fcs, np.ascontiguousarray(t_host["val"]), vr) <= TOL
    from helpers import split_blocks
    assert any(np.trace(b) < 0 for b in split_blocks(fcs, vr)) and any(np.trace(b) > 0 for b in split_blocks(fcs, vr))
    monkeypatch.setenv("CIPC_TRIPLETS_DMA", "1")
    t_dev = ctx.friction_hessian(1e-10, 0.4, True)
    assert np.array_equal(t_dev["row"], t_host["row"]) and np.array_equal(t_dev["col"], t_host["col"])
    assert np.abs(t_dev["val"] - t_host["val"]).max() <= 1e-13 * np.abs(t_host["val"]).max()


def test_friction_reference_named_entry_points():
    import codim_ipc_b200 as cipc
    from codim_ipc_b200 import scenes
    from oracle import cipc_oracle as O
    sc = scenes.mixed_small()
    S = O.OracleScene(sc)
    cs, info = sort_cs(*S.constraint_set(sc["dHat2"], sc["xi"]))
    X4 = np.zeros((len(sc["X"]), 4)); X4[:, :3] = sc["X"]
    fcs, cp, B, nf = cipc.Compute_Friction_Basis(X4, cs, info, sc["dHat2"], sc["kappa"], sc["xi"])
    fo, cpo, Bo, nfo = S.friction_basis(cs, info, sc["dHat2"], sc["kappa"], sc["xi"])
    assert np.array_equal(fcs, fo) and _close(nf, nfo)
    nV = len(sc["X"])
    nf, mu = cipc.Compute_Friction_Coef(fcs, cp, B, np.array([nV], np.int32), np.array([0.25]), nf)
    assert mu == 1.0 and _close(nf, 0.25 * nfo)
    S.friction_coef(np.array([nV], np.int32), np.array([0.25]))
    rng = np.random.default_rng(5)
    Xn4 = X4.copy(); Xn4[:, :3] -= rng.normal(size=sc["X"].shape) * 1e-5
    Xn = np.ascontiguousarray(Xn4[:, :3])
    E = cipc.Compute_Friction_Potential(X4, Xn4, fcs, cp, B, nf, 1e-10, mu, 1.5)
    assert abs((E - 1.5) - S.friction_potential(Xn, 1e-10, mu)) <= TOL * abs(E - 1.5)
    g = np.ones((nV, 4))
    cipc.Compute_Friction_Gradient(X4, Xn4, fcs, cp, B, nf, 1e-10, mu, g)
    gr = S.friction_gradient(Xn, 1e-10, mu)
    assert np.abs(g[:, :3] - 1.0 - gr).max() <= TOL * np.abs(gr).max() and np.all(g[:, 3] == 1.0)
    pre = np.zeros(5, cipc.TRIPLET_DTYPE); pre["val"] = 2.0
    trip = cipc.Compute_Friction_Hessian(X4, Xn4, fcs, cp, B, nf, 1e-10, mu, True, pre)
    rr, cc, vr = S.friction_hessian(Xn, 1e-10, mu, True)
    assert len(trip) == 5 + len(vr) and np.all(trip["val"][:5] == 2.0) and np.array_equal(trip["row"][5:], rr)
    assert max_block_rel_err(fcs, np.ascontiguousarray(trip["val"][5:]), vr) <= TOL


def test_friction_device_resident_accumulates(ctx):
    """_dev variants: friction energy in scalars[4], friction gradient added onto the barrier gradient in HBM"""
    import torch
    from codim_ipc_b200 import scenes, multi
    sc = scenes.cloth_stack(24, 4)
    ctx.set_scene(sc)
    ctx.constraint_set(sc["dHat2"], sc["xi"], fetch=False)
    ctx.friction_basis(sc["dHat2"], sc["kappa"], sc["xi"], fetch=False)
    rng = np.random.default_rng(9)
    Xn = sc["X"] - rng.normal(size=sc["X"].shape) * 1e-5
    ctx.set_prev_positions(Xn)
    gb = ctx.barrier_gradient(sc["dHat2"], sc["kappa"], sc["xi"])
    gf = ctx.friction_gradient(1e-10, 0.4)
    Ef = ctx.friction_energy(1e-10, 0.4)
    ctx.barrier_gradient_dev(sc["dHat2"], sc["kappa"], sc["xi"])
    ctx.friction_gradient_dev(1e-10, 0.4, accumulate=True)
    ctx.friction_energy_dev(1e-10, 0.4)
    ctx.sync()
    nV = len(sc["X"])
    g = multi.wrap_device_f64(ctx.dev_ptrs()["g"], 3 * nV, 0).cpu().numpy().reshape(nV, 3)
    scal = multi.wrap_device_f64(ctx.dev_ptrs()["scalars"], 16, 0).cpu().numpy()
    assert np.abs(g - (gb + gf)).max() <= 1e-12 * np.abs(gb + gf).max()
    assert scal[4] == Ef


def test_fused_friction_hessian_equals_factor_path(ctx):
    """cipc_friction_hessian_dev (fused factor + expansion, stream left in HBM) == factor kernel + host expansion"""
    from codim_ipc_b200 import scenes
    sc = scenes.mixed_small()
    ctx.set_scene(sc)
    ctx.constraint_set(sc["dHat2"], sc["xi"], fetch=False)
    fcs, cp, B, nf = ctx.friction_basis(sc["dHat2"], sc["kappa"], sc["xi"])
    nf = nf.copy(); nf[::4] *= -1.0  # include negated blocks
    ctx.set_friction_basis(fcs, cp, B, nf)
    rng = np.random.default_rng(13)
    ctx.set_prev_positions(sc["X"] - rng.normal(size=sc["X"].shape) * np.where(rng.random(len(sc["X"])) < 0.5, 1e-7, 1e-4)[:, None])
    a = ctx.friction_hessian(1e-10, 0.4, True).copy()
    n = ctx.friction_hessian_dev(1e-10, 0.4, True)
    b = ctx.get_triplets(n)
    assert n == len(a) > 0 and np.array_equal(a["row"], b["row"]) and np.array_equal(a["col"], b["col"])
    assert np.abs(a["val"] - b["val"]).max() <= 1e-13 * np.abs(a["val"]).max()
