"""GPU parity tests proper: the CUDA path (through the C ABI / the reference-named Python mirror)
against the CPU oracle on the same seeded inputs.  Gates (BASELINE.json north_star):
constraint sets bit-exact as sorted index sets; E, g, H within 1e-9 relative; step size within
1e-12 relative and never larger than the oracle's."""
import numpy as np
import pytest

from helpers import sort_cs, max_block_rel_err, kinds

pytestmark = pytest.mark.gpu

TOL = 1e-9


def _scenes():
    from codim_ipc_b200 import scenes
    return {
        "mixed_small": scenes.mixed_small,
        "stack_24x4": lambda: scenes.cloth_stack(24, 4),
        "stack_40x6_xi": lambda: scenes.cloth_stack(40, 6, xi=1e-3),
        "sphere_small": lambda: scenes.cloth_on_sphere(40, draped=True),
        "noodles_small": lambda: scenes.noodles(4, 40),
        "granules_small": lambda: scenes.granules(2000, cloth_n=15),
    }


@pytest.fixture(scope="module", params=list(_scenes().keys()))
def case(request, ctx):
    from oracle import cipc_oracle as O
    sc = _scenes()[request.param]()
    S = O.OracleScene(sc)
    ctx.set_scene(sc)
    cs_o, info_o = S.constraint_set(sc["dHat2"], sc["xi"])
    cs_g, info_g = ctx.constraint_set(sc["dHat2"], sc["xi"])
    return dict(sc=sc, S=S, ctx=ctx, cs_o=cs_o, info_o=info_o, cs_g=cs_g, info_g=info_g, name=request.param)


def test_constraint_set_bit_exact(case):
    a, ia = sort_cs(case["cs_g"], case["info_g"])
    b, ib = sort_cs(case["cs_o"], case["info_o"])
    assert len(a) == len(b) and len(a) > 0, (len(a), len(b))
    assert np.array_equal(a, b)
    assert np.array_equal(ia, ib)  # weight 1, dHat2 = (sqrt(dHat2)+xi)^2


def test_energy_gradient_hessian(case):
    sc, S, ctx = case["sc"], case["S"], case["ctx"]
    cs, info = sort_cs(case["cs_o"], case["info_o"])
    ctx.set_scene(sc)
    ctx.set_constraints(cs, info)
    E_o = S.barrier(cs, info, sc["dHat2"], sc["kappa"], sc["xi"], E0=0.25)
    E_g = ctx.barrier_energy(sc["dHat2"], sc["kappa"], sc["xi"], E=0.25)
    assert abs(E_g - E_o) <= TOL * abs(E_o)
    g0 = np.random.default_rng(0).normal(size=(len(sc["X"]), 3))
    g_o = S.barrier_gradient(cs, info, sc["dHat2"], sc["kappa"], sc["xi"], g=g0.copy())
    g_g = ctx.barrier_gradient(sc["dHat2"], sc["kappa"], sc["xi"], g=g0.copy())
    assert np.abs((g_g - g0) - (g_o - g0)).max() <= TOL * np.abs(g_o - g0).max()
    for spd in (True, False):
        r_o, c_o, v_o = S.barrier_hessian(cs, info, sc["dHat2"], sc["kappa"], sc["xi"], projectSPD=spd)
        t = ctx.barrier_hessian(sc["dHat2"], sc["kappa"], sc["xi"], projectSPD=spd)
        assert len(t) == len(v_o)
        assert np.array_equal(t["row"], r_o) and np.array_equal(t["col"], c_o)
        assert max_block_rel_err(cs, t["val"], v_o) <= TOL


def test_step_size(case):
    sc, S, ctx = case["sc"], case["S"], case["ctx"]
    ctx.set_scene(sc)
    for scale, a0 in ((1.0, 1.0), (1.0, 0.37), (40.0, 1.0)):  # the last one triggers the span-size shrink
        p = sc["p"] * scale
        ctx.set_search_dir(p)
        a_o = S.step_size(p, sc["xi"], a0)
        a_g = ctx.step_size(sc["xi"], a0)
        assert 0 < a_g <= a_o, (a_g, a_o)
        assert (a_o - a_g) <= 1e-12 * a_o, (a_g, a_o)


def test_min_dist(case):
    sc, S, ctx = case["sc"], case["S"], case["ctx"]
    cs, info = sort_cs(case["cs_o"], case["info_o"])
    ctx.set_scene(sc)
    ctx.set_constraints(cs, info)
    d_o, m_o = S.min_dist2(cs, sc["xi"])
    d_g, m_g = ctx.min_dist2(sc["xi"])
    assert np.array_equal(d_g, d_o)  # same expression order, no contraction: bit-exact
    assert m_g == m_o


def test_all_stencil_kinds_seen():
    """the seven stencil kinds of SURVEY Appendix A are all exercised by the parity scenes"""
    from oracle import cipc_oracle as O
    seen = set()
    for f in _scenes().values():
        sc = f()
        cs, _ = O.OracleScene(sc).constraint_set(sc["dHat2"], sc["xi"])
        seen |= set(kinds(cs))
    assert {"-+++", "-++-", "-+--", "++++", "++-+", "+++-"} <= seen


def test_mollified_pp_stencil(ctx):
    """'++--' (mollified PP) rarely occurs naturally: evaluate a hand-built one"""
    from oracle import cipc_oracle as O
    from codim_ipc_b200 import scenes
    sc = scenes.cloth_stack(4, 2)
    X = sc["X"].copy()
    # two nearly parallel edges whose closest points are end points: (0,1) and (2,3)
    X[0] = [0, 0, 0]; X[1] = [1e-2, 0, 0]; X[2] = [-1e-2 - 4e-4, 3e-4, 1e-5]; X[3] = [-4e-4, 3e-4, 2e-5]
    sc["X"] = X
    sc["X0"] = X.copy()
    cs = np.array([[0, 3, -2, -3]], np.int32)  # ea0=0, eb0=3, ea1=1, eb1=2
    info = np.array([[1.0, sc["dHat2"]]])
    S = O.OracleScene(sc)
    ctx.set_scene(sc)
    ctx.set_constraints(cs, info)
    E_o = S.barrier(cs, info, sc["dHat2"], sc["kappa"], 0.0)
    E_g = ctx.barrier_energy(sc["dHat2"], sc["kappa"], 0.0)
    assert E_o != 0 and abs(E_g - E_o) <= TOL * abs(E_o)
    g_o = S.barrier_gradient(cs, info, sc["dHat2"], sc["kappa"], 0.0)
    g_g = ctx.barrier_gradient(sc["dHat2"], sc["kappa"], 0.0)
    assert np.abs(g_g - g_o).max() <= TOL * np.abs(g_o).max()
    _, _, v_o = S.barrier_hessian(cs, info, sc["dHat2"], sc["kappa"], 0.0)
    t = ctx.barrier_hessian(sc["dHat2"], sc["kappa"], 0.0)
    assert max_block_rel_err(cs, t["val"], v_o) <= TOL


def test_error_codes(ctx):
    """non-positive distance -> CIPC_ERR_NONPOSITIVE_DIST (reference: printf + exit(-1), IPC.h:773-776)"""
    import codim_ipc_b200 as cipc
    from codim_ipc_b200 import scenes
    sc = scenes.cloth_stack(4, 2, xi=1e-3)
    ctx.set_scene(sc)
    cs = np.array([[-1, 30, -1, -1]], np.int32)
    X = sc["X"].copy(); X[30] = X[0] + [0, 5e-4, 0]  # closer than xi
    ctx.set_positions(X)
    ctx.set_constraints(cs, np.ones((1, 2)))
    with pytest.raises(cipc.NonPositiveDistance):
        ctx.barrier_energy(sc["dHat2"], sc["kappa"], sc["xi"])
