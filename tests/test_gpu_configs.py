"""Parity on the BASELINE.json configurations at sizes the CPU oracle finishes in seconds:
cfg1 cloth (25K triangles) over sphere + floor, cfg2 draped 85K cloth with a self fold, cfg3 rods over a plate
(PE/PP-heavy, mollified stencils), cfg4 particles over cloth (thickness-offset point-point contact)."""
import numpy as np
import pytest

from helpers import sort_cs, max_block_rel_err

pytestmark = pytest.mark.gpu


def _cfg():
    from codim_ipc_b200 import scenes
    return {
        "cfg1": lambda: scenes.cloth_on_sphere(112, draped=False),
        "cfg2": lambda: scenes.cloth_on_sphere(207, draped=True),
        "cfg3": lambda: scenes.noodles(25, 200),
        "cfg4_20k": lambda: scenes.granules(20000),
    }


@pytest.mark.parametrize("name", ["cfg1", "cfg2", "cfg3", "cfg4_20k"])
def test_config_parity(ctx, name):
    from oracle import cipc_oracle as O
    sc = _cfg()[name]()
    S = O.OracleScene(sc)
    ctx.set_scene(sc)
    cs_g, info_g = ctx.constraint_set(sc["dHat2"], sc["xi"])
    cs_o, info_o = S.constraint_set(sc["dHat2"], sc["xi"])
    assert np.array_equal(sort_cs(cs_g), sort_cs(cs_o)), (len(cs_g), len(cs_o))
    if len(cs_o):
        cs, info = sort_cs(cs_o, info_o)
        ctx.set_constraints(cs, info)
        E_o = S.barrier(cs, info, sc["dHat2"], sc["kappa"], sc["xi"])
        E_g = ctx.barrier_energy(sc["dHat2"], sc["kappa"], sc["xi"])
        assert abs(E_g - E_o) <= 1e-9 * abs(E_o)
        g_o = S.barrier_gradient(cs, info, sc["dHat2"], sc["kappa"], sc["xi"])
        g_g = ctx.barrier_gradient(sc["dHat2"], sc["kappa"], sc["xi"])
        assert np.abs(g_g - g_o).max() <= 1e-9 * np.abs(g_o).max()
        # Hessian on a subset (the oracle's dense eigen-solves dominate the test time)
        sub = np.arange(0, len(cs), max(1, len(cs) // 20000))
        ctx.set_constraints(cs[sub], info[sub])
        _, _, v_o = S.barrier_hessian(cs[sub], info[sub], sc["dHat2"], sc["kappa"], sc["xi"], True)
        t = ctx.barrier_hessian(sc["dHat2"], sc["kappa"], sc["xi"], True)
        assert max_block_rel_err(cs[sub], t["val"], v_o) <= 1e-9
        ctx.set_constraints(cs, info)
        d_o, m_o = S.min_dist2(cs, sc["xi"])
        d_g, m_g = ctx.min_dist2(sc["xi"])
        assert np.array_equal(d_g, d_o) and m_g == m_o
    for a0 in (1.0, 0.3):
        a_o = S.step_size(sc["p"], sc["xi"], a0)
        a_g = ctx.step_size(sc["xi"], a0)
        assert 0 < a_g <= a_o and a_o - a_g <= 1e-12 * a_o, (name, a_g, a_o)
