"""GPU parity against the REFERENCE ITSELF: the CUDA path (through the C ABI mirror) is compared with
(1) tests/golden/ref_drivers.npz -- outputs of the reference's own FEM/IPC.h / Grid/SPATIAL_HASH.h drivers
(generator: tests/golden/make_golden_drivers.py) and (2) oracle/_ref/libcipc_refdrv.so run live on this
machine's host cores when the library travelled here.  Gates as in test_gpu_parity.py: constraint sets
bit-exact as sorted index sets, E / g / H within 1e-9, step size <= the reference's and within 1e-12."""
import os

import numpy as np
import pytest

from golden_cases import CASES, block_summaries
from helpers import sort_cs, max_block_rel_err

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "ref_drivers.npz")
TOL = 1e-9


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


def _trip_arrays(t):
    return np.ascontiguousarray(t["row"]), np.ascontiguousarray(t["col"]), np.ascontiguousarray(t["val"])


@pytest.mark.parametrize("name", list(CASES))
def test_cuda_matches_reference_golden(ctx, gold, name):
    sc = CASES[name]()
    ctx.set_scene(sc)
    cs, info = sort_cs(*ctx.constraint_set(sc["dHat2"], sc["xi"]))
    assert np.array_equal(cs, gold[name + "/cs"]) and np.array_equal(info, gold[name + "/info"])
    ctx.set_constraints(cs, info)
    E = ctx.barrier_energy(sc["dHat2"], sc["kappa"], sc["xi"], E=0.25)
    assert abs(E - float(gold[name + "/E"])) <= TOL * abs(float(gold[name + "/E"]))
    g = ctx.barrier_gradient(sc["dHat2"], sc["kappa"], sc["xi"])
    assert np.abs(g - gold[name + "/g"]).max() <= TOL * np.abs(gold[name + "/g"]).max()
    for spd in (1, 0):
        r, c, v = _trip_arrays(ctx.barrier_hessian(sc["dHat2"], sc["kappa"], sc["xi"], projectSPD=bool(spd)))
        s, ref = block_summaries(cs, r, c, v), gold[name + "/H%d" % spd]
        assert np.array_equal(s[:, 2], ref[:, 2])
        assert np.all(np.abs(s[:, 0] - ref[:, 0]) <= TOL * ref[:, 0])
        assert np.all(np.abs(s[:, 1] - ref[:, 1]) <= TOL * 12 * ref[:, 0])
    d, m = ctx.min_dist2(sc["xi"])
    assert np.array_equal(d, gold[name + "/dist2"]) and m == float(gold[name + "/minDist2"])
    for k, scale in enumerate((1.0, 40.0)):
        ctx.set_search_dir(sc["p"] * scale)
        a, a_ref = ctx.step_size(sc["xi"], 1.0), float(gold[name + "/step"][k])
        assert a <= a_ref and (a_ref - a) <= 1e-12 * a_ref


def _live_cases():
    from codim_ipc_b200 import scenes
    return {
        "stack_48x6": lambda: scenes.cloth_stack(48, 6),
        "stack_40x6_xi": lambda: scenes.cloth_stack(40, 6, xi=1e-3),
        "sphere_64": lambda: scenes.cloth_on_sphere(64, draped=True),
        "noodles_8x80": lambda: scenes.noodles(8, 80),
        "granules_6k": lambda: scenes.granules(6000, cloth_n=25),
    }


def _refdrv_present():
    from oracle import cipc_oracle as O
    return os.path.exists(os.path.join(os.path.dirname(O.__file__), "_ref", "libcipc_refdrv.so"))


@pytest.mark.skipif(not _refdrv_present(), reason="oracle/_ref/libcipc_refdrv.so did not travel to this machine")
@pytest.mark.parametrize("name", list(_live_cases()))
def test_cuda_matches_reference_live(ctx, name):
    from oracle import cipc_oracle as O
    sc = _live_cases()[name]()
    R = O.RefScene(sc)
    ctx.set_scene(sc)
    cs, info = sort_cs(*ctx.constraint_set(sc["dHat2"], sc["xi"]))
    cr, ir = sort_cs(*R.constraint_set(sc["dHat2"], sc["xi"]))
    assert len(cs) > 0 and np.array_equal(cs, cr) and np.array_equal(info, ir)
    ctx.set_constraints(cs, info)
    a = (cs, info, sc["dHat2"], sc["kappa"], sc["xi"])
    E_r = R.barrier(*a, E0=0.5)
    assert abs(ctx.barrier_energy(sc["dHat2"], sc["kappa"], sc["xi"], E=0.5) - E_r) <= TOL * abs(E_r)
    g_r = R.barrier_gradient(*a)
    assert np.abs(ctx.barrier_gradient(sc["dHat2"], sc["kappa"], sc["xi"]) - g_r).max() <= TOL * np.abs(g_r).max()
    for spd in (True, False):
        rr, cc, vr = R.barrier_hessian(*a, projectSPD=spd)
        r, c, v = _trip_arrays(ctx.barrier_hessian(sc["dHat2"], sc["kappa"], sc["xi"], projectSPD=spd))
        assert np.array_equal(r, rr) and np.array_equal(c, cc)
        assert max_block_rel_err(cs, v, vr) <= TOL
    d, m = ctx.min_dist2(sc["xi"]); dr, mr = R.min_dist2(cs, sc["xi"])
    assert np.array_equal(d, dr) and m == mr
    for scale in (1.0, 40.0):
        ctx.set_search_dir(sc["p"] * scale)
        al, ar = ctx.step_size(sc["xi"], 1.0), R.step_size(sc["p"] * scale, sc["xi"])
        assert al <= ar and (ar - al) <= 1e-12 * ar
