"""Drop-in boundary on the reference's REAL types: tests/shim_harness/_build/libcipc_shimdrv.so is the reference's own driver
translation unit (MESH_NODE / MESH_NODE_ATTR on Storage/*.hpp, VECTOR.h, std::vector<bool>, std::map NNExclusion) compiled
through codim-ipc_b200/shim -- every Compute_* call lands on the CUDA path, the reference's templates stay available as
Compute_*_CPU in the same binary.
  * shim_selfcheck: all six contact templates and all five friction templates, GPU vs CPU, compared in C++ in ONE binary;
  * the shim library against oracle/_ref (the same translation unit compiled without the shim) through the common C API;
  * the reference's timer sub-scopes are filled from the device stage times under their top-level scope;
  * merged (default) and raw triplet delivery."""
import os
import subprocess
import sys

import numpy as np
import pytest

import shim_scene
from helpers import sort_cs, max_block_rel_err

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not shim_scene.present(), reason="tests/shim_harness/_build/libcipc_shimdrv.so did not travel here")]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _cases():
    from codim_ipc_b200 import scenes
    return {
        "mixed_small": lambda: scenes.mixed_small(),
        "stack_48x6": lambda: scenes.cloth_stack(48, 6),
        "stack_40x6_xi": lambda: scenes.cloth_stack(40, 6, xi=1e-3),
        "sphere_64": lambda: scenes.cloth_on_sphere(64, draped=True),
        "noodles_8x80": lambda: scenes.noodles(8, 80),
        "granules_6k": lambda: scenes.granules(6000, cloth_n=25),
    }


def _xn(sc, seed=3):
    rng = np.random.default_rng(seed)
    return sc["X"] - rng.normal(size=sc["X"].shape) * np.where(rng.random(len(sc["X"])) < 0.5, 2e-6, 5e-5)[:, None]


_SELFCHECK = r"""
import sys, json
sys.path.insert(0, %(root)r); sys.path.insert(0, %(tests)r)
import numpy as np
import shim_scene
from test_gpu_shim import _cases, _xn
res = {}
for name, mk in _cases().items():
    sc = mk()
    S = shim_scene.ShimScene(sc)
    res[name] = S.selfcheck(sc, _xn(sc), raw=%(raw)d).tolist()
print("RESULT " + json.dumps(res))
"""


@pytest.mark.parametrize("mode", ["raw", "merged"])
def test_shim_selfcheck_gpu_vs_cpu_templates_in_one_binary(mode):
    """CIPC_TRIPLETS is read once per process, so each delivery mode runs in its own interpreter"""
    import json
    env = dict(os.environ, CIPC_TRIPLETS=mode)
    code = _SELFCHECK % dict(root=ROOT, tests=os.path.join(ROOT, "tests"), raw=1 if mode == "raw" else 0)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=900, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    res = json.loads([l for l in r.stdout.splitlines() if l.startswith("RESULT ")][-1][7:])
    for name, o in res.items():
        assert o[15] > 0, name
        assert o[0] == 0, (name, "constraint sets differ", o[0])
        assert o[1] <= 1e-9 and o[2] <= 1e-9, (name, "E / g", o[1], o[2])
        assert o[3] == 0 and o[5] == 0 and o[4] <= 1e-9, (name, "Hessian", o[3], o[4], o[5])
        assert o[6] <= o[7] and o[7] - o[6] <= 1e-12 * o[7], (name, "step", o[6], o[7])
        assert o[8] == 0 and o[9] == 0, (name, "dist2", o[8], o[9])
        assert o[16] > 0 and o[10] == 0, (name, "friction set", o[10])
        assert o[11] <= 1e-9 and o[12] <= 1e-9 and o[13] <= 1e-9 and o[14] <= 1e-9, (name, "friction terms", o[11:15])
        assert o[18] == 0 and o[17] <= 1e-9, (name, "CSR hand-off (Compute_Barrier_Hessian_CSR)", o[17], o[18])


@pytest.mark.parametrize("name", list(_cases()))
def test_shim_library_matches_reference_library(name):
    """the same driver translation unit, with and without the shim, through the common C API (merged delivery: the Hessian is
    compared as an assembled matrix)"""
    import scipy.sparse as sp
    from oracle import cipc_oracle as O
    if O.refdrv() is None:
        pytest.skip("oracle/_ref did not travel here")
    sc = _cases()[name]()
    S, R = shim_scene.ShimScene(sc), O.RefScene(sc)
    cs, info = sort_cs(*S.constraint_set(sc["dHat2"], sc["xi"]))
    cr, ir = sort_cs(*R.constraint_set(sc["dHat2"], sc["xi"]))
    assert len(cs) and np.array_equal(cs, cr) and np.array_equal(info, ir)
    a = (cs, info, sc["dHat2"], sc["kappa"], sc["xi"])
    E, Er = S.barrier(*a, E0=0.5), R.barrier(*a, E0=0.5)
    assert abs(E - Er) <= 1e-9 * abs(Er)
    g0 = np.full((len(sc["X"]), 3), 0.25)
    g, gr = S.barrier_gradient(*a, g=g0.copy()), R.barrier_gradient(*a, g=g0.copy())  # accumulates into nodeAttr.g
    assert np.abs(g - gr).max() <= 1e-9 * np.abs(gr).max()
    n = 3 * len(sc["X"])
    A = sp.coo_matrix((lambda r, c, v: (v, (r, c)))(*S.barrier_hessian(*a, projectSPD=True)), shape=(n, n)).tocsr()
    B = sp.coo_matrix((lambda r, c, v: (v, (r, c)))(*R.barrier_hessian(*a, projectSPD=True)), shape=(n, n)).tocsr()
    assert abs(A - B).max() <= 1e-9 * abs(B).max() and A.nnz == B.nnz
    d, m = S.min_dist2(cs, sc["xi"]); dr, mr = R.min_dist2(cs, sc["xi"])
    assert np.array_equal(d, dr) and m == mr
    al, ar = S.step_size(sc["p"], sc["xi"]), R.step_size(sc["p"], sc["xi"])
    assert al <= ar and ar - al <= 1e-12 * ar


def test_timer_sub_scopes_are_filled_under_their_parents():
    from codim_ipc_b200 import scenes
    sc = scenes.cloth_stack(48, 6)
    S = shim_scene.ShimScene(sc)
    S.timer_reset()
    S.contact_stage(sc)
    for top, subs in (("Compute_Constraint_Set", ("_Build_Hash", "_PT", "_EE", "_Merge")),
                      ("Compute_Intersection_Free_StepSize", ("_Build_Hash", "_PT", "_EE"))):
        total = S.timer(top)
        assert total > 0
        parts = [S.timer(top + sfx) for sfx in subs]
        assert all(p > 0 for p in parts), (top, parts)
        assert sum(parts) <= total  # device time of the stages is part of the call's wall time
        for sfx in subs:
            assert S.timer_parent(top + sfx) == top
    for leaf in ("Compute_Barrier", "Compute_Barrier_Gradient", "Compute_Barrier_Hessian", "Compute_Min_Dist"):
        assert S.timer(leaf) > 0


def test_contact_stage_call_pattern_matches_reference():
    from codim_ipc_b200 import scenes
    from oracle import cipc_oracle as O
    if O.refdrv() is None:
        pytest.skip("oracle/_ref did not travel here")
    sc = scenes.cloth_stack(48, 6)
    _, rs = shim_scene.ShimScene(sc).contact_stage(sc)
    _, rr = O.RefScene(sc).contact_stage(sc)
    assert rs["nC"] == rr["nC"] > 0
    assert abs(rs["E"] - rr["E"]) <= 1e-9 * abs(rr["E"])
    assert rs["step"] <= rr["step"] and rr["step"] - rs["step"] <= 1e-12 * rr["step"]
    assert rs["minDist2"] == rr["minDist2"]
    assert 0 < rs["nTriplets"] < rr["nTriplets"]  # merged delivery: one triplet per distinct (row, col)
