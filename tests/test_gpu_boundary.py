"""Boundary-primitive construction on the device (cipc_build_boundary) against the oracle's literal
restatement of Utils/MESHIO.h:768-834 + Shell/IMPLICIT_EULER.h:245-277: index lists identical, areas bit-identical."""
import numpy as np
import pytest

from oracle import cipc_oracle as O

pytestmark = pytest.mark.gpu


def _scenes():
    from codim_ipc_b200 import scenes
    return {
        "cfg1": lambda: scenes.cloth_on_sphere(112, draped=False),
        "cfg2": lambda: scenes.cloth_on_sphere(207, draped=True),
        "cfg3": lambda: scenes.noodles(25, 200),
        "cfg4_50k": lambda: scenes.granules(50000),
        "cfg5_250k": lambda: scenes.cloth_stack(112, 10),
        "mixed": lambda: scenes.mixed_small(),
    }


def _same(a, b):
    for k in ("BN", "BE", "BT", "codim"):
        assert np.array_equal(a[k], b[k]), k
    for k in ("BNArea", "BEArea", "BTArea"):
        assert np.array_equal(a[k], b[k]), k  # same operation order, no contraction: bit-identical


@pytest.mark.parametrize("name", list(_scenes()))
def test_build_boundary_matches_literal_restatement(ctx, name):
    sc = _scenes()[name]()
    rr = np.full(len(sc["rodE"]), max(sc["xi"], 1e-3))
    g = ctx.build_boundary(sc["X"], sc["F"], rod=sc["rodE"], rodRadius=rr, particle=sc["particles"])
    o = O.build_boundary(sc["X"], sc["F"], rod=sc["rodE"], rodRadius=rr, particle=sc["particles"])
    _same(g, o)
    assert np.array_equal(g["BN"], sc["BN"]) and np.array_equal(g["BE"], sc["BE"]) and tuple(g["codim"]) == tuple(sc["codim"])
    g2 = ctx.build_boundary(sc["X"], sc["F"], rod=sc["rodE"], rodRadius=rr, particle=sc["particles"])  # cached on the content hash
    _same(g2, o)
    X2 = sc["X"] * 1.25  # new positions: areas change, lists stay
    _same(ctx.build_boundary(X2, sc["F"], rod=sc["rodE"], rodRadius=rr, particle=sc["particles"]),
          O.build_boundary(X2, sc["F"], rod=sc["rodE"], rodRadius=rr, particle=sc["particles"]))


def test_irregular_meshes(ctx):
    """inconsistent orientations (same-direction mentions overwrite), non-manifold fans, degenerate triangles (zero area: the
    vertex is dropped from boundaryNode), duplicated triangles, seg + rod + particle appends, strided inputs"""
    rng = np.random.default_rng(11)
    nV = 300
    X = rng.normal(size=(nV, 3))
    F = rng.integers(0, nV, size=(900, 3)).astype(np.int32)
    F = F[(F[:, 0] != F[:, 1]) & (F[:, 1] != F[:, 2]) & (F[:, 0] != F[:, 2])]
    F = np.concatenate([F, F[:40], F[40:80, ::-1]])          # duplicates and flipped duplicates
    X[250:] = X[250]                                         # vertices 250.. coincide: triangles among them have zero area
    F = np.concatenate([F, np.array([[250, 251, 252], [253, 254, 255]], np.int32)])
    seg = np.array([[5, 9], [9, 5]], np.int32)
    rod = rng.integers(0, nV, size=(50, 2)).astype(np.int32)
    rr = rng.random(50) * 0.01
    part = np.array([7, 7, 299], np.int32)
    o = O.build_boundary(X, F, seg=seg, rod=rod, rodRadius=rr, particle=part)
    _same(ctx.build_boundary(X, F, seg=seg, rod=rod, rodRadius=rr, particle=part), o)
    X4 = np.zeros((nV, 4)); X4[:, :3] = X
    F4 = np.zeros((len(F), 4), np.int32); F4[:, :3] = F; F4[:, 3] = -7
    rod4 = np.full((len(rod), 4), 99, np.int32); rod4[:, :2] = rod
    _same(ctx.build_boundary(X4, F4, seg=seg, rod=rod4, rodRadius=rr, particle=part), o)
    e = O.build_boundary(X, np.zeros((0, 3), np.int32), particle=part)
    _same(ctx.build_boundary(X, np.zeros((0, 3), np.int32), particle=part), e)
