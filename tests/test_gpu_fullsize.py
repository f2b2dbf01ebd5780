"""BASELINE.json full size (cfg5: 1M-triangle cloth stack): the oracle would need minutes, so parity is checked through
size-independent properties of the path: idempotence, activity of every stencil, translation invariance of
g and H, PSD-ness and symmetry of projected blocks, directional-derivative consistency E <-> g, intersection-free step."""
import numpy as np
import pytest

from helpers import sort_cs, split_blocks

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def big(ctx):
    from codim_ipc_b200 import scenes
    sc = scenes.cloth_stack(224, 10)  # 1 003 520 triangles
    ctx.set_scene(sc)
    cs, info = ctx.constraint_set(sc["dHat2"], sc["xi"])
    return sc, cs, info


def test_constraint_set_properties(ctx, big):
    sc, cs, info = big
    assert len(sc["BT"]) == 1003520 and len(cs) > 5_000_000
    d2, m = ctx.min_dist2(sc["xi"])
    assert np.all(d2 < info[:, 1]) and m > 0  # every stencil is active and separated
    dd = cs[(cs[:, 0] < 0) & (cs[:, 3] < 0)]
    assert len(np.unique(dd[:, :3], axis=0)) == len(dd)  # PP/PE keys unique after the merge
    assert np.all(cs[:, 1] >= 0)
    ctx.set_scene(sc)
    cs2, _ = ctx.constraint_set(sc["dHat2"], sc["xi"])
    assert np.array_equal(sort_cs(cs), sort_cs(cs2))  # idempotent, order-independent


def test_energy_gradient_consistency(ctx, big):
    sc, cs, info = big
    ctx.set_scene(sc)
    ctx.set_constraints(cs, info)
    g = ctx.barrier_gradient(sc["dHat2"], sc["kappa"], sc["xi"])
    assert np.abs(g.sum(0)).max() <= 1e-9 * np.abs(g).sum()  # translation invariance
    rng = np.random.default_rng(0)
    dx = rng.normal(size=sc["X"].shape)
    h = 1e-9
    ctx.set_positions(sc["X"] + h * dx); Ep = ctx.barrier_energy(sc["dHat2"], sc["kappa"], sc["xi"])
    ctx.set_positions(sc["X"] - h * dx); Em = ctx.barrier_energy(sc["dHat2"], sc["kappa"], sc["xi"])
    ctx.set_positions(sc["X"])
    fd = (Ep - Em) / (2 * h)
    assert abs(fd - np.sum(g * dx)) <= 1e-4 * abs(fd)


def test_hessian_block_properties(ctx, big):
    sc, cs, info = big
    rng = np.random.default_rng(1)
    pick = rng.choice(len(cs), 60000, replace=False)
    sub, subinfo = cs[pick], info[pick]
    ctx.set_positions(sc["X"])
    ctx.set_constraints(sub, subinfo)
    t = ctx.barrier_hessian(sc["dHat2"], sc["kappa"], sc["xi"], True)
    tu = ctx.barrier_hessian(sc["dHat2"], sc["kappa"], sc["xi"], False)
    blocks = split_blocks(sub, t["val"])
    ublocks = split_blocks(sub, tu["val"])
    for k in range(0, len(blocks), 97):
        H, U = blocks[k], ublocks[k]
        n = H.shape[0]
        assert np.abs(H - H.T).max() <= 1e-12 * np.abs(H).max()
        w = np.linalg.eigvalsh(H)
        assert w.min() >= -1e-9 * np.abs(w).max()
        T = np.tile(np.eye(3), (n // 3, 1))
        assert np.abs(H @ T).max() <= 1e-9 * np.abs(H).max()  # rigid translations are in the null space
        wu, V = np.linalg.eigh(U)
        P = (V * np.maximum(wu, 0)) @ V.T
        assert np.linalg.norm(H - P) <= 1e-9 * np.linalg.norm(U)  # the projected block IS the PSD part of the raw block
    # row / column indices follow the stencil's vertex order
    st = sub[0]
    v0 = -st[0] - 1 if st[0] < 0 else st[0]
    assert t["row"][0] == 3 * v0 and t["col"][0] == 3 * v0


def test_step_is_intersection_free(ctx, big):
    sc, cs, info = big
    ctx.set_scene(sc)
    a = ctx.step_size(sc["xi"], 1.0)
    assert 0 < a <= 1.0
    assert ctx.step_size(sc["xi"], 0.5 * a) == 0.5 * a  # a smaller bound is returned unchanged
    ctx.set_constraints(cs, info)
    ctx.set_positions(sc["X"] + a * sc["p"])
    _, m = ctx.min_dist2(sc["xi"])
    assert m > 0
    ctx.set_positions(sc["X"])
