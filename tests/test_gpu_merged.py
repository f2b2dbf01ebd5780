"""Merged-triplet delivery (cipc_barrier_hessian_merged / cipc_friction_hessian_merged, merge.cuh): the matrix assembled from
the merged triplets equals the one Eigen's setFromTriplets builds from the raw 144/81/36-triplet blocks of
Compute_Barrier_Hessian / Compute_Friction_Hessian (Shell/INC_POTENTIAL.h:374-382, Math/CSR_MATRIX.h:49-56): identical
sparsity structure, one triplet per distinct (row, col), values within 1e-9 of the row scale."""
import numpy as np
import pytest
import scipy.sparse as sp

pytestmark = pytest.mark.gpu


def _cases():
    from codim_ipc_b200 import scenes
    return {
        "mixed_small": lambda: scenes.mixed_small(),
        "stack_48x6": lambda: scenes.cloth_stack(48, 6),
        "stack_40x6_xi": lambda: scenes.cloth_stack(40, 6, xi=1e-3),
        "sphere_64": lambda: scenes.cloth_on_sphere(64, draped=True),
        "noodles_8x80": lambda: scenes.noodles(8, 80),
        "granules_6k": lambda: scenes.granules(6000, cloth_n=25),
        "stack_112x10": lambda: scenes.cloth_stack(112, 10),
    }


def _csr(t, n):
    return sp.coo_matrix((t["val"], (t["row"], t["col"])), shape=(n, n)).tocsr()


def _check(raw, merged, n):
    A, M = _csr(raw, n), _csr(merged, n)
    key = merged["row"].astype(np.int64) * n + merged["col"]
    assert len(np.unique(key)) == len(key), "merged triplets repeat a (row, col)"
    A.sort_indices(); M.sort_indices()
    assert A.nnz == M.nnz == len(merged)  # same structure as setFromTriplets (explicit zeros kept)
    assert np.array_equal(A.indptr, M.indptr) and np.array_equal(A.indices, M.indices)
    scale = np.abs(A).max()
    assert np.abs(A.data - M.data).max() <= 1e-9 * scale
    assert abs(A - A.T).max() <= 1e-9 * scale and abs(M - M.T).max() <= 1e-12 * scale  # the merged matrix is mirrored exactly


@pytest.mark.parametrize("name", list(_cases()))
@pytest.mark.parametrize("spd", [True, False])
def test_barrier_hessian_merged_equals_assembled_raw(ctx, name, spd):
    if name == "stack_112x10" and not spd:
        pytest.skip("one large case is enough")
    sc = _cases()[name]()
    ctx.set_scene(sc)
    ctx.constraint_set(sc["dHat2"], sc["xi"], fetch=False)
    raw = ctx.barrier_hessian(sc["dHat2"], sc["kappa"], sc["xi"], spd).copy()
    merged = ctx.barrier_hessian_merged(sc["dHat2"], sc["kappa"], sc["xi"], spd)
    assert 0 < len(merged) < len(raw)
    _check(raw, merged, 3 * len(sc["X"]))


@pytest.mark.parametrize("name", ["mixed_small", "stack_48x6", "granules_6k", "noodles_8x80"])
def test_friction_hessian_merged_equals_assembled_raw(ctx, name):
    sc = _cases()[name]()
    rng = np.random.default_rng(5)
    Xn = sc["X"] - rng.normal(size=sc["X"].shape) * np.where(rng.random(len(sc["X"])) < 0.5, 2e-6, 5e-5)[:, None]
    ctx.set_scene(sc)
    ctx.set_prev_positions(Xn)
    ctx.constraint_set(sc["dHat2"], sc["xi"], fetch=False)
    assert ctx.friction_basis(sc["dHat2"], sc["kappa"], sc["xi"], fetch=False) > 0
    raw = ctx.friction_hessian(1e-10, 0.4, True).copy()
    merged = ctx.friction_hessian_merged(1e-10, 0.4, True)
    _check(raw, merged, 3 * len(sc["X"]))


def test_merged_on_empty_set(ctx):
    from codim_ipc_b200 import scenes
    sc = scenes.cloth_on_sphere(24, draped=False)
    ctx.set_scene(sc)
    if ctx.constraint_set(sc["dHat2"], sc["xi"], fetch=False) == 0:
        assert len(ctx.barrier_hessian_merged(sc["dHat2"], sc["kappa"], sc["xi"], True)) == 0
