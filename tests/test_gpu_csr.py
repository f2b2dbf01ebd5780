"""Device-side triplets -> CSR assembly against scipy's COO -> CSR conversion of the same triplets,
which is what Math/CSR_MATRIX.h:49-56 (Eigen setFromTriplets) produces: duplicates summed, sorted column indices per row,
explicit zeros kept.  Structure must match exactly, values to 1e-9 of the matrix scale (only the summation order differs)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _ref_csr(trips, n):
    import scipy.sparse as sp
    r = np.concatenate([t["row"] for t in trips]); c = np.concatenate([t["col"] for t in trips]); v = np.concatenate([t["val"] for t in trips])
    # duplicates summed without dropping entries that cancel to zero: accumulate on the unique (row, col) pattern
    key = r.astype(np.int64) * n + c
    uk, inv = np.unique(key, return_inverse=True)
    val = np.zeros(len(uk)); np.add.at(val, inv, v)
    rows = (uk // n).astype(np.int64)
    rowPtr = np.zeros(n + 1, np.int64); np.add.at(rowPtr, rows + 1, 1)
    A = sp.coo_matrix((v, (r, c)), shape=(n, n)).tocsr()
    A.sort_indices()
    return np.cumsum(rowPtr).astype(np.int32), (uk % n).astype(np.int32), val, A


def _scenes():
    from codim_ipc_b200 import scenes
    return [scenes.mixed_small(), scenes.cloth_stack(24, 4), scenes.noodles(4, 40), scenes.granules(2000, cloth_n=15)]


@pytest.mark.parametrize("k", range(4))
def test_csr_of_barrier_and_friction_hessians(ctx, k):
    sc = _scenes()[k]
    n = 3 * len(sc["X"])
    ctx.set_scene(sc)
    ctx.constraint_set(sc["dHat2"], sc["xi"], fetch=False)
    tb = ctx.barrier_hessian(sc["dHat2"], sc["kappa"], sc["xi"], True).copy()
    assert len(tb) > 0
    ctx.csr_begin(); ctx.csr_add()
    rp, ci, v = ctx.csr_finish()
    rp_r, ci_r, v_r, A = _ref_csr([tb], n)
    assert np.array_equal(rp, rp_r) and np.array_equal(ci, ci_r)
    assert np.abs(v - v_r).max() <= 1e-9 * np.abs(v_r).max()
    # the same matrix through scipy: symmetric, and A x agrees
    x = np.cos(0.3 * np.arange(n))
    y = np.zeros(n); np.add.at(y, np.repeat(np.arange(n), np.diff(rp)), v * x[ci])
    assert np.abs(y - A @ x).max() <= 1e-9 * np.abs(A @ x).max()
    # barrier (fused device path) + friction in one matrix
    rng = np.random.default_rng(k)
    ctx.set_prev_positions(sc["X"] - rng.normal(size=sc["X"].shape) * 1e-5)
    ctx.csr_begin()
    nb = ctx.barrier_hessian_dev(sc["dHat2"], sc["kappa"], sc["xi"], True); ctx.csr_add()
    ctx.friction_basis(sc["dHat2"], sc["kappa"], sc["xi"], fetch=False)
    tf = ctx.friction_hessian(1e-10, 0.4, True).copy(); ctx.csr_add()
    rp2, ci2, v2 = ctx.csr_finish()
    rp_r2, ci_r2, v_r2, _ = _ref_csr([tb, tf], n)
    assert nb == len(tb) and np.array_equal(rp2, rp_r2) and np.array_equal(ci2, ci_r2)
    assert np.abs(v2 - v_r2).max() <= 1e-9 * np.abs(v_r2).max()


def test_csr_empty_matrix(ctx):
    from codim_ipc_b200 import scenes
    sc = scenes.cloth_on_sphere(24, draped=False)
    ctx.set_scene(sc)
    ctx.constraint_set(sc["dHat2"], sc["xi"], fetch=False)
    ctx.barrier_hessian(sc["dHat2"], sc["kappa"], sc["xi"], True)
    ctx.csr_begin(); ctx.csr_add()
    rp, ci, v = ctx.csr_finish()
    assert len(ci) == 0 and len(v) == 0 and not rp.any() and len(rp) == 3 * len(sc["X"]) + 1
