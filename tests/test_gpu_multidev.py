"""Several GPUs behind ONE context and one calling thread (cipc_create_multi, multidev.h): results equal the single-device
context's -- the constraint set as a sorted set (PP / PE multiplicities merged across ranks), E and g to summation order,
the Hessian as an assembled matrix, the step size and dist2 exactly.  The rank logic (slab partition, peer gathers, re-balanced
contiguous chunks, per-rank deliveries) is also exercised on a one-GPU box by placing two and three ranks on device 0."""
import numpy as np
import pytest
import scipy.sparse as sp

from helpers import sort_cs

pytestmark = pytest.mark.gpu


def _ndev():
    import torch
    return torch.cuda.device_count()


def _device_lists():
    n = _ndev()
    lists = [[0], [0, 0], [0, 0, 0]]
    if n >= 2:
        lists.append(list(range(min(n, 2))))
    if n >= 4:
        lists.append(list(range(4)))
    if n >= 8:
        lists.append(list(range(8)))
    return lists


def _cases():
    from codim_ipc_b200 import scenes
    return {
        "stack_48x6": lambda: scenes.cloth_stack(48, 6),
        "mixed_small": lambda: scenes.mixed_small(),
        "noodles_8x80": lambda: scenes.noodles(8, 80),
        "granules_6k": lambda: scenes.granules(6000, cloth_n=25),
    }


def _csr(t, n):
    return sp.coo_matrix((t["val"], (t["row"], t["col"])), shape=(n, n)).tocsr()


@pytest.mark.parametrize("name", list(_cases()))
def test_multi_context_equals_single_device(ctx, name):
    import codim_ipc_b200 as cipc
    sc = _cases()[name]()
    n3 = 3 * len(sc["X"])
    ctx.set_scene(sc)
    cs1, info1 = sort_cs(*ctx.constraint_set(sc["dHat2"], sc["xi"]))
    a = (sc["dHat2"], sc["kappa"], sc["xi"])
    ctx.set_constraints(cs1, info1)
    E1 = ctx.barrier_energy(*a, E=0.5)
    g1 = ctx.barrier_gradient(*a)
    H1 = _csr(ctx.barrier_hessian(*a, True), n3)
    d1, m1 = ctx.min_dist2(sc["xi"])
    s1 = ctx.step_size(sc["xi"], 1.0)
    for devs in _device_lists():
        M = cipc.ContactContext(devices=devs)
        try:
            M.set_scene(sc)
            cs, info = M.constraint_set(sc["dHat2"], sc["xi"])
            assert np.array_equal(sort_cs(cs), cs1), (name, devs)
            assert np.all(info[:, 0] == 1.0) and np.all(info[:, 1] == info1[0, 1])
            # the resident, re-balanced per-rank chunks
            E = M.barrier_energy(*a, E=0.5)
            assert abs(E - E1) <= 1e-12 * abs(E1)
            g = M.barrier_gradient(*a)
            assert np.abs(g - g1).max() <= 1e-12 * np.abs(g1).max()
            for merged in (True, False):
                t = (M.barrier_hessian_merged if merged else M.barrier_hessian)(*a, True)
                H = _csr(t, n3)
                assert abs(H - H1).max() <= 1e-9 * abs(H1).max() and H.nnz == H1.nnz, (name, devs, merged)
            d, mm = M.min_dist2(sc["xi"])
            assert len(d) == len(cs) and mm == m1
            assert np.array_equal(np.sort(d), np.sort(d1))
            assert M.step_size(sc["xi"], 1.0) == s1
            # a caller-owned set, split evenly across the ranks: outputs line up with the caller's order
            M.set_constraints(cs1, info1)
            d2, m2 = M.min_dist2(sc["xi"])
            assert np.array_equal(d2, d1) and m2 == m1
            t = M.barrier_hessian(*a, True)
            t1 = ctx.barrier_hessian(*a, True)
            assert np.array_equal(t["row"], t1["row"]) and np.array_equal(t["col"], t1["col"]) and np.array_equal(t["val"], t1["val"])
            assert abs(M.barrier_energy(*a) - (E1 - 0.5)) <= 1e-12 * abs(E1)
        finally:
            M.close()


def test_multi_context_rejects_single_device_only_calls():
    import codim_ipc_b200 as cipc
    from codim_ipc_b200 import scenes
    sc = scenes.cloth_stack(12, 3)
    M = cipc.ContactContext(devices=[0, 0])
    try:
        M.set_scene(sc)
        with pytest.raises(cipc.CipcError):
            M.barrier_energy_dev(sc["dHat2"], sc["kappa"], sc["xi"])
        with pytest.raises(cipc.CipcError):
            M.csr_begin()
    finally:
        M.close()
