"""Device primitives (scan, radix sort, PSD projection) against numpy."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _u32p(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint32))


@pytest.mark.parametrize("n", [1, 31, 2048, 2049, 100003, 3_000_001])
def test_scan(ctx, n):
    rng = np.random.default_rng(n)
    a = rng.integers(0, 50, n, dtype=np.uint32)
    out = np.zeros(n, np.uint32)
    tot = C.c_uint32(0)
    assert ctx.L.cipc_test_scan(ctx.h, _u32p(a), _u32p(out), C.c_int64(n), C.byref(tot)) == 0
    ref = np.concatenate([[0], np.cumsum(a, dtype=np.uint64)[:-1]]).astype(np.uint32)
    assert np.array_equal(out, ref)
    assert tot.value == int(a.sum())


@pytest.mark.parametrize("n,bits", [(1, 8), (777, 10), (2048, 32), (65537, 17), (2_000_003, 26)])
def test_radix_sort_stable(ctx, n, bits):
    rng = np.random.default_rng(n)
    k = rng.integers(0, 1 << min(bits, 31), n, dtype=np.uint32)
    v = np.arange(n, dtype=np.uint32)
    k2, v2 = k.copy(), v.copy()
    assert ctx.L.cipc_test_sort(ctx.h, _u32p(k2), _u32p(v2), C.c_int64(n), bits) == 0
    order = np.argsort(k, kind="stable")
    assert np.array_equal(k2, k[order])
    assert np.array_equal(v2, v[order])  # stability


@pytest.mark.parametrize("n", [6, 9, 12])
def test_psd_projection(ctx, n):
    rng = np.random.default_rng(n)
    count = 500
    H = rng.normal(size=(count, n, n))
    H = H + H.transpose(0, 2, 1)
    H[:50] = np.einsum("kij,klj->kil", H[:50], H[:50])  # already PSD: must come back unchanged
    # low-rank indefinite blocks like the barrier Hessians (rank 5, two negative eigenvalues)
    for k in range(50, 150):
        U = rng.normal(size=(n, 5))
        H[k] = (U * np.array([3.0, 1.0, 0.5, -2.0, -0.1])) @ U.T
    G = np.ascontiguousarray(H.copy())
    assert ctx.L.cipc_test_make_pd(ctx.h, G.ctypes.data_as(C.POINTER(C.c_double)), n, count) == 0
    for k in range(count):
        w, U = np.linalg.eigh(H[k])
        P = (U * np.maximum(w, 0)) @ U.T
        assert np.linalg.norm(G[k] - P) <= 1e-12 * np.linalg.norm(H[k]), k
