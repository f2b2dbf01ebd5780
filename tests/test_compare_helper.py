"""The native triplet-stream comparator used by the full-size parity gates (oracle.compare_triplet_blocks): checked here on
the CPU by comparing the oracle port's Hessian with the reference drivers' and with a perturbed copy."""
import os

import numpy as np
import pytest

from oracle import cipc_oracle as O

HAVE_REF = os.path.exists(os.path.join(os.path.dirname(O.__file__), "_ref", "libcipc_refdrv.so"))


def _scene():
    from codim_ipc_b200 import scenes
    return scenes.cloth_stack(16, 3)


def test_comparator_flags_differences():
    sc = _scene()
    S = O.OracleScene(sc)
    cs, info = S.constraint_set(sc["dHat2"], sc["xi"])
    n = S.barrier_hessian_notfetch(cs, info, sc["dHat2"], sc["kappa"], sc["xi"], True)
    p, cnt = S.triplets_data()
    assert cnt == n > 0
    a = np.ctypeslib.as_array((np.ctypeslib.ctypes.c_uint8 * (16 * n)).from_address(p)).copy()
    t = a.view(np.dtype([("row", np.int32), ("col", np.int32), ("val", np.float64)]))
    same = O.compare_triplet_blocks(t.ctypes.data, p, cs)
    assert same["max_block_rel_err"] == 0 and same["index_mismatches"] == 0 and same["triplets"] == n and same["blocks"] == len(cs)
    t["val"][int(np.argmax(np.abs(t["val"][:36])))] *= 1 + 1e-6
    t["row"][200] += 1
    bad = O.compare_triplet_blocks(t.ctypes.data, p, cs)
    assert bad["index_mismatches"] == 1 and 1e-9 < bad["max_block_rel_err"] < 1e-5


@pytest.mark.skipif(not HAVE_REF, reason="oracle/_ref not built")
def test_oracle_port_matches_reference_drivers_blockwise():
    sc = _scene()
    S, R = O.OracleScene(sc), O.RefScene(sc)
    cs, info = R.constraint_set(sc["dHat2"], sc["xi"])
    n1 = S.barrier_hessian_notfetch(cs, info, sc["dHat2"], sc["kappa"], sc["xi"], True)
    n2 = R.barrier_hessian_notfetch(cs, info, sc["dHat2"], sc["kappa"], sc["xi"], True)
    assert n1 == n2
    c = O.compare_triplet_blocks(S.triplets_data()[0], R.triplets_data()[0], cs)
    assert c["index_mismatches"] == 0 and c["max_block_rel_err"] <= 1e-10
    R.release_triplets()
