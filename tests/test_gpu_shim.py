"""Drop-in boundary: codim-ipc_b200/shim/FEM/IPC.h compiled against stand-ins of the reference's container
types (tests/shim_harness/stub) and driven like the reference's time stepper; results vs the oracle."""
import os
import struct
import subprocess

import numpy as np
import pytest

from helpers import sort_cs, max_block_rel_err

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shim_templates_through_reference_call_pattern(tmp_path):
    from codim_ipc_b200 import scenes
    from oracle import cipc_oracle as O
    sc = scenes.mixed_small()
    exe = str(tmp_path / "shim_harness")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-I", os.path.join(ROOT, "codim-ipc_b200", "shim"), "-I", os.path.join(ROOT, "include"),
                           "-I", os.path.join(ROOT, "tests", "shim_harness", "stub"), "-o", exe, os.path.join(ROOT, "tests", "shim_harness", "main.cpp"),
                           "-L", os.path.join(ROOT, "codim-ipc_b200"), "-lcipc_b200", "-Wl,-rpath," + os.path.join(ROOT, "codim-ipc_b200")])
    nV = len(sc["X"])
    a0 = 0.8
    with open(tmp_path / "scene.bin", "wb") as f:
        f.write(np.array([nV, len(sc["BN"]), len(sc["BE"]), len(sc["BT"]), sc["nRod"], sc["codim"][0], sc["codim"][1], len(sc["NNX"])], np.int32).tobytes())
        f.write(np.array([sc["dHat2"], sc["xi"], *sc["kappa"], a0], np.float64).tobytes())
        for k in ("X", "X0", "p"):
            f.write(np.ascontiguousarray(sc[k], np.float64).tobytes())
        f.write(np.ascontiguousarray(sc["BN"], np.int32).tobytes()); f.write(np.ascontiguousarray(sc["BE"], np.int32).tobytes())
        f.write(np.ascontiguousarray(sc["BT"], np.int32).tobytes()); f.write(np.ascontiguousarray(sc["DBC"], np.uint8).tobytes())
        f.write(np.ascontiguousarray(sc["NNX"], np.int32).tobytes())
    r = subprocess.run([exe, str(tmp_path / "scene.bin"), str(tmp_path / "out.bin")], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    buf = open(tmp_path / "out.bin", "rb").read()
    n, nt = struct.unpack_from("qq", buf, 0)
    off = 16
    cs = np.frombuffer(buf, np.int32, 4 * n, off).reshape(n, 4); off += 16 * n
    info = np.frombuffer(buf, np.float64, 2 * n, off).reshape(n, 2); off += 16 * n
    E, step, mind, E2 = np.frombuffer(buf, np.float64, 4, off); off += 32
    g = np.frombuffer(buf, np.float64, 3 * nV, off).reshape(nV, 3); off += 24 * nV
    trip = np.frombuffer(buf, np.dtype([("r", np.int32), ("c", np.int32), ("v", np.float64)]), nt, off); off += 16 * nt
    dist2 = np.frombuffer(buf, np.float64, n, off)

    S = O.OracleScene(sc)
    cs_o, info_o = S.constraint_set(sc["dHat2"], sc["xi"])
    assert np.array_equal(sort_cs(cs), sort_cs(cs_o))
    E_o = S.barrier(cs, info, sc["dHat2"], sc["kappa"], sc["xi"])
    assert abs((E - 2.0) - E_o) <= 1e-9 * abs(E_o) and abs(E2 - E_o) <= 1e-9 * abs(E_o)
    g_o = S.barrier_gradient(cs, info, sc["dHat2"], sc["kappa"], sc["xi"])
    assert np.abs((g - 0.5) - g_o).max() <= 1e-9 * np.abs(g_o).max()
    r_o, c_o, v_o = S.barrier_hessian(cs, info, sc["dHat2"], sc["kappa"], sc["xi"], True)
    assert nt == 5 + len(v_o) and np.all(trip["v"][:5] == 3.0)
    assert np.array_equal(trip["r"][5:], r_o) and np.array_equal(trip["c"][5:], c_o)
    assert max_block_rel_err(cs, trip["v"][5:], v_o) <= 1e-9
    a_o = S.step_size(sc["p"], sc["xi"], a0)
    assert step <= a_o and a_o - step <= 1e-12 * a_o
    d_o, m_o = S.min_dist2(cs, sc["xi"])
    assert np.array_equal(dist2, d_o) and mind == m_o
