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


def test_friction_shim_templates_through_reference_call_pattern(tmp_path):
    """codim-ipc_b200/shim/FEM/FRICTION.h driven like Shell/IMPLICIT_EULER.h:419-464 drives the reference's FRICTION.h"""
    from codim_ipc_b200 import scenes
    from oracle import cipc_oracle as O
    sc = scenes.mixed_small()
    exe = str(tmp_path / "friction_harness")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-I", os.path.join(ROOT, "codim-ipc_b200", "shim"), "-I", os.path.join(ROOT, "include"),
                           "-I", os.path.join(ROOT, "tests", "shim_harness", "stub"), "-o", exe,
                           os.path.join(ROOT, "tests", "shim_harness", "friction_main.cpp"),
                           "-L", os.path.join(ROOT, "codim-ipc_b200"), "-lcipc_b200", "-Wl,-rpath," + os.path.join(ROOT, "codim-ipc_b200")])
    nV = len(sc["X"])
    rng = np.random.default_rng(21)
    Xn = sc["X"] - rng.normal(size=sc["X"].shape) * np.where(rng.random(nV) < 0.5, 1e-7, 1e-4)[:, None]
    with open(tmp_path / "scene.bin", "wb") as f:
        f.write(np.array([nV, len(sc["BN"]), len(sc["BE"]), len(sc["BT"]), sc["nRod"], sc["codim"][0], sc["codim"][1], len(sc["NNX"])], np.int32).tobytes())
        f.write(np.array([sc["dHat2"], sc["xi"], *sc["kappa"], 1.0], np.float64).tobytes())
        for k in ("X", "X0", "p"):
            f.write(np.ascontiguousarray(sc[k], np.float64).tobytes())
        f.write(np.ascontiguousarray(sc["BN"], np.int32).tobytes()); f.write(np.ascontiguousarray(sc["BE"], np.int32).tobytes())
        f.write(np.ascontiguousarray(sc["BT"], np.int32).tobytes()); f.write(np.ascontiguousarray(sc["DBC"], np.uint8).tobytes())
        f.write(np.ascontiguousarray(sc["NNX"], np.int32).tobytes())
        f.write(np.ascontiguousarray(Xn, np.float64).tobytes())
    r = subprocess.run([exe, str(tmp_path / "scene.bin"), str(tmp_path / "out.bin")], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    buf = open(tmp_path / "out.bin", "rb").read()
    n, nF, nt = struct.unpack_from("qqq", buf, 0)
    off = 24
    cs = np.frombuffer(buf, np.int32, 4 * n, off).reshape(n, 4); off += 16 * n
    info = np.frombuffer(buf, np.float64, 2 * n, off).reshape(n, 2); off += 16 * n
    fcs = np.frombuffer(buf, np.int32, 4 * nF, off).reshape(nF, 4); off += 16 * nF
    cp = np.frombuffer(buf, np.float64, 2 * nF, off).reshape(nF, 2); off += 16 * nF
    B = np.frombuffer(buf, np.float64, 6 * nF, off).reshape(nF, 6); off += 48 * nF
    nf0 = np.frombuffer(buf, np.float64, nF, off); off += 8 * nF
    nf = np.frombuffer(buf, np.float64, nF, off); off += 8 * nF
    E, E2, mu = np.frombuffer(buf, np.float64, 3, off); off += 24
    g = np.frombuffer(buf, np.float64, 3 * nV, off).reshape(nV, 3); off += 24 * nV
    trip = np.frombuffer(buf, np.dtype([("r", np.int32), ("c", np.int32), ("v", np.float64)]), nt, off)

    S = O.OracleScene(sc)
    fo, cpo, Bo, nfo = S.friction_basis(cs, info, sc["dHat2"], sc["kappa"], sc["xi"])
    close = lambda a, b: np.abs(a - b).max() <= 1e-9 * np.abs(b).max()
    assert nF > 0 and np.array_equal(fcs, fo) and close(cp, cpo) and close(B, Bo) and close(nf0, nfo)
    S.friction_coef(np.array([nV // 2, nV], np.int32), np.array([0.3, 0.7, 0.5, 0.2]))
    assert mu == 1.0 and close(nf, S.fetch_friction(nF)[3])
    E_o = S.friction_potential(Xn, 1e-10, mu)
    assert abs((E - 2.0) - E_o) <= 1e-9 * abs(E_o) and abs(E2 - E_o) <= 1e-9 * abs(E_o)
    g_o = S.friction_gradient(Xn, 1e-10, mu)
    assert np.abs((g - 0.5) - g_o).max() <= 1e-9 * np.abs(g_o).max()
    r_o, c_o, v_o = S.friction_hessian(Xn, 1e-10, mu, True)
    assert nt == 5 + len(v_o) and np.all(trip["v"][:5] == 3.0)
    assert np.array_equal(trip["r"][5:], r_o) and np.array_equal(trip["c"][5:], c_o)
    assert max_block_rel_err(fcs, trip["v"][5:], v_o) <= 1e-9
