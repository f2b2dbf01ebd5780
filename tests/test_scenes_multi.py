"""Host logic: boundary-primitive extraction semantics, scene sizes, multi-rank merge, and the
world_size-2 gloo path of the collectives used at N > 1."""
import os
import socket
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _surface_primitives_loop(nV, F):
    """literal restatement of Find_Surface_Primitives (Library/Utils/MESHIO.h:728-766) with Python containers"""
    edges = set()
    order = []
    for t in F:
        for a, b in ((t[0], t[1]), (t[1], t[2]), (t[2], t[0])):
            if (b, a) not in edges:
                if (a, b) not in edges:
                    order.append((a, b))
                edges.add((a, b))
    return sorted(edges)


def test_surface_primitives_match_reference_semantics():
    from codim_ipc_b200 import scenes
    V, F = scenes.grid_mesh(7)
    rng = np.random.default_rng(0)
    F = F[rng.permutation(len(F))]  # arbitrary triangle order
    BN, BE, BT, na, ea, ta = scenes.surface_primitives(V, F)
    assert [tuple(e) for e in BE] == _surface_primitives_loop(len(V), F)
    assert np.array_equal(BT, F) and np.array_equal(BN, np.arange(len(V)))
    assert len(BE) == 3 * 49 + 2 * 7  # Euler: E = 3 n^2 + 2 n for an n x n grid
    assert np.isclose(na.sum(), 1.0) and np.isclose(ta.sum(), 0.5) and np.isclose(ea.sum(), 0.5)


def test_config_sizes():
    from codim_ipc_b200 import scenes
    sc = scenes.cloth_on_sphere(112)
    assert len(sc["BT"]) == 25088 + 2 * 36 * 35 + 800  # square113-like cloth + sphere + floor
    sc = scenes.cloth_stack(8, 3)
    assert len(sc["BT"]) == 3 * 128 and sc["codim"] == (len(sc["BN"]), len(sc["BN"])) and sc["nRod"] == 0
    sc = scenes.mixed_small()
    assert sc["nRod"] == 80 and sc["codim"][0] < sc["codim"][1] < len(sc["BN"]) and sc["DBC"].sum() > 0


def test_merge_constraint_sets_adds_multiplicities():
    from codim_ipc_b200 import multi
    a = np.array([[-5, 7, 8, 9], [3, 4, 5, 6], [-2, 9, -1, -2], [-3, 1, 2, -1]], np.int32)
    b = np.array([[-2, 9, -1, -1], [-3, 1, 2, -3], [1, 2, -4, 5], [-9, 9, -1, -1]], np.int32)
    m = multi.merge_constraint_sets([a, b])
    got = {tuple(r) for r in m}
    assert got == {(-5, 7, 8, 9), (3, 4, 5, 6), (1, 2, -4, 5), (-2, 9, -1, -3), (-3, 1, 2, -4), (-9, 9, -1, -1)}


_WORKER = r"""
import os, sys
import numpy as np
sys.path.insert(0, %r)
import torch.distributed as dist
from codim_ipc_b200 import multi
dist.init_process_group("gloo")
r, w = dist.get_rank(), dist.get_world_size()
dc = multi.DistContact()
assert dc.min_step(0.5 + 0.25 * r) == 0.5
assert dc.sum_scalar(1.0 + r) == 3.0
g = dc.sum_gradient(np.full((5, 3), float(r + 1)))
assert np.all(g == 3.0)
parts = [np.array([[-2, 9, -1, -1], [3, 4, 5, 6]], np.int32), np.array([[-2, 9, -1, -2], [-7, 1, 2, 3], [-8, 1, -1, -1]], np.int32)]
m = dc.gather_constraints(parts[r])
got = {tuple(x) for x in m}
assert got == {(3, 4, 5, 6), (-7, 1, 2, 3), (-2, 9, -1, -3), (-8, 1, -1, -1)}, got
T = np.dtype([("row", np.int32), ("col", np.int32), ("val", np.float64)])
mine = np.zeros(3 + 2 * r, T); mine["row"] = 100 * r + np.arange(len(mine)); mine["val"] = r + 0.5
off, total, counts = dc.triplet_offsets(len(mine))
assert counts == [3, 5] and total == 8 and off == (0 if r == 0 else 3)
full = dc.gather_triplets(mine)
assert len(full) == 8 and list(full["row"]) == [0, 1, 2, 100, 101, 102, 103, 104] and np.all(full["val"][:3] == 0.5) and np.all(full["val"][3:] == 1.5)
dist.destroy_process_group()
print("rank", r, "ok")
"""


def test_gloo_world2_collectives(tmp_path):
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    script = tmp_path / "worker.py"
    script.write_text(_WORKER % ROOT)
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                          "--master-port", str(port), str(script)], capture_output=True, text=True, timeout=240)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.count("ok") == 2
