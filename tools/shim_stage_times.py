"""Per-call host wall times of one contact stage through the COMPILED shim harness (tests/shim_harness) on cfg5 (default 1M)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import shim_scene
from codim_ipc_b200 import scenes
n, L = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (224, 10)
sc = scenes.cloth_stack(n, L)
S = shim_scene.ShimScene(sc)
for i in range(6):
    tm, rs = S.contact_stage(sc)
    print("stage %d: total %.2f ms " % (i, 1e3 * sum(tm.values())), {k: round(1e3 * v, 2) for k, v in tm.items()}, rs["nC"], rs["nTriplets"])
