import time, sys
sys.path.insert(0, ".")
import numpy as np
import codim_ipc_b200 as cipc
from codim_ipc_b200 import scenes
from oracle import cipc_oracle as O
sc = scenes.cloth_stack(224, 10)
ctx = cipc.ContactContext(0)
for i in range(3):
    X = sc["X"] * (1 + 0.01 * i)
    t = time.perf_counter(); g = ctx.build_boundary(X, sc["F"]); t1 = time.perf_counter() - t
    print("gpu build_boundary 1M tris: %.2f ms (device stage %.2f ms)" % (1e3 * t1, ctx.stage_ms("build_boundary")))
t = time.perf_counter(); o = O.build_boundary(sc["X"], sc["F"]); print("literal std::map restatement: %.1f ms" % (1e3 * (time.perf_counter() - t)))
