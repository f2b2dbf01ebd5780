"""One contact stage for profiling (ncu): the device-resident stage bench.py times plus the merged Hessian path of the shim.
Usage: profile_stage.py [n layers] (default cfg5_1m = 224 10)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import codim_ipc_b200 as cipc
from codim_ipc_b200 import scenes
n, L = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (224, 10)
sc = scenes.cloth_stack(n, L)
ctx = cipc.ContactContext(0)
ctx.set_scene(sc)
a = (sc["dHat2"], sc["kappa"], sc["xi"])
for rep in range(2):  # the second pass is the one to read (buffers allocated)
    ctx.constraint_set(sc["dHat2"], sc["xi"], fetch=False)
    ctx.barrier_energy_dev(*a)
    ctx.barrier_gradient_hessian_dev(*a)
    ctx.barrier_hessian_merged(*a, True, fetch=False)
    ctx.step_size_dev(sc["xi"], 1.0)
    ctx.min_dist2_dev(sc["xi"]); ctx.min_dist2_dev(sc["xi"])
ctx.sync()
print("profiled stage done", ctx.counter("constraints"))
