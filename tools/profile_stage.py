"""One contact stage for profiling (ncu): the device-resident stage bench.py times plus the merged Hessian path of the shim.
Usage: profile_stage.py [n layers] (default cfg5_1m = 224 10)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import codim_ipc_b200 as cipc
from codim_ipc_b200 import scenes
n, L = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (224, 10)
sc = scenes.cloth_stack(n, L)
rank, world = int(os.environ.get("CIPC_PROFILE_RANK", "0")), int(os.environ.get("CIPC_PROFILE_WORLD", "1"))  # one rank's share on one device
ctx = cipc.ContactContext(0, rank, world)
ctx.set_scene(sc)
a = (sc["dHat2"], sc["kappa"], sc["xi"])
for rep in range(3 if world > 1 else 2):  # the last pass is the one to read (buffers allocated, slab bounds cached)
    ctx.constraint_set(sc["dHat2"], sc["xi"], fetch=False)
    ctx.barrier_energy_dev(*a)
    ctx.barrier_gradient_hessian_dev(*a)
    if world == 1:
        ctx.barrier_hessian_merged(*a, True, fetch=False)
    ctx.step_size_dev(sc["xi"], 1.0)
    ctx.min_dist2_dev(sc["xi"]); ctx.min_dist2_dev(sc["xi"])
ctx.sync()
print("profiled stage done", ctx.counter("constraints"))
