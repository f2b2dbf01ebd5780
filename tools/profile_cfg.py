"""Barrier Hessian of a named configuration, twice (for ncu): profile_cfg.py cfg3"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import codim_ipc_b200 as cipc
from codim_ipc_b200 import scenes
sc = scenes.CONFIGS[sys.argv[1]]()
ctx = cipc.ContactContext(0); ctx.set_scene(sc)
a = (sc["dHat2"], sc["kappa"], sc["xi"])
ctx.constraint_set(sc["dHat2"], sc["xi"], fetch=False)
for rep in range(2):
    ctx.barrier_gradient_hessian_dev(*a)
ctx.sync()
print("done", ctx.counter("constraints"), ctx.counter("hessian_mollified"))
