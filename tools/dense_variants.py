"""Times the barrier-Hessian stage (dense path included) of experimental builds of the library (CIPC_LIB) on cfg3 and cfg5_1m."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = r'''
import sys; sys.path.insert(0, %r)
import codim_ipc_b200 as cipc
from codim_ipc_b200 import scenes
for name in ("cfg3", "cfg5_1m"):
    sc = scenes.CONFIGS[name]() if name in scenes.CONFIGS else scenes.cloth_stack(224, 10)
    ctx = cipc.ContactContext(0); ctx.set_scene(sc)
    ctx.constraint_set(sc["dHat2"], sc["xi"], fetch=False)
    a = (sc["dHat2"], sc["kappa"], sc["xi"])
    best = 1e9
    for rep in range(5):
        ctx.barrier_gradient_hessian_dev(*a); best = min(best, ctx.stage_ms("barrier_H"))
    print(name, "gH stage %%.3f ms, mollified %%d" %% (best, ctx.counter("hessian_mollified")), end="; ")
    ctx.close()
print()
''' % ROOT
for lib in sys.argv[1:]:
    env = dict(os.environ, CIPC_LIB=os.path.abspath(lib))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env)
    print(os.path.basename(lib), out.stdout.strip().splitlines()[-1] if out.stdout.strip() else out.stderr[-800:])
