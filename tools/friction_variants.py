"""Times the fused friction Hessian of experimental builds of the library (CIPC_LIB) on cfg5_1m."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = r'''
import sys; sys.path.insert(0, %r)
import numpy as np
import codim_ipc_b200 as cipc
from codim_ipc_b200 import scenes
sc = scenes.cloth_stack(224, 10)
ctx = cipc.ContactContext(0); ctx.set_scene(sc)
ctx.constraint_set(sc["dHat2"], sc["xi"], fetch=False)
rng = np.random.default_rng(17)
ctx.set_prev_positions(sc["X"] - rng.normal(size=sc["X"].shape) * 2e-5)
ctx.friction_basis(sc["dHat2"], sc["kappa"], sc["xi"], fetch=False)
best = [1e9, 1e9]
for rep in range(5):
    ctx.friction_hessian_dev(1e-10, 0.4, True); t = [ctx.stage_ms("friction_H")]
    ctx.friction_hessian_merged(1e-10, 0.4, True, fetch=False); t.append(ctx.stage_ms("friction_H"))
    best = [min(a, b) for a, b in zip(best, t)]
print("friction H fused (triplets) %%.3f ms, (blocks + merge) %%.3f ms" %% tuple(best))
''' % ROOT
for lib in sys.argv[1:]:
    env = dict(os.environ, CIPC_LIB=os.path.abspath(lib))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env)
    print(os.path.basename(lib), out.stdout.strip().splitlines()[-1] if out.stdout.strip() else out.stderr[-800:])
