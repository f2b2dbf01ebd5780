"""Times the fused gradient+Hessian pass of experimental builds of the library (CIPC_LIB) on cfg5_1m."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = r'''
import sys; sys.path.insert(0, %r)
import codim_ipc_b200 as cipc
from codim_ipc_b200 import scenes
sc = scenes.cloth_stack(224, 10)
ctx = cipc.ContactContext(0); ctx.set_scene(sc)
ctx.constraint_set(sc["dHat2"], sc["xi"], fetch=False)
a = (sc["dHat2"], sc["kappa"], sc["xi"])
best = [1e9, 1e9, 1e9, 1e9]; bpe = 1e9
for rep in range(6):
    ctx.barrier_gradient_hessian_dev(*a); t = [ctx.stage_ms("barrier_H"), ctx.stage_ms("k_hessian_fused0")]; pe = ctx.stage_ms("k_hessian_fused12")
    ctx.barrier_hessian_merged(*a, True, fetch=False); t += [ctx.stage_ms("k_hessian_fused0"), ctx.stage_ms("barrier_H")]
    best = [min(x, y) for x, y in zip(best, t)]; bpe = min(bpe, pe)
print("gH stage %%.3f  fused0(triplets+grad) %%.3f  fused0(blocks) %%.3f  merged stage %%.3f  fused12(triplets+grad) %%.3f" %% tuple(best + [bpe]))
''' % ROOT
for lib in sys.argv[1:]:
    env = dict(os.environ, CIPC_LIB=os.path.abspath(lib))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env)
    print(os.path.basename(lib), out.stdout.strip().splitlines()[-1] if out.stdout.strip() else out.stderr[-500:])
