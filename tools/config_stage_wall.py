"""Wall time of the device-resident contact stage (bench.py's device_step) on every named configuration, event scopes off."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import codim_ipc_b200 as cipc
from codim_ipc_b200 import scenes

for name in (sys.argv[1:] or ["cfg1", "cfg2", "cfg3", "cfg4_50k", "cfg4_500k", "cfg5_250k", "cfg5_1m"]):
    sc = scenes.CONFIGS[name]()
    a = (sc["dHat2"], sc["kappa"], sc["xi"])
    ctx = cipc.ContactContext(0); ctx.set_scene(sc)
    def step():
        n = ctx.constraint_set(sc["dHat2"], sc["xi"], fetch=False)
        ctx.barrier_energy_dev(*a)
        if n: ctx.barrier_gradient_hessian_dev(*a)
        ctx.step_size_dev(sc["xi"], 1.0)
        ctx.min_dist2_dev(sc["xi"]); ctx.min_dist2_dev(sc["xi"])
        return n
    for _ in range(3): n = step()
    ctx.sync(); ctx.set_timing(False)
    l0 = cipc.kernel_launches(); t0 = time.perf_counter(); K = 20
    for _ in range(K): step()
    ctx.sync()
    print(f"{name}: {1e3 * (time.perf_counter() - t0) / K:.3f} ms per stage, {n} constraints, {(cipc.kernel_launches() - l0) // K} launches", flush=True)
    ctx.close()
