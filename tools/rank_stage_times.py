"""Fixed cost of one rank's share of the stage: rank 0 of world W on ONE device (no collectives), cfg5_1m.  Prints the wall time
of the device-resident stage with the event scopes off, the per-scope device times, and the number of kernel launches --
what bounds the N = 8 line of the scaling run (launch / round-trip bound)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import codim_ipc_b200 as cipc
from codim_ipc_b200 import scenes

n, layers = (224, 10) if len(sys.argv) < 3 else (int(sys.argv[1]), int(sys.argv[2]))
sc = scenes.cloth_stack(n, layers)
a = (sc["dHat2"], sc["kappa"], sc["xi"])
for world in (1, 2, 4, 8):
    for rank in sorted({0, world // 2}):
        ctx = cipc.ContactContext(0, rank, world); ctx.set_scene(sc)
        def step():
            ctx.constraint_set(sc["dHat2"], sc["xi"], fetch=False)
            ctx.barrier_energy_dev(*a)
            ctx.barrier_gradient_hessian_dev(*a)
            ctx.step_size_dev(sc["xi"], 1.0)
            ctx.min_dist2_dev(sc["xi"]); ctx.min_dist2_dev(sc["xi"])
        for _ in range(3): step()
        ctx.sync()
        st = {}
        def grab(*names):
            for k in names: st[k] = round(ctx.stage_ms(k), 3)
        ctx.constraint_set(sc["dHat2"], sc["xi"], fetch=False); grab("ccs_hash_build", "ccs_pairs", "ccs_narrow", "ccs_merge")
        ctx.barrier_energy_dev(*a); grab("barrier_E")
        ctx.barrier_gradient_hessian_dev(*a); grab("barrier_H")
        ctx.step_size_dev(sc["xi"], 1.0); grab("ccd_hash_build", "ccd_pairs", "ccd_accd")
        ctx.min_dist2_dev(sc["xi"]); grab("min_dist")
        ctx.set_timing(False)
        l0 = cipc.kernel_launches()
        t0 = time.perf_counter()
        K = 10
        for _ in range(K): step()
        ctx.sync()
        wall = (time.perf_counter() - t0) / K * 1e3
        print(f"world {world} rank {rank}: wall {wall:.3f} ms  launches/stage {(cipc.kernel_launches() - l0) // K}  sum(scopes) {sum(st.values()):.3f}  {st}", flush=True)
        ctx.close()
