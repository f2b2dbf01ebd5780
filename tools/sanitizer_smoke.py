"""Every stage of the library (constraint set, E, g, fused Hessians in triplet and block mode, device-side merge + merged
delivery, friction, CSR, step size, min-dist, boundary-primitive construction, a multi-device context with two ranks on one
device) on three small scenes; run it under `compute-sanitizer --tool memcheck|racecheck|synccheck python tools/sanitizer_smoke.py` on a GPU box."""
import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np
import codim_ipc_b200 as cipc
from codim_ipc_b200 import scenes
ctx = cipc.ContactContext(0)
for sc in (scenes.mixed_small(), scenes.cloth_stack(12, 3), scenes.granules(800, cloth_n=9)):
    ctx.set_scene(sc)
    n = ctx.constraint_set(sc["dHat2"], sc["xi"], fetch=False)
    ctx.barrier_energy_dev(sc["dHat2"], sc["kappa"], sc["xi"]); ctx.barrier_gradient_dev(sc["dHat2"], sc["kappa"], sc["xi"])
    nt = ctx.barrier_hessian_dev(sc["dHat2"], sc["kappa"], sc["xi"], True)
    ctx.csr_begin(); ctx.csr_add()
    ctx.friction_basis(sc["dHat2"], sc["kappa"], sc["xi"], fetch=False)
    ctx.set_prev_positions(sc["X"] - 1e-5)
    ctx.friction_energy_dev(1e-10, 0.4); ctx.friction_gradient_dev(1e-10, 0.4, True)
    ctx.friction_hessian_dev(1e-10, 0.4, True); ctx.csr_add()
    nnz = ctx.csr_finish(fetch=False)
    t = ctx.barrier_hessian(sc["dHat2"], sc["kappa"], sc["xi"], True)
    a = ctx.step_size(sc["xi"], 1.0)
    ctx.min_dist2_dev(sc["xi"]); ctx.sync()
    ng = ctx.barrier_gradient_hessian_dev(sc["dHat2"], sc["kappa"], sc["xi"])
    m = ctx.barrier_hessian_merged(sc["dHat2"], sc["kappa"], sc["xi"], True)
    mu = ctx.barrier_hessian_merged(sc["dHat2"], sc["kappa"], sc["xi"], False)
    mf = ctx.friction_hessian_merged(1e-10, 0.4, True)
    b = ctx.build_boundary(sc["X"], sc["F"], rod=sc["rodE"], rodRadius=np.full(len(sc["rodE"]), 1e-3), particle=sc["particles"])
    print(sc.get("name"), n, nt, nnz, len(t), a, ng, len(m), len(mu), len(mf), len(b["BE"]))
M = cipc.ContactContext(devices=[0, 0])
sc = scenes.mixed_small()
M.set_scene(sc)
cs, info = M.constraint_set(sc["dHat2"], sc["xi"])
print("multi", len(cs), M.barrier_energy(sc["dHat2"], sc["kappa"], sc["xi"]), np.abs(M.barrier_gradient(sc["dHat2"], sc["kappa"], sc["xi"])).max(),
      len(M.barrier_hessian_merged(sc["dHat2"], sc["kappa"], sc["xi"], True)), M.step_size(sc["xi"], 1.0), M.min_dist2(sc["xi"])[1])
M.close()
print("sanitizer script done")
