"""Device / delivery times of the merged-triplet Hessian path on one scene (default cfg5_1m)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import codim_ipc_b200 as cipc
from codim_ipc_b200 import scenes
n, L = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (224, 10)
sc = scenes.cloth_stack(n, L)
ctx = cipc.ContactContext(0)
ctx.set_scene(sc)
nC = ctx.constraint_set(sc["dHat2"], sc["xi"], fetch=False)
a = (sc["dHat2"], sc["kappa"], sc["xi"])
for rep in range(3):
    nT = ctx.barrier_hessian_merged(*a, True, fetch=False)
    print("constraints", nC, "merged triplets", nT, {k: round(ctx.stage_ms(k), 3) for k in ("barrier_H", "k_hessian_fused0", "k_hessian_fused12", "hessian_merge", "mg_count", "mg_scatter", "mg_sort", "mg_scan", "mg_uniq", "mg_sum")},
          {k: ctx.counter(k) for k in ("merge_blocks_in", "merge_blocks_unique", "merge_blocks_diag")})
buf = torch.empty((nT, 2), dtype=torch.float64).pin_memory().numpy().view(cipc.TRIPLET_DTYPE).reshape(-1)
pag = np.empty(nT, cipc.TRIPLET_DTYPE)
for name, out in (("pinned", buf), ("pageable", pag), ("pageable", pag)):
    for rep in range(3):
        t0 = time.perf_counter()
        ctx.barrier_hessian_merged(*a, True, out=out)
        print(name, "hessian_merged + delivery ms", round(1e3 * (time.perf_counter() - t0), 2), {k: ctx.counter(k) for k in ("deliver_us_keys", "deliver_us_count", "deliver_us_copied", "deliver_us_total")})
for rep in range(2):
    t0 = time.perf_counter()
    fresh = np.empty(nT, cipc.TRIPLET_DTYPE)
    ctx.barrier_hessian_merged(*a, True, out=fresh)
    print("fresh pageable buffer (first touch inside)", round(1e3 * (time.perf_counter() - t0), 2))
    del fresh
