"""Per-stage device times of one contact stage on a named configuration (diagnostics, not the benchmark)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import codim_ipc_b200 as cipc
from codim_ipc_b200 import scenes

name = sys.argv[1]
sc = scenes.CONFIGS[name]()
ctx = cipc.ContactContext(0)
ctx.set_scene(sc)
for rep in range(2):
    out = {}
    n = ctx.constraint_set(sc["dHat2"], sc["xi"], fetch=False)
    for s in ("ccs_hash_build", "ccs_pairs", "ccs_narrow", "ccs_merge"):
        out[s] = ctx.stage_ms(s)
    hb = {s: round(ctx.stage_ms(s), 3) for s in ("hb_count_scan", "hb_emit", "hb_sort", "hb_heads", "hb_tables")}
    cnt = {k: ctx.counter(k) for k in ("hash_entries", "hash_cells", "candidates_pt", "candidates_ee", "candidates_pe", "candidates_pp", "constraints")}
    ctx.barrier_energy_dev(sc["dHat2"], sc["kappa"], sc["xi"]); out["E"] = ctx.stage_ms("barrier_E")
    ctx.barrier_gradient_dev(sc["dHat2"], sc["kappa"], sc["xi"]); out["g"] = ctx.stage_ms("barrier_g")
    nT = ctx.barrier_hessian(sc["dHat2"], sc["kappa"], sc["xi"], True, fetch=False); out["H_factor"] = ctx.stage_ms("barrier_H")
    t0 = time.perf_counter(); ctx.dev_triplets(); ctx.sync(); out["H_expand_wall"] = 1e3 * (time.perf_counter() - t0)
    ctx.barrier_hessian_dev(sc["dHat2"], sc["kappa"], sc["xi"], True); out["H_fused"] = ctx.stage_ms("barrier_H")
    cnt.update({k: ctx.counter(k) for k in ("hessian_4pt", "hessian_pe", "hessian_pp", "hessian_mollified")})
    ctx.step_size_dev(sc["xi"], 1.0)
    for s in ("ccd_hash_build", "ccd_pairs", "ccd_accd"):
        out[s] = ctx.stage_ms(s)
    cnt["ccd_pairs"] = ctx.counter("ccd_pairs")
    ctx.min_dist2_dev(sc["xi"]); out["min_dist"] = ctx.stage_ms("min_dist")
print(name, "nV", len(sc["X"]), "tris", len(sc["BT"]), "edges", len(sc["BE"]))
print({k: round(v, 3) for k, v in out.items()}, "total", round(sum(out.values()), 2))
print("ccs hash build detail", hb)
print(cnt)
