#!/usr/bin/env python
"""bench.py -- contact-stage benchmark of the B200 C-IPC hot path (BASELINE.json metric).

A "step" is one contact stage of a Newton iteration on one synthetic scene:
    1x Compute_Constraint_Set + 1x Compute_Barrier + 1x Compute_Barrier_Gradient
    + 1x Compute_Barrier_Hessian (PSD-projected) + 1x Compute_Intersection_Free_StepSize + 2x Compute_Min_Dist2
Workload: cfg5, the 1M-triangle stacked-cloth scene the BASELINE target is quoted on
(10 layers x 224 x 224 x 2 triangles, gap xi + dHat/2, dense self contact); `--workload` picks another.

  value  : ms per step with every input resident in HBM (device-side CUDA events, max over ranks)
  e2e    : ms per step through the reference-named host API (codim_ipc_b200.Compute_* call pattern of the
           shim: host numpy/pinned buffers in, constraint set / gradient / Hessian triplets / dist2 out),
           host<->device copies inside the timed region
  N > 1  : the same scene, candidate pairs partitioned across ranks by cell ranges of the sorted hash
           ("strong" scaling); NCCL all-reduces the step size (min), energy and gradient (sum), min-dist (min)

`--impl reference` times the reference's own CPU implementation of the path -- FEM/IPC.h + Grid/SPATIAL_HASH.h
compiled from its sources into oracle/_ref/libcipc_refdrv_fast.so (the reference's own compiler flags; stand-ins only
for Eigen/Cabana/pybind11), or the oracle port when that library is absent -- with all host threads ON THE SAME WORKLOAD
(every warm-up and timed step is one full contact stage of `--workload`; only cfg5_4m, whose 4.28G triplets overflow the
reference's `int curStartInd` (IPC.h:1368), is sampled and marked `extrapolated`).

The GPU arm's `cpu_baseline` leg (rank 0, N=1) runs the same reference drivers on the same scene (>= 2 samples, min and
mean reported) and a PARITY leg: the parity build of the reference (-ffp-contract=off) and the CUDA path evaluate the same
scene and the SURVEY 8(d) gates are asserted and printed as `"parity": {...}`.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "IPC contact-stage ms/Newton iter (hash+barrier Hessian+ACCD)"
WORKLOADS = {  # name -> (n, layers)
    "cfg5_4m": (354, 16),
    "cfg5_1m": (224, 10),
    "cfg5_250k": (112, 10),
    "cfg5_62k": (56, 10),
}
# the CPU arms run the requested workload itself; only cfg5_4m is sampled (see the module docstring) and flagged as such
CPU_SAMPLE = {"cfg5_4m": ("cfg5_1m", None)}
L2_NOTE = "GPU arm: L2 flushed between timed steps (256 MiB write), working set >> L2"
PAR_NOTE = ("GPU arm: candidate pairs partitioned by voxel slabs across n_gpus ranks, per stage two NCCL all-reduces (sum: gradient + energy; "
            "min: step size + min distance); reference arm: all host cores of rank 0")


def scene_config(name, sc):
    """`config` of the JSON line: names the workload; identical in the GPU arm and the reference arm"""
    return {"workload": name, "triangles": int(len(sc["BT"])), "nodes": int(len(sc["X"])), "boundary_edges": int(len(sc["BE"])),
            "dHat": float(np.sqrt(sc["dHat2"])), "xi": float(sc["xi"]), "l2": L2_NOTE, "parallelism": PAR_NOTE}


def make_scene(name):
    from codim_ipc_b200 import scenes
    n, L = WORKLOADS[name]
    return scenes.cloth_stack(n, L)


# ----------------------------------------------------------------------------- clocks
class ClockSampler(threading.Thread):
    """samples nvidia-smi every 200 ms while the timed region runs (B200_PROFILING.md clocks line)"""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.rows = index, False, []

    def run(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 3 + i and r[3 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(self.rows[0][1]) if self.rows[0][1].replace(".", "").isdigit() else None,
                "reasons": reasons, "samples": len(self.rows)}


# ----------------------------------------------------------------------------- CPU arm
def cpu_backend(parity=False):
    """(kind, Scene class).  kind = "reference": oracle/_ref -- the reference's own FEM/IPC.h + SPATIAL_HASH.h compiled from its
    sources (oracle/Makefile `ref`): the TIMING build (the reference's own flags) unless parity=True (the -ffp-contract=off
    build whose constraint-set membership is bit-comparable).  kind = "port": the oracle restatement (when oracle/_ref is absent)."""
    from oracle import cipc_oracle as O
    if not parity and O.refdrv_fast() is not None:
        return "reference", O.RefSceneFast
    if O.refdrv() is not None:
        return "reference", O.RefScene
    return "port", O.OracleScene


def cpu_set_threads(Scene, threads):
    from oracle import cipc_oracle as O
    O.set_num_threads(threads)
    for L in (O.refdrv(), O.refdrv_fast()):
        if L is not None:
            L.ref_set_num_threads(int(threads))


def cpu_contact_stage(S, sc):
    """one contact stage on the CPU scene object S; returns (seconds, per-stage dict, nConstraints)"""
    st = {}
    t0 = time.perf_counter()
    cs, info = S.constraint_set(sc["dHat2"], sc["xi"]); t1 = time.perf_counter(); st["constraint_set"] = t1 - t0
    S.barrier(cs, info, sc["dHat2"], sc["kappa"], sc["xi"]); t2 = time.perf_counter(); st["barrier_E"] = t2 - t1
    S.barrier_gradient(cs, info, sc["dHat2"], sc["kappa"], sc["xi"]); t3 = time.perf_counter(); st["barrier_g"] = t3 - t2
    S.barrier_hessian_notfetch(cs, info, sc["dHat2"], sc["kappa"], sc["xi"], True); t4 = time.perf_counter(); st["barrier_H"] = t4 - t3
    S.step_size(sc["p"], sc["xi"], 1.0); t5 = time.perf_counter(); st["step_size"] = t5 - t4
    S.min_dist2(cs, sc["xi"]); S.min_dist2(cs, sc["xi"]); t6 = time.perf_counter(); st["min_dist_x2"] = t6 - t5
    return t6 - t0, st, len(cs)


def cpu_describe(kind):
    return ("the reference's own FEM/IPC.h + Grid/SPATIAL_HASH.h drivers (oracle/_ref timing build: -O3 -mavx2 -mfma -mbmi2 -fopenmp as the "
            "reference's CMakeLists.txt:20, Par_Each on OpenMP)" if kind == "reference"
            else "oracle port (-O3 -mavx2 -mfma -fopenmp, the reference's parallel structure)")


def cpu_samples(workload, n_warm, n_timed, budget_s=None):
    """run n_warm + n_timed full contact stages of `workload` (cfg5_4m: of its sample) on all host cores.  budget_s: wall-clock
    bound of the whole run -- when the first stage shows that n_warm + n_timed stages would exceed it, fewer stages are run
    (never fewer than 1 warm-up + 3 timed) and the executed counts are returned"""
    sample, _ = CPU_SAMPLE.get(workload, (workload, 1.0))
    sc = make_scene(sample)
    cores = os.cpu_count()  # torchrun exports OMP_NUM_THREADS=1; the CPU arm uses every host core
    kind, Scene = cpu_backend()
    cpu_set_threads(Scene, cores)
    S = Scene(sc)
    times, st, nC = [], {}, 0
    i = 0
    while i < n_warm + n_timed:
        t, st, nC = cpu_contact_stage(S, sc)
        if i == 0 and budget_s is not None and t * (n_warm + n_timed) > budget_s:
            fit = max(4, int(budget_s / t))
            n_warm = max(1, min(n_warm, fit - 3)) if n_warm else 0
            n_timed = min(n_timed, max(3, fit - n_warm))  # cut down to what fits, never below 3 (nor above the request)
        if i >= n_warm:
            times.append(t)
        i += 1
    scale, extrap = 1.0, None
    if sample != workload:
        # cfg5_4m only: scaled by the constraint count of the full workload (known from the scene generator's density) -- an
        # extrapolation, flagged at the top level of the JSON line
        full = WORKLOADS[workload][0] ** 2 * 2 * WORKLOADS[workload][1]
        scale = full / float(len(sc["BT"]))
        extrap = {"extrapolated": True, "sample_workload": sample, "scale": scale,
                  "why": "4.28G triplets overflow the reference's int triplet offsets (FEM/IPC.h:1368) and 68 GB of host triplets"}
    return dict(kind=kind, cores=cores, times=[t * scale for t in times], stages=st, nC=nC, sample=sample, scene=sc, scale=scale, extrap=extrap,
                n_warm=n_warm, n_timed=len(times))


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # every warm-up and timed step is one FULL contact stage of the workload (~13 s at cfg5_1m on 16 cores).  The run is bounded by
    # CIPC_REF_BUDGET_S seconds of CPU work (default 420): if W + K stages do not fit, fewer are executed and "steps" / "warmup"
    # report what was executed, with the request beside them.
    budget = float(os.environ.get("CIPC_REF_BUDGET_S", "420"))
    r = cpu_samples(args.workload, args.warmup, args.steps, budget)
    ms = 1e3 * float(np.mean(r["times"]))
    sc_full = r["scene"] if r["sample"] == args.workload else make_scene(args.workload)
    desc = "%s: %d warm-up + %d timed full contact stages of %s (%d triangles, %d constraints)%s" % (
        cpu_describe(r["kind"]), r["n_warm"], r["n_timed"], r["sample"], len(r["scene"]["BT"]), r["nC"],
        "" if r["extrap"] is None else ", scaled x%.2f to %s" % (r["scale"], args.workload))
    line = {"impl": "reference", "metric": METRIC, "value": ms, "unit": "ms", "n_gpus": args.gpus, "steps": r["n_timed"], "warmup": r["n_warm"],
            "ms_per_step": ms, "higher_is_better": False, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": scene_config(args.workload, sc_full),
            "cpu_baseline": {"value": ms, "unit": "ms", "cores": r["cores"], "kind": r["kind"], "sample": desc,
                             "min_ms": 1e3 * min(r["times"]), "max_ms": 1e3 * max(r["times"]),
                             "stages_s": {k: round(v, 4) for k, v in r["stages"].items()}},
            "e2e": {"value": ms, "unit": "ms", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0,
            "counts": {"constraints": int(r["nC"])}}
    if (r["n_timed"], r["n_warm"]) != (args.steps, args.warmup):
        line["requested"] = {"steps": args.steps, "warmup": args.warmup, "budget_s": budget}
    if r["extrap"] is not None:
        line.update(r["extrap"])
    print(json.dumps(line))


# ----------------------------------------------------------------------------- parity leg (GPU arm, rank 0, N=1)
TOL_EGH = 1e-9   # north_star: energy, gradient and Hessian entries within 1e-9 relative in fp64
TOL_STEP = 1e-12  # CCD step size within 1e-12 relative and never larger than the reference's


def parity_leg(ctx, sc, local, workload):
    """The CUDA path and the reference (parity build of oracle/_ref, else the oracle port) evaluate the SAME scene; the SURVEY
    8(d) gates are evaluated and returned.  The Hessian compared is the one the timed kernel produces: the device-resident
    fused gradient + Hessian pass (k_hessian_fused) read back with cipc_get_triplets, block by block against the reference's
    triplet vector; the factor + host-expansion delivery of Compute_Barrier_Hessian is compared as well."""
    import psutil
    from codim_ipc_b200 import multi
    from oracle import cipc_oracle as O
    kind, Scene = cpu_backend(parity=True)
    cpu_set_threads(Scene, os.cpu_count())
    R = Scene(sc)
    dHat2, xi, kappa = sc["dHat2"], sc["xi"], sc["kappa"]
    nV = len(sc["X"])
    out = {"workload": workload, "against": ("oracle/_ref parity build (the reference's FEM/IPC.h, -ffp-contract=off)" if kind == "reference"
                                             else "oracle port"), "tolerances": {"E_g_H": TOL_EGH, "step": TOL_STEP}}
    key = lambda a: np.lexsort(a.T[::-1])
    ctx.set_scene(sc)
    cs_g, info_g = ctx.constraint_set(dHat2, xi)
    cs_r, info_r = R.constraint_set(dHat2, xi)
    og, orr = key(cs_g), key(cs_r)
    cs_g, info_g, cs, info = cs_g[og], info_g[og], cs_r[orr], info_r[orr]
    same = cs_g.shape == cs.shape and np.array_equal(cs_g, cs) and np.array_equal(info_g, info)
    out["constraint_set"] = {"n_gpu": int(len(cs_g)), "n_ref": int(len(cs)), "identical_as_sorted_sets": bool(same)}
    def step_gate():
        a_g = ctx.step_size(xi, 1.0); a_r = R.step_size(sc["p"], xi, 1.0)
        return {"gpu": a_g, "ref": a_r, "rel_diff": (a_r - a_g) / a_r, "ok": bool(a_g <= a_r and (a_r - a_g) <= TOL_STEP * a_r)}

    if not same or len(cs) == 0:
        out["step_size"] = step_gate()
        out["pass"] = bool(same and out["step_size"]["ok"])
        return out
    ctx.set_constraints(cs, info)
    E_r = R.barrier(cs, info, dHat2, kappa, xi, E0=0.5); E_g = ctx.barrier_energy(dHat2, kappa, xi, E=0.5)
    out["energy"] = {"rel_err": abs(E_g - E_r) / abs(E_r), "ok": bool(abs(E_g - E_r) <= TOL_EGH * abs(E_r))}
    g_r = R.barrier_gradient(cs, info, dHat2, kappa, xi); g_g = ctx.barrier_gradient(dHat2, kappa, xi)
    gs = float(np.abs(g_r).max())
    # the timed kernels: gradient + Hessian in one pass, results resident on the device
    nT = ctx.barrier_gradient_hessian_dev(dHat2, kappa, xi)
    ctx.sync()  # the library may run on its own stream: torch's copy below is not ordered against it
    g_f = multi.wrap_device_f64(ctx.dev_ptrs()["g"], 3 * nV, local).cpu().numpy().reshape(nV, 3)
    e1, e2 = float(np.abs(g_g - g_r).max()) / gs, float(np.abs(g_f - g_r).max()) / gs
    out["gradient"] = {"inf_norm_ref": gs, "rel_err_host_api": e1, "rel_err_fused_kernel": e2, "ok": bool(max(e1, e2) <= TOL_EGH)}
    need = 2 * nT * 16 + (4 << 30)
    sub = None
    if psutil.virtual_memory().available < need:  # bounded host memory: every k-th stencil instead of all of them
        k = int(np.ceil(need / max(psutil.virtual_memory().available, 1))) * 2
        sub = np.arange(0, len(cs), k)
        ctx.set_constraints(cs[sub], info[sub])
        nT = ctx.barrier_gradient_hessian_dev(dHat2, kappa, xi)
    csH, infoH = (cs, info) if sub is None else (cs[sub], info[sub])
    trip = ctx.get_triplets(nT)
    n_r = R.barrier_hessian_notfetch(csH, infoH, dHat2, kappa, xi, True)
    ptr_r, _ = R.triplets_data()
    c1 = O.compare_triplet_blocks(trip.ctypes.data, ptr_r, csH) if n_r == nT else None
    trip = ctx.barrier_hessian(dHat2, kappa, xi, True, out=trip)  # factor kernels + host expansion (the shim's delivery)
    c2 = O.compare_triplet_blocks(trip.ctypes.data, ptr_r, csH) if n_r == len(trip) else None
    del trip
    if hasattr(R, "release_triplets"):
        R.release_triplets()
    okH = all(c is not None and c["index_mismatches"] == 0 and c["max_block_rel_err"] <= TOL_EGH for c in (c1, c2))
    out["hessian"] = {"triplets_gpu": int(nT), "triplets_ref": int(n_r), "blocks": int(len(csH)), "all_blocks": sub is None,
                      "fused_kernel_vs_ref": c1, "host_delivery_vs_ref": c2, "ok": bool(okH)}
    if sub is not None:
        ctx.set_constraints(cs, info)
    d_g, m_g = ctx.min_dist2(xi); d_r, m_r = R.min_dist2(cs, xi)
    out["min_dist2"] = {"dist2_bit_identical": bool(np.array_equal(d_g, d_r)), "min_equal": bool(m_g == m_r),
                        "ok": bool(np.array_equal(d_g, d_r) and m_g == m_r)}
    out["step_size"] = step_gate()
    out["pass"] = bool(same and all(out[k]["ok"] for k in ("energy", "gradient", "hessian", "min_dist2", "step_size")))
    return out


# ----------------------------------------------------------------------------- GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg5_1m", choices=list(WORKLOADS))
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-friction", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--stage-report", action="store_true", help="print the per-stage device times to stderr")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)

    import torch
    import codim_ipc_b200 as cipc
    from codim_ipc_b200 import multi

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        host_pg = dist.new_group(backend="gloo")  # host-side waits that must not occupy the GPUs (a NCCL barrier spins in a kernel)
    dc = multi.DistContact() if world > 1 else None

    sc = make_scene(args.workload)
    nV, nT = len(sc["X"]), len(sc["BT"])
    ctx = cipc.ContactContext(local, rank, world)
    # one explicit stream for the library's kernels, torch's events / copies and NCCL (a handle of 0 -- torch's default
    # stream -- would make cipc_set_stream fall back to the library's own non-blocking stream, which nothing else orders against)
    stream = torch.cuda.Stream(device=local)
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    ctx.set_stream(stream.cuda_stream)
    ctx.set_scene(sc)
    dHat2, xi, kappa = sc["dHat2"], sc["xi"], sc["kappa"]
    scal = multi.wrap_device_f64(ctx.dev_ptrs()["scalars"], 16, local)

    dc_ranks = dc
    g_views = {}

    def g_view():  # torch view of the library's gradient buffer (3 nV + 1 doubles), cached per device pointer
        ptr = ctx.L.cipc_dev_gradient(ctx.h)
        if ptr not in g_views:
            g_views[ptr] = multi.wrap_device_f64(ptr, 3 * nV + 1, local)
        return g_views[ptr]

    def device_step(collectives=True):
        """inputs resident; results stay on the device.  N > 1: two collectives per stage -- one all-reduce(sum) over the 3 nV
        gradient with the energy in slot 3 nV (cipc_dev_gradient) and one all-reduce(min) over (step, min distance)"""
        nC = ctx.constraint_set(dHat2, xi, fetch=False)
        ctx.barrier_energy_dev(dHat2, kappa, xi)
        # gradient and PSD-projected Hessian in one pass over the stencils (the Newton iteration evaluates them back to back);
        # the triplet stream is materialised in HBM (what a device-side solver / CSR assembly consumes)
        nTrip = ctx.barrier_gradient_hessian_dev(dHat2, kappa, xi)
        work = None
        dc = dc_ranks if collectives else None
        if dc is not None:  # asynchronous: the sum travels over NVLink while the step-size search (independent of it) runs
            work = dc.dist.all_reduce(g_view(), op=dc.dist.ReduceOp.SUM, async_op=True)
        ctx.step_size_dev(xi, 1.0)
        for _ in range(2):
            ctx.min_dist2_dev(xi)
        if dc is not None:
            dc.dist.all_reduce(scal[1:3], op=dc.dist.ReduceOp.MIN)  # step size, min dist2 (bit pattern of a positive double: same order)
            work.wait()  # the stage ends with every collective complete (stream-ordered: the timing event is recorded after it)
        return nC, nTrip

    def sync_all():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    # ---- device-resident timing
    for _ in range(args.warmup):
        nC, nTrip = device_step()
    sync_all()
    sampler = ClockSampler(local)
    sampler.start()
    stage_names = ["ccs_hash_build", "ccs_pairs", "ccs_narrow", "ccs_merge", "barrier_E", "barrier_g", "barrier_H", "k_barrier_hessian",
                   "ccd_hash_build", "ccd_pairs", "ccd_accd", "min_dist"]
    launches0 = cipc.kernel_launches()
    ctx.set_timing(False)  # the per-stage event scopes are instrumentation: off inside the timed loop, on again for the stage report
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    kH_ms = []
    for i in range(args.steps):
        flush.zero_()  # L2 flush between timed iterations (outside the per-step event pair)
        ev[i][0].record(stream)
        nC, nTrip = device_step()
        ev[i][1].record(stream)
    sync_all()
    launches = cipc.kernel_launches() - launches0
    ctx.set_timing(True)
    dev_ms = sum(a.elapsed_time(b) for a, b in ev) / args.steps
    # the dominant kernel, timed live with CUDA events on its own stream (stage timers of the last step)
    ctx.barrier_gradient_hessian_dev(dHat2, kappa, xi)
    kH = ctx.stage_ms("k_hessian_fused0")
    t = torch.tensor([dev_ms], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms = float(t.item())

    # N > 1: where the step goes -- every rank's stage without the collectives, and the collectives alone (untimed extra steps,
    # L2 not flushed); the gap between max(local) + collectives and `value` is rank skew inside the collectives
    scaling_breakdown = None
    if dist is not None:
        def avg_ms(fn, k):
            sync_all()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            for _ in range(k):
                fn()
            b.record(stream)
            torch.cuda.synchronize()
            return a.elapsed_time(b) / k

        def collectives_only():
            w = dist.all_reduce(g_view(), op=dist.ReduceOp.SUM, async_op=True)
            dist.all_reduce(scal[1:3], op=dist.ReduceOp.MIN)
            w.wait()

        ctx.set_timing(False)
        tl = torch.tensor([avg_ms(lambda: device_step(False), args.steps), avg_ms(collectives_only, args.steps)], dtype=torch.float64, device="cuda")
        ctx.set_timing(True)
        tall = [torch.zeros_like(tl) for _ in range(world)]
        dist.all_gather(tall, tl)
        scaling_breakdown = {"local_ms_per_rank": [round(float(x[0]), 4) for x in tall],
                             "collectives_only_ms": round(max(float(x[1]) for x in tall), 4)}

    # per-stage report (one extra untimed step)
    stages = {}
    ctx.constraint_set(dHat2, xi, fetch=False)
    for s in stage_names[:4]:
        stages[s] = ctx.stage_ms(s)
    counters = {k: ctx.counter(k) for k in ("hash_entries", "hash_cells", "candidates_pt", "candidates_ee", "candidates_pe", "candidates_pp", "constraints")}
    ctx.barrier_energy_dev(dHat2, kappa, xi); stages["barrier_E"] = ctx.stage_ms("barrier_E")
    ctx.barrier_gradient_dev(dHat2, kappa, xi); stages["barrier_g"] = ctx.stage_ms("barrier_g")
    ctx.barrier_gradient_hessian_dev(dHat2, kappa, xi); stages["barrier_gH_fused"] = ctx.stage_ms("barrier_H")
    ctx.barrier_hessian_dev(dHat2, kappa, xi, True); stages["barrier_H_fused"] = ctx.stage_ms("barrier_H")
    for _ in range(2):  # the first call may allocate the factor buffers inside the timed scope
        ctx.barrier_hessian(dHat2, kappa, xi, True, fetch=False); stages["barrier_H_factor_only"] = ctx.stage_ms("barrier_H")
    ctx.dev_triplets(); stages["barrier_H_expand_only"] = ctx.stage_ms("k_barrier_hessian")
    ctx.step_size_dev(xi, 1.0)
    for s in ("ccd_hash_build", "ccd_pairs", "ccd_accd"):
        stages[s] = ctx.stage_ms(s)
    counters["ccd_pairs"] = ctx.counter("ccd_pairs")
    ctx.min_dist2_dev(xi); stages["min_dist"] = ctx.stage_ms("min_dist")

    # lagged friction (FEM/FRICTION.h) on the same constraint set: reported beside the metric, not inside `value`
    friction = {}
    if args.workload == "cfg5_4m":
        args.no_friction = True  # the side measurements (friction, CSR assembly) need a second 68 GB triplet stream
    if not args.no_friction:
        rng = np.random.default_rng(17)
        Xn = sc["X"] - rng.normal(size=sc["X"].shape) * np.where(rng.random(nV) < 0.5, 2e-6, 5e-5)[:, None]  # both sides of eps_v h = 1e-5
        ctx.set_prev_positions(Xn)
        for _ in range(2):
            nF = ctx.friction_basis(dHat2, kappa, xi, fetch=False); friction["basis"] = ctx.stage_ms("friction_basis")
            ctx.friction_energy_dev(1e-10, 0.4); friction["E"] = ctx.stage_ms("friction_E")
            ctx.friction_gradient_dev(1e-10, 0.4, accumulate=False); friction["g"] = ctx.stage_ms("friction_g")
            nFT = ctx.friction_hessian(1e-10, 0.4, True, fetch=False); friction["H_factor_only"] = ctx.stage_ms("friction_H")
            ctx.dev_triplets(); friction["H_expand_only"] = ctx.stage_ms("k_barrier_hessian")
            ctx.friction_hessian_dev(1e-10, 0.4, True); friction["H_fused"] = ctx.stage_ms("friction_H")
        friction = {k: round(v, 4) for k, v in friction.items()}
        friction["stencils"] = int(nF); friction["triplets"] = int(nFT)

    # triplets -> CSR assembly on the device, barrier Hessian only: reported beside the metric
    csr = {}
    if not args.no_friction:
        for _ in range(2):
            ctx.csr_begin()
            ctx.barrier_hessian_dev(dHat2, kappa, xi, True)
            ctx.csr_add(); csr["blocks"] = ctx.stage_ms("csr_add")
            nnz = ctx.csr_finish(fetch=False)
            for k in ("csr_sort", "csr_pattern", "csr_emit"):
                csr[k[4:]] = ctx.stage_ms(k)
        csr = {k: round(v, 4) for k, v in csr.items()}
        csr["nnz"] = int(nnz); csr["blocks_in"] = ctx.counter("csr_blocks_in"); csr["blocks_unique"] = ctx.counter("csr_blocks_unique")

    # ---- end-to-end through the host API (pinned host buffers; copies inside the timed region)
    e2e = None
    if not args.no_e2e:
        pin = lambda shape, dt: torch.empty(shape, dtype=dt).pin_memory().numpy()
        X4 = pin((nV, 4), torch.float64); X4[:, :3] = sc["X"]; X4[:, 3] = 0
        X04 = pin((nV, 4), torch.float64); X04[:, :3] = sc["X0"]; X04[:, 3] = 0
        p_h = pin((nV * 3,), torch.float64); p_h[:] = sc["p"].ravel()
        g_h = pin((nV, 4), torch.float64)
        big = args.workload == "cfg5_4m"  # 4.28G raw triplets = 68 GB: only the merged delivery is timed end to end, side measurements off
        ctx.set_positions(sc["X"])
        nMerged = ctx.barrier_hessian_merged(dHat2, kappa, xi, True, fetch=False)
        nC_cap, nT_cap = int(nC * 1.05) + 1024, int((nMerged if big else nTrip) * 1.05) + 1024
        cs_h = pin((nC_cap, 4), torch.int32); info_h = pin((nC_cap, 2), torch.float64); d_h = pin((nC_cap,), torch.float64)
        trip_raw = torch.empty((nT_cap, 2), dtype=torch.float64).pin_memory().numpy()
        trip_h = trip_raw.view(cipc.TRIPLET_DTYPE).reshape(-1)
        import ctypes as C
        # the reference's containers: VECTOR<int,2|3> are 16-byte records (stride 4 ints)
        BE4 = np.zeros((len(sc["BE"]), 4), np.int32); BE4[:, :2] = sc["BE"]
        BT4 = np.zeros((len(sc["BT"]), 4), np.int32); BT4[:, :3] = sc["BT"]
        ctx.set_topology(nV, sc["BN"], BE4, BT4, 0, sc["codim"], sc["DBC"])
        ctx.set_positions(X4); ctx.set_rest_positions(X04); ctx.set_search_dir(p_h)
        pcie = [0]

        calls = {}

        def tick(name, t0):
            calls[name] = calls.get(name, 0.0) + (time.perf_counter() - t0)
            return time.perf_counter()

        def e2e_step(merged):
            h2d = d2h = 0
            t0 = time.perf_counter()
            # Compute_Constraint_Set: topology (content-hashed, re-uploaded only on change), X, x0 in; constraintSet, stencilInfo out
            ctx.set_topology(nV, sc["BN"], BE4, BT4, 0, sc["codim"], sc["DBC"])
            ctx.set_positions(X4); ctx.set_rest_positions(X04); h2d += 2 * X4.nbytes
            n = ctx.constraint_set(dHat2, xi, fetch=False)
            ctx._ck(ctx.L.cipc_get_constraints(ctx.h, cs_h.ctypes.data_as(C.POINTER(C.c_int32)), info_h.ctypes.data_as(C.POINTER(C.c_double))))
            d2h += n * 32
            t0 = tick("Compute_Constraint_Set", t0)
            # Compute_Barrier / _Gradient / _Hessian: X in (the shim keeps the constraint set it just produced resident)
            ctx.set_positions(X4); h2d += X4.nbytes
            E = ctx.barrier_energy(dHat2, kappa, xi, 0.0); d2h += 8
            if dc is not None:
                E = dc.sum_scalar(E)  # N > 1: every rank ends up with the complete energy / gradient / step / min distance
            t0 = tick("Compute_Barrier", t0)
            ctx.set_positions(X4); h2d += X4.nbytes
            g_h[:] = 0
            ctx.barrier_gradient(dHat2, kappa, xi, g_h); d2h += nV * 24
            if dc is not None:
                g_h[:, :3] = dc.sum_gradient(np.ascontiguousarray(g_h[:, :3]))
            t0 = tick("Compute_Barrier_Gradient", t0)
            ctx.set_positions(X4); h2d += X4.nbytes
            # the Hessian stays additive across ranks: every rank delivers the triplets of ITS constraints (the matrix is their sum)
            tr = (ctx.barrier_hessian_merged if merged else ctx.barrier_hessian)(dHat2, kappa, xi, True, out=trip_h); d2h += len(tr) * 16
            if merged:
                pcie[0] = ctx.counter("merge_blocks_unique") * 80
            else:
                pcie[0] = sum(ctx.counter(k) * b for k, b in (("hessian_4pt", 320), ("hessian_pe", 176), ("hessian_pp", 80))) + ctx.counter("hessian_mollified") * 2312
            ntr[0] = len(tr)
            t0 = tick("Compute_Barrier_Hessian", t0)
            # Compute_Intersection_Free_StepSize: topology check, X, searchDir in; step out
            ctx.set_topology(nV, sc["BN"], BE4, BT4, 0, sc["codim"], sc["DBC"])
            ctx.set_positions(X4); ctx.set_search_dir(p_h); h2d += X4.nbytes + p_h.nbytes
            a = ctx.step_size(xi, 1.0); d2h += 8
            if dc is not None:
                a = dc.min_step(a)
            t0 = tick("Compute_Intersection_Free_StepSize", t0)
            # Compute_Min_Dist2 x2: X in; dist2, min out
            for _ in range(2):
                ctx.set_positions(X4); h2d += X4.nbytes
                m = C.c_double(0)
                ctx._ck(ctx.L.cipc_min_dist2(ctx.h, C.c_double(xi), d_h.ctypes.data_as(C.POINTER(C.c_double)), C.byref(m))); d2h += n * 8 + 8
                mval = dc.min_step(m.value) if dc is not None else m.value
            t0 = tick("Compute_Min_Dist2_x2", t0)
            return h2d, d2h, (E, a, mval)

        ntr = [0]

        def e2e_run(merged):
            for _ in range(max(1, args.warmup - 1)):
                e2e_step(merged)
            sync_all()
            calls.clear()
            t0 = time.perf_counter()
            for _ in range(args.steps):
                h2d, d2h, res = e2e_step(merged)
            sync_all()
            ms = 1e3 * (time.perf_counter() - t0) / args.steps
            t = torch.tensor([ms], dtype=torch.float64, device="cuda")
            if dist is not None:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item()), int(h2d), int(d2h), {k: round(1e3 * v / args.steps, 3) for k, v in calls.items()}, ntr[0], pcie[0]

        raw_ms, raw_h2d, raw_d2h, raw_calls, raw_n, raw_pcie = e2e_run(False) if not big else (None, 0, 0, {}, int(nTrip), 0)
        mg_ms, mg_h2d, mg_d2h, mg_calls, mg_n, mg_pcie = e2e_run(True)
        e2e = {"value": mg_ms, "unit": "ms", "h2d_bytes_per_step": mg_h2d, "d2h_bytes_per_step": mg_d2h,
               "via": "C ABI through the ctypes mirror of the shim's call pattern, pinned host buffers, merged-triplet Hessian delivery",
               "note": "d2h = bytes delivered into host buffers.  Hessian: one triplet per distinct (row, col) (%d triplets = %d B; the duplicates "
                       "setFromTriplets would sum are summed on the device), crossing PCIe as %d B of unique upper 3x3 blocks and mirrored by the "
                       "host cores.  N > 1: energy, gradient, step and min distance are all-reduced inside the timed region; every rank delivers the "
                       "triplets of its own constraints (the matrix is their sum)." % (mg_n, mg_n * 16, mg_pcie),
               "calls_ms": mg_calls,
               "raw_triplets": {"value": raw_ms, "unit": "ms", "d2h_bytes_per_step": raw_d2h, "calls_ms": raw_calls,
                                "note": "the reference's exact 144/81/36 triplets per stencil (%d triplets = %d B), crossing PCIe as %d B of factors "
                                        "and expanded by the host cores (CIPC_TRIPLETS=raw in the shim)" % (raw_n, raw_n * 16, raw_pcie)}}
        # the same Hessian delivered as CSR assembled on the device instead of 16-byte triplets: what an
        # integration that feeds the solver's A->p/i/x directly would pay (reported beside e2e, not part of it)
        if csr:
            rp_h = pin((3 * nV + 1,), torch.int32); ci_h = pin((csr["nnz"] + 1024,), torch.int32); cv_h = pin((csr["nnz"] + 1024,), torch.float64)
            ts = []
            for _ in range(3):
                t0 = time.perf_counter()
                ctx.set_positions(X4)
                ctx.csr_begin(); ctx.barrier_hessian_dev(dHat2, kappa, xi, True); ctx.csr_add()
                nnz = ctx.csr_finish(fetch=False)
                ctx._ck(ctx.L.cipc_get_csr(ctx.h, rp_h.ctypes.data_as(C.POINTER(C.c_int32)), ci_h.ctypes.data_as(C.POINTER(C.c_int32)),
                                           cv_h.ctypes.data_as(C.POINTER(C.c_double))))
                ts.append(1e3 * (time.perf_counter() - t0))
            e2e["hessian_as_csr_ms"] = round(min(ts), 3)
            e2e["hessian_as_csr_bytes"] = int(nnz * 12 + (3 * nV + 1) * 4)
        # ---- the same stage through the COMPILED shim: the reference's driver translation unit on its real MESH_NODE / AoSoA /
        # std::vector containers (pageable, the triplet vector and dist2 fresh per call as in INC_POTENTIAL.h:321 /
        # IMPLICIT_EULER.h:122), built through codim-ipc_b200/shim (tests/shim_harness).  When the harness travelled here its
        # number is the headline e2e; the ctypes-mirror figure stays beside it.
        # N > 1: the shim is ONE calling thread over all N GPUs (CIPC_DEVICES -> cipc_create_multi: slab-partitioned pairs, NVLink peer
        # exchanges inside the library); rank 0 runs it while the other ranks have released their device buffers and wait.
        n4_keep = ctx.counter("hessian_4pt")
        if True:
            sys.path.insert(0, os.path.join(ROOT, "tests"))
            try:
                import shim_scene
                have_shim = shim_scene.present()
            except Exception:
                have_shim = False
            if world > 1:
                hs = torch.tensor([1 if have_shim else 0], device="cuda")
                dist.all_reduce(hs, op=dist.ReduceOp.MIN)
                have_shim = bool(hs.item())
                if have_shim:
                    ctx.close()  # every rank frees its contact context; rank 0 re-enters through the shim on all devices
                    del flush
                    torch.cuda.empty_cache()
                    sync_all()
                    if rank == 0:
                        os.environ["CIPC_DEVICES"] = ",".join(str(i) for i in range(world))
                    else:
                        have_shim = False
            if have_shim:
                if world == 1:
                    ctx.sync()
                    if big:  # 4M triangles: the bench context's 100+ GB of buffers make room for the shim's own context
                        ctx.close()
                        del flush
                        torch.cuda.empty_cache()
                S = shim_scene.ShimScene(sc)
                tot, per = [], {}
                for i in range(max(1, args.warmup - 1) + args.steps):
                    tm, rs = S.contact_stage(sc)
                    if i >= max(1, args.warmup - 1):
                        tot.append(sum(tm.values()))
                        for k, v in tm.items():
                            per[k] = per.get(k, 0.0) + v
                nCs, nTs = rs["nC"], rs["nTriplets"]
                shim_h2d = nV * (32 * 8 + 24 + 24)  # X per call (8 calls), x0 and searchDir once each
                shim_d2h = nCs * 32 + 8 + nV * 24 + nTs * 16 + 8 + 2 * (nCs * 8 + 8)
                e2e["ctypes_mirror"] = {"value": e2e["value"], "calls_ms": e2e["calls_ms"], "via": e2e["via"]}
                e2e.update({"value": 1e3 * float(np.mean(tot)), "h2d_bytes_per_step": int(shim_h2d), "d2h_bytes_per_step": int(shim_d2h),
                            "min_ms": 1e3 * min(tot), "max_ms": 1e3 * max(tot),
                            "via": "compiled shim harness (tests/shim_harness/_build/libcipc_shimdrv.so): the reference's six templates called in the "
                                   "reference's own call pattern on MESH_NODE / AoSoA / std::vector containers (pageable; triplet vector and dist2 "
                                   "freshly allocated per call), CIPC_TRIPLETS=merged" + ("" if world == 1 else
                                   ", CIPC_DEVICES=0..%d: one calling thread over all %d GPUs (cipc_create_multi)" % (world - 1, world)),
                            "calls_ms": {k: round(1e3 * v / args.steps, 3) for k, v in per.items()}})
                # The fresh std::vector<Eigen::Triplet> of every Hessian evaluation (INC_POTENTIAL.h:321) costs ~30 ms of page faults
                # at this size under glibc's default policy (allocations > 32 MB are mmap'ed and returned on free).  A process that
                # keeps freed memory (MALLOC_MMAP_MAX_=0 MALLOC_TRIM_THRESHOLD_=2147483647) re-uses the pages
                # of the previous Newton iteration; the same stage under that policy is reported beside the headline.
                try:
                    import ctypes as C2
                    libc = C2.CDLL("libc.so.6")
                    if libc.mallopt(-4, 0) == 1 and libc.mallopt(-1, 2147483647) == 1:  # M_MMAP_MAX = 0, M_TRIM_THRESHOLD = max
                        tot2 = []
                        for i in range(2 + args.steps):
                            tm, _ = S.contact_stage(sc)
                            if i >= 2:
                                tot2.append(sum(tm.values()))
                        e2e["with_malloc_reuse"] = {"value": 1e3 * float(np.mean(tot2)), "unit": "ms", "calls_ms": {k: round(1e3 * v, 3) for k, v in tm.items()},
                                                    "policy": "mallopt(M_MMAP_MAX, 0); mallopt(M_TRIM_THRESHOLD, INT_MAX)"}
                except Exception as ex:  # noqa: BLE001 -- a side measurement must not take the bench line down
                    e2e["with_malloc_reuse"] = {"error": str(ex)}
                del S
    sampler.stop_flag = True
    sampler.join(timeout=2)
    if dist is not None:
        dist.barrier(group=host_pg)  # the ranks that released their GPU for rank 0's multi-device shim run wait here on the host

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel: k_hessian_fused<0> (factor + triplet expansion of the PT/EE blocks)
    #   algorithmic bytes per 4-point stencil: write 144 triplets x 16 B; read stencil 16 + info 16 + offset 4 + index 4 + 4 positions x 32 B
    #   (the 12 gradient atomics of the stencil stay in L2 and are not counted)
    peaks = {}
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peaks = json.load(open(pk))
    peak = float(peaks.get("hbm_gbs", 6650.0))
    n4 = n4_keep if not args.no_e2e else ctx.counter("hessian_4pt")
    per_unit = 144 * 16 + 16 + 16 + 4 + 4 + 4 * 32
    alg_bytes = n4 * per_unit
    achieved = alg_bytes / (kH * 1e-3) / 1e9 if kH and kH > 0 else None
    roof = {"bound": "hbm", "kernel": "k_hessian_fused<0>", "achieved": achieved, "peak": peak, "unit": "GB/s",
            "frac": (achieved / peak) if achieved else None, "traffic": None, "kernel_ms": kH, "algorithmic_bytes": int(alg_bytes),
            "units_per_launch": int(n4), "bytes_per_unit": per_unit,
            "peak_source": "MEASURED_PEAKS.json (of measured)" if peaks else "fallback 6650 GB/s (of fallback)",
            "note": "write-dominated stream co-limited by the FP64 pipe: the same launch runs the 5x5 Jacobi PSD projection of every stencil "
                    "(~6K fp64 instructions each); the expansion-only kernel it replaces runs at 0.98 of the copy peak (stages_ms.barrier_H_expand_only)"}
    # FP64 side of the same kernel (it is co-limited): fp64 thread instructions per PT/EE stencil counted by ncu on this build
    # (smsp__sass_thread_inst_executed_op_{dfma,dmul,dadd}_pred_on, profiles/r2_fp64_ops_k_hessian_fused.csv) against the
    # non-tensor FP64 FMA rate measured on THIS box by tools/fp64_fma_probe (dependent DFMA chains, CUDA events)
    FP64_OPS_4PT = {"dfma": 1865, "dmul": 1410, "dadd": 258}
    fp64_peak, fp64_src = 34.2, "recorded on this pool's B200 (tools/fp64_fma_probe, 2026-10-17)"
    probe = os.path.join(ROOT, "tools", "fp64_fma_probe")
    if os.path.exists(probe) and world == 1:
        try:
            pj = json.loads(subprocess.run([probe, str(local)], capture_output=True, text=True, timeout=60).stdout.strip().splitlines()[-1])
            fp64_peak, fp64_src = float(pj["fp64_tflops"]), "tools/fp64_fma_probe run in this process' job: %s" % pj["how"]
        except Exception:
            pass
    if kH and kH > 0:
        flops = n4 * (2 * FP64_OPS_4PT["dfma"] + FP64_OPS_4PT["dmul"] + FP64_OPS_4PT["dadd"])
        slots = n4 * sum(FP64_OPS_4PT.values())
        roof_fp64 = {"bound": "fp64", "kernel": "k_hessian_fused<0>", "achieved": flops / (kH * 1e-3) / 1e12, "peak": fp64_peak, "unit": "TFLOP/s",
                     "frac": flops / (kH * 1e-3) / 1e12 / fp64_peak,
                     "pipe_frac": 2.0 * slots / (kH * 1e-3) / 1e12 / fp64_peak,  # every fp64 instruction occupies one FMA issue slot
                     "fp64_thread_instructions_per_unit": FP64_OPS_4PT, "peak_source": fp64_src}
        if args.workload == "cfg5_1m" and world == 1 and stages.get("ccd_accd"):
            # additive CCD (k_accd_pt + k_accd_ee on this workload's 1.17M swept pairs): fp64 thread instructions per call from the
            # same ncu pass -- dadd 154.6M, dfma 78.7M, dmul 155.1M -- against the live stage time
            accd_flops = 154.6e6 + 2 * 78.7e6 + 155.1e6
            roof_fp64["accd"] = {"kernels": "k_accd_pt + k_accd_ee", "achieved": accd_flops / (stages["ccd_accd"] * 1e-3) / 1e12, "unit": "TFLOP/s",
                                 "frac": accd_flops / (stages["ccd_accd"] * 1e-3) / 1e12 / fp64_peak,
                                 "note": "divergent per-pair iteration (1-40 ACCD steps), latency-bound: 0.1 ms of the stage"}
    else:
        roof_fp64 = None
    tr_path = os.path.join(ROOT, "profiles", "traffic_k_hessian_fused.json")
    if os.path.exists(tr_path):
        try:
            tj = json.load(open(tr_path))
            if tj.get("workload") == args.workload and world == 1:  # captured for this workload at N=1 only
                roof["traffic"] = tj.get("dram_bytes_per_launch")
        except Exception:
            pass

    cpu = None
    extrap = None
    if not args.no_cpu and world == 1:
        r = cpu_samples(args.workload, 0, 2)  # the first sample warms the allocator / page cache of the CPU path
        cpu = {"value": 1e3 * min(r["times"]), "unit": "ms", "cores": r["cores"], "kind": r["kind"],
               "sample": "%s: 2 full contact stages of %s (%d triangles, %d constraints)%s; value = min, mean_ms beside it" % (
                   cpu_describe(r["kind"]), r["sample"], len(r["scene"]["BT"]), r["nC"],
                   "" if r["extrap"] is None else ", scaled x%.2f to %s (EXTRAPOLATED)" % (r["scale"], args.workload)),
               "min_ms": 1e3 * min(r["times"]), "mean_ms": 1e3 * float(np.mean(r["times"])), "samples_ms": [round(1e3 * t, 1) for t in r["times"]],
               "stages_s": {k: round(v, 4) for k, v in r["stages"].items()}}
        if r["extrap"] is not None:
            cpu["extrapolated"] = r["extrap"]
        del r
    parity = None
    if not args.no_parity and world == 1 and args.workload not in CPU_SAMPLE:
        parity = parity_leg(ctx, sc, local, args.workload)

    line = {"metric": METRIC, "value": dev_ms, "unit": "ms", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms,
            "higher_is_better": False, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": scene_config(args.workload, sc),
            "counts": {"constraints_rank0": int(nC), "triplets_rank0": int(nTrip), "world": world},
            "roofline": roof, "roofline_fp64": roof_fp64, "cpu_baseline": cpu, "parity": parity, "e2e": e2e, "gpu_launches": int(launches), "clocks": sampler.summary(),
            "stages_ms": {k: (round(v, 4) if v is not None else None) for k, v in stages.items()}, "counters": counters,
            "friction_stages_ms": friction, "csr_stages_ms": csr}
    if scaling_breakdown is not None:
        line["scaling_breakdown"] = scaling_breakdown
    if args.stage_report:
        print(json.dumps(line["stages_ms"], indent=1), file=sys.stderr)
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
