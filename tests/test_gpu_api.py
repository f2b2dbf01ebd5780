"""The reference-named entry points (codim_ipc_b200.Compute_*) keep the reference's call semantics:
outputs resized and filled, E accumulated, g accumulated in place, triplets appended, step in/out."""
import numpy as np
import pytest

from helpers import sort_cs

pytestmark = pytest.mark.gpu


def test_reference_named_entry_points():
    import codim_ipc_b200 as cipc
    from codim_ipc_b200 import scenes
    from oracle import cipc_oracle as O
    sc = scenes.mixed_small()
    S = O.OracleScene(sc)
    rod = sc["BE"][len(sc["BE"]) - sc["nRod"]:]
    particle = sc["BN"][sc["codim"][1]:]
    X4 = np.zeros((len(sc["X"]), 4)); X4[:, :3] = sc["X"]  # VECTOR<double,3> storage: 32-byte elements
    cs, ptee, info = cipc.Compute_Constraint_Set(X4, sc["X0"], sc["BN"], sc["BE"], sc["BT"], particle, rod, sc["NNX"], sc["BNArea"],
                                                 sc["BEArea"], sc["BTArea"], sc["codim"], sc["DBC"], sc["dHat2"], sc["xi"], False)
    cs_o, info_o = S.constraint_set(sc["dHat2"], sc["xi"])
    assert np.array_equal(sort_cs(cs), sort_cs(cs_o)) and len(ptee) == 0
    E = cipc.Compute_Barrier(X4, sc["X0"], cs, info, sc["dHat2"], sc["kappa"], sc["xi"], 1.5)
    assert abs((E - 1.5) - S.barrier(cs, info, sc["dHat2"], sc["kappa"], sc["xi"])) <= 1e-9 * abs(E - 1.5)
    g = np.ones((len(sc["X"]), 4))
    cipc.Compute_Barrier_Gradient(X4, cs, info, sc["dHat2"], sc["kappa"], sc["xi"], sc["X0"], g)
    g_o = S.barrier_gradient(cs, info, sc["dHat2"], sc["kappa"], sc["xi"])
    assert np.abs(g[:, :3] - 1.0 - g_o).max() <= 1e-9 * np.abs(g_o).max() and np.all(g[:, 3] == 1.0)
    pre = np.zeros(7, cipc.TRIPLET_DTYPE); pre["val"] = 3.0
    trip = cipc.Compute_Barrier_Hessian(X4, sc["X0"], cs, info, sc["dHat2"], sc["kappa"], sc["xi"], True, pre)
    r_o, c_o, v_o = S.barrier_hessian(cs, info, sc["dHat2"], sc["kappa"], sc["xi"])
    assert len(trip) == 7 + len(v_o) and np.all(trip["val"][:7] == 3.0) and np.array_equal(trip["row"][7:], r_o)
    a = cipc.Compute_Intersection_Free_StepSize(X4, sc["BN"], sc["BE"], sc["BT"], particle, rod, sc["NNX"], sc["codim"], sc["DBC"],
                                                sc["p"].ravel(), sc["xi"], 1.0)
    a_o = S.step_size(sc["p"], sc["xi"], 1.0)
    assert a <= a_o and a_o - a <= 1e-12 * a_o
    d, m = cipc.Compute_Min_Dist2(X4, cs, sc["xi"])
    d_o, m_o = S.min_dist2(cs, sc["xi"])
    assert np.array_equal(d, d_o) and m == m_o


def test_topology_reupload_only_on_change(ctx):
    from codim_ipc_b200 import scenes
    sc = scenes.cloth_stack(10, 2)
    ctx.set_scene(sc)
    n1 = ctx.constraint_set(sc["dHat2"], sc["xi"], fetch=False)
    ctx.set_scene(sc)  # identical content: the hash matches, resident copy kept
    n2 = ctx.constraint_set(sc["dHat2"], sc["xi"], fetch=False)
    sc2 = scenes.cloth_stack(10, 3)
    ctx.set_scene(sc2)
    n3 = ctx.constraint_set(sc2["dHat2"], sc2["xi"], fetch=False)
    assert n1 == n2 and n3 > n1


def test_rank_partition_covers_all_pairs():
    """world=2 contexts on one device: the union of the two ranks' pass-through constraints and the
    merged PP/PE multiplicities equals the single-rank result."""
    import codim_ipc_b200 as cipc
    from codim_ipc_b200 import scenes, multi
    sc = scenes.cloth_stack(24, 4)
    full = cipc.ContactContext(0)
    full.set_scene(sc)
    cs_full, _ = full.constraint_set(sc["dHat2"], sc["xi"])
    parts = []
    for r in range(2):
        c = cipc.ContactContext(0, rank=r, world=2)
        c.set_scene(sc)
        parts.append(c.constraint_set(sc["dHat2"], sc["xi"])[0])
        a_r = c.step_size(sc["xi"], 1.0)
        parts.append(a_r)
        c.close()
    merged = multi.merge_constraint_sets([parts[0], parts[2]])
    assert np.array_equal(sort_cs(merged), sort_cs(cs_full))
    assert min(parts[1], parts[3]) == full.step_size(sc["xi"], 1.0)
    full.close()


def test_triplet_delivery_paths_agree(ctx, monkeypatch):
    """host expansion of the compact factors == PCIe copy of the device-expanded stream"""
    from codim_ipc_b200 import scenes
    sc = scenes.mixed_small()
    ctx.set_scene(sc)
    ctx.constraint_set(sc["dHat2"], sc["xi"], fetch=False)
    a = ctx.barrier_hessian(sc["dHat2"], sc["kappa"], sc["xi"], True).copy()
    monkeypatch.setenv("CIPC_TRIPLETS_DMA", "1")
    b = ctx.barrier_hessian(sc["dHat2"], sc["kappa"], sc["xi"], True).copy()
    monkeypatch.delenv("CIPC_TRIPLETS_DMA")
    assert len(a) == len(b) > 0
    assert np.array_equal(a["row"], b["row"]) and np.array_equal(a["col"], b["col"])
    assert np.abs(a["val"] - b["val"]).max() <= 1e-13 * np.abs(b["val"]).max()
    monkeypatch.setenv("CIPC_HESSIAN_DENSE", "1")  # dense 12x12 eigen path for every stencil: the cross-check of the low-rank path
    c = ctx.barrier_hessian(sc["dHat2"], sc["kappa"], sc["xi"], True).copy()
    monkeypatch.delenv("CIPC_HESSIAN_DENSE")
    assert np.array_equal(a["row"], c["row"])
    from helpers import max_block_rel_err
    cs, _ = ctx.get_constraints()
    assert max_block_rel_err(cs, a["val"], c["val"]) <= 1e-10


def test_device_resident_line_search(ctx):
    """X = Xprev + alpha p formed on the device gives the same constraint set / min distance as uploading
    the host-formed trial positions (Shell/IMPLICIT_EULER.h:102-131), bit for bit"""
    from codim_ipc_b200 import scenes
    sc = scenes.cloth_stack(24, 4)
    ctx.set_scene(sc)
    ctx.save_positions()
    for alpha in (0.5, 0.125):
        ctx.step_positions(alpha)
        Xt = sc["X"] + alpha * sc["p"]
        assert np.array_equal(ctx.get_positions(), Xt)
        cs_d = sort_cs(ctx.constraint_set(sc["dHat2"], sc["xi"])[0])
        d_d, m_d = ctx.min_dist2(sc["xi"])
        ctx.set_positions(Xt)
        cs_h = sort_cs(ctx.constraint_set(sc["dHat2"], sc["xi"])[0])
        d_h, m_h = ctx.min_dist2(sc["xi"])
        assert np.array_equal(cs_d, cs_h) and m_d == m_h and len(cs_d) > 0


def test_fused_device_hessian_equals_factor_path(ctx):
    """cipc_barrier_hessian_dev (fused factor + expansion, stream left in HBM) == factor kernel + host expansion"""
    from codim_ipc_b200 import scenes
    for sc in (scenes.mixed_small(), scenes.cloth_stack(20, 4)):
        ctx.set_scene(sc)
        ctx.constraint_set(sc["dHat2"], sc["xi"], fetch=False)
        a = ctx.barrier_hessian(sc["dHat2"], sc["kappa"], sc["xi"], True).copy()
        n = ctx.barrier_hessian_dev(sc["dHat2"], sc["kappa"], sc["xi"], True)
        assert ctx.dev_triplets()
        b = ctx.get_triplets(n)
        assert n == len(a) > 0 and np.array_equal(a["row"], b["row"]) and np.array_equal(a["col"], b["col"])
        assert np.abs(a["val"] - b["val"]).max() <= 1e-13 * np.abs(a["val"]).max()


def test_empty_constraint_set_through_every_stage(ctx):
    """a scene with nothing in contact (cloth 1.5 dHat above the sphere): every stage returns an empty / neutral result"""
    from codim_ipc_b200 import scenes
    sc = scenes.cloth_on_sphere(24, draped=False)
    ctx.set_scene(sc)
    cs, info = ctx.constraint_set(sc["dHat2"], sc["xi"])
    assert cs.shape == (0, 4) and info.shape == (0, 2)
    assert ctx.barrier_energy(sc["dHat2"], sc["kappa"], sc["xi"], E=0.75) == 0.75
    assert not ctx.barrier_gradient(sc["dHat2"], sc["kappa"], sc["xi"]).any()
    assert len(ctx.barrier_hessian(sc["dHat2"], sc["kappa"], sc["xi"], True)) == 0
    assert ctx.barrier_hessian_dev(sc["dHat2"], sc["kappa"], sc["xi"], True) == 0 and ctx.dev_triplets()
    fcs, cp, B, nf = ctx.friction_basis(sc["dHat2"], sc["kappa"], sc["xi"])
    assert len(fcs) == len(cp) == len(B) == len(nf) == 0
    ctx.set_prev_positions(sc["X"])
    assert ctx.friction_energy(1e-10, 0.4, E=0.5) == 0.5
    assert not ctx.friction_gradient(1e-10, 0.4).any()
    assert len(ctx.friction_hessian(1e-10, 0.4, True)) == 0
    a = ctx.step_size(sc["xi"], 1.0)
    assert 0.0 < a <= 1.0
    d, m = ctx.min_dist2(sc["xi"])  # the reference returns early on an empty set and leaves its outputs untouched (IPC.h:2253)
    assert len(d) == 0 and m == 0.0
    assert len(ctx.barrier_hessian_merged(sc["dHat2"], sc["kappa"], sc["xi"], True)) == 0


def test_fused_gradient_hessian_equals_separate_calls(ctx):
    """cipc_barrier_gradient_hessian_dev == cipc_barrier_gradient + cipc_barrier_hessian on every stencil kind"""
    from codim_ipc_b200 import scenes, multi
    for sc in (scenes.mixed_small(), scenes.noodles(4, 40), scenes.cloth_stack(20, 4)):
        ctx.set_scene(sc)
        ctx.constraint_set(sc["dHat2"], sc["xi"], fetch=False)
        g_ref = ctx.barrier_gradient(sc["dHat2"], sc["kappa"], sc["xi"])
        t_ref = ctx.barrier_hessian(sc["dHat2"], sc["kappa"], sc["xi"], True).copy()
        n = ctx.barrier_gradient_hessian_dev(sc["dHat2"], sc["kappa"], sc["xi"])
        ctx.sync()
        nV = len(sc["X"])
        g = multi.wrap_device_f64(ctx.dev_ptrs()["g"], 3 * nV, 0).cpu().numpy().reshape(nV, 3)
        t = ctx.get_triplets(n)
        assert np.abs(g - g_ref).max() <= 1e-12 * np.abs(g_ref).max()
        assert n == len(t_ref) and np.array_equal(t["row"], t_ref["row"]) and np.abs(t["val"] - t_ref["val"]).max() <= 1e-13 * np.abs(t_ref["val"]).max()


def test_rank_partition_sums_to_single_rank_terms():
    """world=3 contexts on one device: barrier and friction energies / gradients and the per-rank CSR matrices of the
    Hessians add up to the single-rank results (what the all-reduces of DESIGN.md section 6 assemble)."""
    import scipy.sparse as sp
    import codim_ipc_b200 as cipc
    from codim_ipc_b200 import scenes
    sc = scenes.mixed_small()
    nV = len(sc["X"])
    rng = np.random.default_rng(4)
    Xn = sc["X"] - rng.normal(size=sc["X"].shape) * 1e-5

    def terms(c):
        c.set_scene(sc)
        c.constraint_set(sc["dHat2"], sc["xi"], fetch=False)
        E = c.barrier_energy(sc["dHat2"], sc["kappa"], sc["xi"])
        g = c.barrier_gradient(sc["dHat2"], sc["kappa"], sc["xi"])
        c.csr_begin()
        c.barrier_gradient_hessian_dev(sc["dHat2"], sc["kappa"], sc["xi"]); c.csr_add()
        c.friction_basis(sc["dHat2"], sc["kappa"], sc["xi"], fetch=False)
        c.set_prev_positions(Xn)
        Ef = c.friction_energy(1e-10, 0.4)
        gf = c.friction_gradient(1e-10, 0.4)
        c.friction_hessian_dev(1e-10, 0.4, True); c.csr_add()
        rp, ci, v = c.csr_finish()
        A = sp.csr_matrix((v, ci, rp), shape=(3 * nV, 3 * nV))
        return E, g, Ef, gf, A

    full = cipc.ContactContext(0)
    E, g, Ef, gf, A = terms(full)
    full.close()
    Es = Efs = 0.0; gs = np.zeros_like(g); gfs = np.zeros_like(gf); As = None
    for r in range(3):
        c = cipc.ContactContext(0, rank=r, world=3)
        e, gg, ef, ggf, a = terms(c)
        c.close()
        Es += e; Efs += ef; gs += gg; gfs += ggf
        As = a if As is None else As + a
    assert abs(Es - E) <= 1e-12 * abs(E) and abs(Efs - Ef) <= 1e-12 * abs(Ef)
    assert np.abs(gs - g).max() <= 1e-12 * np.abs(g).max() and np.abs(gfs - gf).max() <= 1e-12 * np.abs(gf).max()
    D = (As - A).tocoo()
    assert (np.abs(D.data).max() if D.nnz else 0.0) <= 1e-11 * np.abs(A.data).max()
