"""Synthetic contact scenes for the C-IPC hot path.

The reference's large input meshes are absent (.MISSING_LARGE_BLOBS), so every configuration is
generated here, seeded, in fp64 / int32.  A scene is a plain dict:

    X, X0        (nV,3) float64   current / rest positions (MESH_NODE X, nodeAttr.x0)
    BN           (nBN,)  int32    boundaryNode   (ordering: Library/FEM/Shell/IMPLICIT_EULER.h:224-300)
    BE           (nBE,2) int32    boundaryEdge   (lexicographic oriented edges, Library/Utils/MESHIO.h:792-826;
                                                  then rod edges)
    BT           (nBT,3) int32    boundaryTri
    nRod, codim  number of rod edges at the tail of BE; codimBNStartInd (2,)
    DBC          (nV,)   uint8    Dirichlet flags (std::vector<bool> DBCb)
    NNX          (k,2)   int32    NNExclusion as (key, member) pairs
    BNArea/BEArea/BTArea          area weights (only read when elasticIPC)
    dHat2, xi, kappa              contact parameters (Library/FEM/Shell/DISCRETE_SHELL.h:555-577)
    p            (nV,3) float64   a CCD search direction
"""
import math
import numpy as np


# ----------------------------------------------------------------------------- mesh pieces
def grid_mesh(n, side=1.0):
    """(n+1)^2 vertices in the xz-plane, 2 n^2 triangles (square<N>.obj-like)."""
    lin = np.linspace(-0.5 * side, 0.5 * side, n + 1)
    gx, gz = np.meshgrid(lin, lin, indexing="ij")
    V = np.stack([gx.ravel(), np.zeros((n + 1) ** 2), gz.ravel()], axis=1)
    i, j = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
    v00 = (i * (n + 1) + j).ravel()
    v01 = v00 + 1
    v10 = v00 + (n + 1)
    v11 = v10 + 1
    F = np.concatenate([np.stack([v00, v01, v11], 1), np.stack([v00, v11, v10], 1)], axis=0)
    # interleave the two triangles of each quad so that triangle order is spatially coherent
    F = F.reshape(2, -1, 3).transpose(1, 0, 2).reshape(-1, 3)
    return V, F.astype(np.int32)


def uv_sphere(nu, nv, r):
    """Closed UV sphere: 2 + (nu-1)*nv vertices, 2*nv*(nu-1) triangles."""
    th = np.pi * np.arange(1, nu) / nu
    ph = 2 * np.pi * np.arange(nv) / nv
    T, P = np.meshgrid(th, ph, indexing="ij")
    ring = np.stack([np.sin(T) * np.cos(P), np.cos(T), np.sin(T) * np.sin(P)], -1).reshape(-1, 3)
    V = np.concatenate([[[0, 1, 0]], ring, [[0, -1, 0]]], 0) * r
    F = []
    top, bot = 0, 1 + (nu - 1) * nv
    idx = lambda a, b: 1 + a * nv + (b % nv)
    for b in range(nv):
        F.append([top, idx(0, b + 1), idx(0, b)])
        F.append([bot, idx(nu - 2, b), idx(nu - 2, b + 1)])
    for a in range(nu - 2):
        for b in range(nv):
            F.append([idx(a, b), idx(a, b + 1), idx(a + 1, b + 1)])
            F.append([idx(a, b), idx(a + 1, b + 1), idx(a + 1, b)])
    return V, np.asarray(F, np.int32)


def rot_y(V, deg):
    a = math.radians(deg)
    c, s = math.cos(a), math.sin(a)
    R = np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]])
    return V @ R.T


# ----------------------------------------------------------------------------- boundary primitives
def surface_primitives(X, F):
    """Restates Find_Surface_Primitives_And_Compute_Area (Library/Utils/MESHIO.h:768-834), vectorised.

    boundaryTri = F in order; boundaryEdge = one entry per undirected edge, oriented like the FIRST
    triangle that mentions it, listed in lexicographic (a,b) order (std::map<VECTOR<int,2>>);
    boundaryNode = ascending vertices used by F.  Areas follow :777,:818,:825.
    """
    F = np.asarray(F, np.int64)
    T = len(F)
    area = 0.5 * np.linalg.norm(np.cross(X[F[:, 1]] - X[F[:, 0]], X[F[:, 2]] - X[F[:, 0]]), axis=1)
    a = np.stack([F[:, 0], F[:, 1], F[:, 2]], 1).ravel()  # directed edges in visiting order
    b = np.stack([F[:, 1], F[:, 2], F[:, 0]], 1).ravel()
    lo, hi = np.minimum(a, b), np.maximum(a, b)
    key = lo * (int(X.shape[0]) + 1) + hi
    order = np.argsort(key, kind="stable")
    ks = key[order]
    first = np.ones(len(ks), bool)
    first[1:] = ks[1:] != ks[:-1]
    grp = np.cumsum(first) - 1
    rep = order[first]  # first occurrence (stable sort keeps visiting order inside a group)
    ea, eb = a[rep], b[rep]
    earea = np.zeros(len(rep))
    np.add.at(earea, grp, np.repeat(area / 3, 3)[order])
    lex = np.lexsort((eb, ea))
    BE = np.stack([ea[lex], eb[lex]], 1).astype(np.int32)
    BEArea = earea[lex] / 2
    narea = np.zeros(X.shape[0])
    np.add.at(narea, F.ravel(), np.repeat(area / 3, 3))
    used = np.zeros(X.shape[0], bool)
    used[F.ravel()] = True
    BN = np.nonzero(used)[0].astype(np.int32)
    return BN, BE, F.astype(np.int32), narea[BN], BEArea, area / 2


def oipc_kappa(dHat2, mean_mass=1e-4, stiff_mult=1.0, dim=3):
    """Initialize_EIPC<T,false> (Library/FEM/Shell/DISCRETE_SHELL.h:555-577)."""
    d1 = 1.0e-16
    t2 = d1 - dHat2
    H_b = (math.log(d1 / dHat2) * -2.0 - t2 * 4.0 / d1) + 1.0 / (d1 * d1) * (t2 * t2)
    k0 = stiff_mult * 1.0e11 * mean_mass * dim / (4.0e-16 * H_b)
    return np.array([k0, 100 * k0, 0.0])


def assemble(parts, rods=(), particles=None, dHat=1e-3, xi=0.0, seed=2, p_scale=None, nnx=None):
    """parts: list of (V, F, is_dbc).  rods: list of (V_polyline, is_dbc).  particles: (V, is_dbc) or None.

    Follows Library/FEM/Shell/IMPLICIT_EULER.h:224-300: surface primitives of all triangles, then
    rod edges appended to BE, codimBNStartInd[0], rod nodes ascending, codimBNStartInd[1], particles.
    """
    Vs, Fs, dbc = [], [], []
    off = 0
    for V, F, fixed in parts:
        Vs.append(V); Fs.append(F + off); dbc.append(np.full(len(V), 1 if fixed else 0, np.uint8))
        off += len(V)
    rod_edges = []
    for V, fixed in rods:
        n = len(V)
        Vs.append(V); dbc.append(np.full(n, 1 if fixed else 0, np.uint8))
        rod_edges.append(np.stack([np.arange(n - 1), np.arange(1, n)], 1) + off)
        off += n
    part_ids = np.zeros(0, np.int32)
    if particles is not None:
        V, fixed = particles
        Vs.append(V); dbc.append(np.full(len(V), 1 if fixed else 0, np.uint8))
        part_ids = (np.arange(len(V)) + off).astype(np.int32)
        off += len(V)
    X = np.concatenate(Vs, 0).astype(np.float64)
    F = np.concatenate(Fs, 0).astype(np.int32) if Fs else np.zeros((0, 3), np.int32)
    BN, BE, BT, BNA, BEA, BTA = surface_primitives(X, F)
    nRod = 0
    if rod_edges:
        RE = np.concatenate(rod_edges, 0).astype(np.int32)
        nRod = len(RE)
        BE = np.concatenate([BE, RE], 0)
        rlen = np.linalg.norm(X[RE[:, 0]] - X[RE[:, 1]], axis=1)
        rarea = rlen * math.pi * xi / 6
        BEA = np.concatenate([BEA, rarea / 2])
        rn = np.unique(RE)
        codim0 = len(BN)
        na = np.zeros(X.shape[0]); np.add.at(na, RE.ravel(), np.repeat(rarea / 2, 2))
        BN = np.concatenate([BN, rn.astype(np.int32)])
        BNA = np.concatenate([BNA, na[rn]])
    else:
        codim0 = len(BN)
    codim1 = len(BN)
    BN = np.concatenate([BN, part_ids]).astype(np.int32)
    rng = np.random.default_rng(seed)
    gap = xi + 0.5 * dHat
    if p_scale is None:
        p_scale = gap
    sc = dict(X=X, X0=X.copy(), BN=BN, BE=BE.astype(np.int32), BT=BT, nRod=nRod, codim=(codim0, codim1),
              DBC=np.concatenate(dbc), NNX=(np.zeros((0, 2), np.int32) if nnx is None else np.asarray(nnx, np.int32)),
              BNArea=BNA, BEArea=BEA, BTArea=BTA, dHat2=dHat * dHat, xi=xi, kappa=oipc_kappa(dHat * dHat),
              p=rng.normal(0.0, 0.2 * p_scale, size=X.shape),
              # raw element lists the boundary primitives were built from (inputs of cipc_build_boundary)
              F=F, rodE=(np.concatenate(rod_edges, 0).astype(np.int32) if rod_edges else np.zeros((0, 2), np.int32)), particles=part_ids)
    return sc


# ----------------------------------------------------------------------------- configurations
def cloth_stack(n, layers, dHat=1e-3, xi=0.0, seed=0, rest_flat=True):
    """cfg5: `layers` n x n grids, gap xi+0.5*dHat, layer i rotated by (i+1)*90/(L+1) deg about y and
    jittered N(0,(0.05 dHat)^2) (pattern of Projects/FEMShell/18_sphere_on_cloth_stack.py:47-49)."""
    rng = np.random.default_rng(seed)
    gap = xi + 0.5 * dHat
    V0, F = grid_mesh(n)
    parts, rest = [], []
    for i in range(layers):
        V = rot_y(V0, (i + 1) * 90.0 / (layers + 1))
        V[:, 1] = i * gap
        rest.append(V.copy())
        parts.append((V + rng.normal(0, 0.05 * dHat, V.shape), F, False))
    sc = assemble(parts, dHat=dHat, xi=xi, seed=seed + 2)
    sc["X0"] = np.concatenate(rest, 0)
    # CCD direction: alternate layers approach each other by 0.5*gap plus noise
    nv = len(V0)
    lay = np.arange(len(sc["X"])) // nv
    sc["p"][:, 1] += np.where(lay % 2 == 0, 0.5, -0.5) * 0.5 * gap
    sc["name"] = "cloth_stack_%dx%dx%d" % (layers, n, n)
    return sc


def cloth_on_sphere(n=112, dHat=1e-3, xi=0.0, seed=1, draped=False):
    """cfg1 (n=112: 25 088 triangles, square113.obj-like) / cfg2 (n=207, draped): a cloth over a
    r=0.3 UV sphere (~2.5K triangles, DBC) and a 20x20 floor (DBC), after
    Projects/FEMShell/6_cloth_on_rotating_sphere.py:32-45."""
    rng = np.random.default_rng(seed)
    V, F = grid_mesh(n)
    r = 0.3
    Vs, Fs = uv_sphere(36, 36, r)
    Vf, Ff = grid_mesh(20, side=2.0)
    Vf[:, 1] = -r - 0.2
    rest = V.copy()
    gap = xi + 0.6 * dHat
    if not draped:
        V[:, 1] = r + 1.5 * (dHat + xi)
    else:
        # sphere-wrapped cap: cloth lies `gap` above the sphere where it covers it, hangs flat at the
        # equator height elsewhere, and one strip is folded back over itself at distance `gap`
        rr = np.sqrt(V[:, 0] ** 2 + V[:, 2] ** 2)
        R = r + gap
        cap = np.sqrt(np.maximum(R * R - rr * rr, 0.0))
        V[:, 1] = np.where(rr < R, cap, 0.0)
        fold = V[:, 0] > 0.4
        V[fold, 0] = 0.8 - V[fold, 0]
        V[fold, 1] += gap
        V += rng.normal(0, 0.1 * dHat, V.shape)
    sc = assemble([(V, F, False), (Vs, Fs, True), (Vf, Ff, True)], dHat=dHat, xi=xi, seed=seed + 2)
    sc["X0"][:len(rest)] = rest
    sc["name"] = "cloth_on_sphere_n%d%s" % (n, "_draped" if draped else "")
    return sc


def noodles(nrod=25, nseg=200, dHat=5e-4, xi=1e-3, seed=3):
    """cfg3: nrod x nrod discrete rods of nseg segments (length 0.4) packed in crossing layers at centre
    spacing xi+0.5*dHat above a DBC plate (Projects/FEMShell/10_noodles.py:26-31,44)."""
    rng = np.random.default_rng(seed)
    gap = xi + 0.5 * dHat
    rods = []
    t = np.linspace(0, 0.4, nseg + 1)
    for i in range(nrod):  # layer i at height (i+1)*gap; even layers run along x, odd layers along z
        for j in range(nrod):
            off = (j - nrod / 2) * gap
            if i % 2 == 0:
                V = np.stack([t - 0.2, np.full_like(t, (i + 1) * gap), np.full_like(t, off)], 1)
            else:
                V = np.stack([np.full_like(t, off), np.full_like(t, (i + 1) * gap), t - 0.2], 1)
            V = V + rng.normal(0, 0.05 * dHat, V.shape)
            rods.append((V, False))
    # obstacle under the pile (the reference uses bowl.obj, absent here): a flat ~2.3K-triangle DBC
    # plate 1.04*gap below the lowest rod layer, i.e. inside the activation distance
    Vb, Fb = grid_mesh(34, side=0.5)
    Vb = rot_y(Vb, 9.0)
    Vb[:, 1] = -0.04 * gap
    sc = assemble([(Vb, Fb, True)], rods=rods, dHat=dHat, xi=xi, seed=seed + 2)
    sc["name"] = "noodles_%dx%dx%d" % (nrod, nrod, nseg)
    return sc


def granules(npart=50000, dHat=1e-3, xi=2e-3, seed=4, cloth_n=63):
    """cfg4: particles on a jittered lattice of spacing xi+0.5*dHat over an 8K-triangle cloth
    (Projects/FEMShell/22_granules_on_cloth.py:45-47,101-103)."""
    rng = np.random.default_rng(seed)
    gap = xi + 0.5 * dHat
    m = int(round((npart / 10) ** 0.5))
    gx, gy, gz = np.meshgrid(np.arange(m), np.arange(max(1, npart // (m * m))), np.arange(m), indexing="ij")
    P = np.stack([gx.ravel(), gy.ravel(), gz.ravel()], 1).astype(np.float64) * gap
    P[:, 0] -= P[:, 0].mean(); P[:, 2] -= P[:, 2].mean()
    P[:, 1] += gap
    P += rng.normal(0, 0.05 * dHat, P.shape)
    V, F = grid_mesh(cloth_n, side=max(1.0, 1.2 * m * gap))
    sc = assemble([(V, F, False)], particles=(P, False), dHat=dHat, xi=xi, seed=seed + 2)
    sc["name"] = "granules_%d" % len(P)
    return sc


def mixed_small(seed=5):
    """A small scene with every codimension (cloth + DBC obstacle + rods + particles + a stitch
    exclusion) used by the parity tests to hit all seven stencil kinds and all filters."""
    rng = np.random.default_rng(seed)
    dHat, xi = 1e-3, 5e-4
    gap = xi + 0.5 * dHat
    V0, F = grid_mesh(12, side=0.12)
    parts = []
    for i in range(3):
        V = rot_y(V0, 17.0 * i)
        V[:, 1] = i * gap
        parts.append((V + rng.normal(0, 0.05 * dHat, V.shape), F, i == 0))
    # a parallel copy (no rotation) to provoke mollified (near-parallel) edge pairs
    V = rot_y(V0, 34.0); V[:, 1] = 3 * gap; V[:, 0] += 0.002
    parts.append((V + rng.normal(0, 0.02 * dHat, V.shape), F, False))
    rods = []
    t = np.linspace(-0.05, 0.05, 21)
    for k in range(4):
        Vr = np.stack([t, np.full_like(t, 4 * gap + (k % 2) * gap), np.full_like(t, (k - 1.5) * gap * 1.5)], 1)
        if k % 2:
            Vr = Vr[:, [2, 1, 0]]
        rods.append((Vr + rng.normal(0, 0.05 * dHat, Vr.shape), False))
    g = np.arange(6)
    gx, gy, gz = np.meshgrid(g, np.arange(2), g, indexing="ij")
    P = np.stack([gx.ravel(), gy.ravel(), gz.ravel()], 1) * gap
    P[:, 0] -= 2.5 * gap; P[:, 2] -= 2.5 * gap; P[:, 1] += 6.1 * gap
    P = P + rng.normal(0, 0.05 * dHat, P.shape)
    nv = len(V0)
    nnx = [[nv + 5, 2 * nv + 5], [nv + 5, 2 * nv + 6], [2 * nv + 5, nv + 5], [2 * nv + 6, nv + 5]]
    sc = assemble(parts, rods=rods, particles=(P, False), dHat=dHat, xi=xi, seed=seed + 2, nnx=nnx)
    sc["name"] = "mixed_small"
    return sc


CONFIGS = {
    "cfg1": lambda: cloth_on_sphere(112, draped=False),
    "cfg2": lambda: cloth_on_sphere(207, draped=True),
    "cfg3": lambda: noodles(),
    "cfg4_50k": lambda: granules(50000),
    "cfg4_500k": lambda: granules(500000),
    "cfg5_250k": lambda: cloth_stack(112, 10),
    "cfg5_1m": lambda: cloth_stack(224, 10),
    "cfg5_4m": lambda: cloth_stack(354, 16),
}
