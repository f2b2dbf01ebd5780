"""Seeded scenes and summaries shared by tests/golden/make_golden_drivers.py (which runs the reference's own
drivers on them) and the tests that compare the oracle / the CUDA path with the resulting fixture."""
import numpy as np

from helpers import block_dims


def _scenes():
    from codim_ipc_b200 import scenes
    return scenes


CASES = {
    "mixed_small": lambda: _scenes().mixed_small(),
    "stack_16x3": lambda: _scenes().cloth_stack(16, 3),
    "stack_20x3_xi": lambda: _scenes().cloth_stack(20, 3, xi=1e-3),
    "sphere_small": lambda: _scenes().cloth_on_sphere(40, draped=True),
    "noodles_small": lambda: _scenes().noodles(4, 40),
    "granules_small": lambda: _scenes().granules(2000, cloth_n=15),
}
# lagged-friction inputs: Xn = X - N(0, mag^2) (relative sliding below / above eps_v h), eps_v h = 1e-5
FRICTION = dict(seed=7, mags=(1e-7, 1e-4), epsvh2=1e-10, mu=0.4)


def quad_vector(n):
    return np.cos(1.0 + 0.37 * np.arange(n))


def block_summaries(cs, rows, cols, vals):
    """per constraint block: (Frobenius norm, v^T H v, index checksum) with the triplet stream in constraint order"""
    dims = block_dims(cs)
    out = np.zeros((len(cs), 3))
    off = 0
    for i, n in enumerate(dims):
        H = vals[off:off + n * n].reshape(n, n)
        v = quad_vector(n)
        out[i, 0] = np.linalg.norm(H)
        out[i, 1] = v @ H @ v
        out[i, 2] = float(np.sum(rows[off:off + n * n].astype(np.int64) * 3 + cols[off:off + n * n].astype(np.int64) * 7) % 1000003)
        off += n * n
    assert off == len(vals)
    return out


def used_closest(fcs, cp):
    """the reference leaves closestPoint components it never reads uninitialised (PP: both, PE: the second)"""
    fcs = np.asarray(fcs); cp = np.array(cp, copy=True)
    pp = (fcs[:, 0] < 0) & (fcs[:, 2] < 0)
    pe = (fcs[:, 0] < 0) & (fcs[:, 2] >= 0) & (fcs[:, 3] < 0)
    cp[pp] = 0.0
    cp[pe, 1] = 0.0
    return cp
