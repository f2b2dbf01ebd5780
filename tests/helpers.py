"""Shared helpers of the parity tests (test infrastructure)."""
import numpy as np


def sort_cs(cs, info=None):
    """constraint sets are compared as sorted int4 multisets (the reference's order is unordered_set-dependent)"""
    cs = np.asarray(cs).reshape(-1, 4)
    order = np.lexsort(cs.T[::-1])
    if info is None:
        return cs[order]
    return cs[order], np.asarray(info)[order]


def block_dims(cs):
    cs = np.asarray(cs)
    return np.where((cs[:, 0] >= 0) | (cs[:, 3] >= 0), 12, np.where(cs[:, 2] >= 0, 9, 6))


def split_blocks(cs, vals):
    """dense blocks (list of (n,n)) from the flat triplet value stream in constraint order"""
    out, off = [], 0
    for n in block_dims(cs):
        out.append(vals[off:off + n * n].reshape(n, n))
        off += n * n
    assert off == len(vals)
    return out


def max_block_rel_err(cs, vals_a, vals_b):
    """max over constraints of ||Ha - Hb||_F / ||Hb||_F (SURVEY 8(d) parity gate for the Hessian)"""
    worst = 0.0
    off = 0
    for n in block_dims(cs):
        a = vals_a[off:off + n * n]; b = vals_b[off:off + n * n]
        nb = np.linalg.norm(b)
        if nb > 0:
            worst = max(worst, np.linalg.norm(a - b) / nb)
        off += n * n
    return worst


def kinds(cs):
    return ["".join("+" if x >= 0 else "-" for x in c) for c in np.asarray(cs)]
