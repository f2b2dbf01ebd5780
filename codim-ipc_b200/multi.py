"""Multi-GPU glue for the contact path: one process per GPU, `torch.distributed` for the plumbing.

Partitioning: positions, rest positions, search direction and the primitive
lists are replicated; every rank owns a slab of voxel layers (balanced on the number of hash entries),
builds the cell-sorted hash of that slab only and enumerates the candidate pairs whose min-corner voxel
lies in it.  Pairs are therefore disjoint across ranks and their union is the single-GPU candidate set.  The only
exchanges the path needs are

    step size        all-reduce(min) of one f64                 (IPC.h:2166,2241: min over all pairs)
    barrier energy   all-reduce(sum) of one f64                 (IPC.h:940)
    barrier gradient all-reduce(sum) of 3*nV f64                (IPC.h:1034-1042)
    constraint set   all-gather of (count, int4 payload); PP/PE stencils that were produced on several
                     ranks are merged by adding their multiplicities (IPC.h:599-654 keys on the raw tuple)
    Hessian triplets stay on the rank that owns the constraint; an all-gather of the per-rank counts gives every rank the
                     offset of its slice in the global (row, col, value) stream, and `gather_triplets` assembles that
                     stream on every rank when one process must own it

Nothing here computes contact terms: it only combines per-rank results.
"""
import numpy as np


def merge_constraint_sets(parts):
    """Combine per-rank constraint sets into the global one (host logic; also used with all_gather).

    PT / EE / mollified stencils are unique to the rank that enumerated the pair.  A PP / PE stencil
    (c0 < 0 and c3 < 0, multiplicity -c3) with the same (c0,c1,c2) may come from several ranks: the
    reference counts every generating pair, so multiplicities add."""
    parts = [np.asarray(p, np.int32).reshape(-1, 4) for p in parts]
    allc = np.concatenate(parts, 0) if parts else np.zeros((0, 4), np.int32)
    dup = (allc[:, 0] < 0) & (allc[:, 3] < 0)
    keep = allc[~dup]
    d = allc[dup]
    if len(d) == 0:
        return keep
    order = np.lexsort((d[:, 2], d[:, 1], d[:, 0]))
    d = d[order]
    first = np.ones(len(d), bool)
    first[1:] = np.any(d[1:, :3] != d[:-1, :3], axis=1)
    grp = np.cumsum(first) - 1
    mult = np.zeros(grp[-1] + 1, np.int64)
    np.add.at(mult, grp, -d[:, 3].astype(np.int64))
    out = d[first].copy()
    out[:, 3] = -mult.astype(np.int32)
    return np.concatenate([keep, out], 0)


class DistContact:
    """Collectives of the contact path over an initialised torch.distributed process group
    (backend "nccl" on GPUs; "gloo" in the CPU tests of the host logic)."""

    def __init__(self, group=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.group = torch, dist, group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")

    def min_step(self, alpha):
        t = self.torch.tensor([alpha], dtype=self.torch.float64, device=self.device)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MIN, group=self.group)
        return float(t.item())

    def sum_scalar(self, v):
        t = self.torch.tensor([v], dtype=self.torch.float64, device=self.device)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM, group=self.group)
        return float(t.item())

    def sum_gradient(self, g):
        """g: torch tensor (device) or numpy (nV,3); summed over ranks in place / returned"""
        if isinstance(g, np.ndarray):
            t = self.torch.from_numpy(np.ascontiguousarray(g)).to(self.device)
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM, group=self.group)
            return t.cpu().numpy()
        self.dist.all_reduce(g, op=self.dist.ReduceOp.SUM, group=self.group)
        return g

    def gather_constraints(self, cs):
        """all-gather per-rank (n_r,4) int32 sets and merge them (every rank gets the global set)"""
        torch, dist = self.torch, self.dist
        cs = np.ascontiguousarray(cs, np.int32).reshape(-1, 4)
        n = torch.tensor([len(cs)], dtype=torch.int64, device=self.device)
        counts = [torch.zeros_like(n) for _ in range(self.world)]
        dist.all_gather(counts, n, group=self.group)
        counts = [int(c.item()) for c in counts]
        mx = max(max(counts), 1)
        buf = torch.zeros((mx, 4), dtype=torch.int32, device=self.device)
        if len(cs):
            buf[:len(cs)] = torch.from_numpy(cs).to(self.device)
        outs = [torch.zeros_like(buf) for _ in range(self.world)]
        dist.all_gather(outs, buf, group=self.group)
        return merge_constraint_sets([o[:c].cpu().numpy() for o, c in zip(outs, counts)])


    def triplet_offsets(self, n_local):
        """-> (offset of this rank's triplets in the global stream, global count, per-rank counts): the ranks' slices are
        laid out in rank order, like the reference appends blocks in constraint order"""
        torch, dist = self.torch, self.dist
        n = torch.tensor([int(n_local)], dtype=torch.int64, device=self.device)
        counts = [torch.zeros_like(n) for _ in range(self.world)]
        dist.all_gather(counts, n, group=self.group)
        counts = [int(c.item()) for c in counts]
        return sum(counts[:self.rank]), sum(counts), counts

    def gather_triplets(self, trip):
        """trip: this rank's triplets as a torch uint8 tensor of 16-byte records (device for nccl, host for gloo) or a numpy
        structured array; returns the global stream (all ranks' slices in rank order) in the same kind of container"""
        torch, dist = self.torch, self.dist
        as_np = isinstance(trip, np.ndarray)
        t = torch.from_numpy(np.ascontiguousarray(trip).view(np.uint8).reshape(-1)).to(self.device) if as_np else trip.reshape(-1)
        assert t.dtype == torch.uint8 and t.numel() % 16 == 0
        off, total, counts = self.triplet_offsets(t.numel() // 16)
        mx = max(max(counts), 1) * 16
        pad = torch.zeros(mx, dtype=torch.uint8, device=self.device)
        pad[:t.numel()] = t
        outs = [torch.empty_like(pad) for _ in range(self.world)]
        dist.all_gather(outs, pad, group=self.group)
        full = torch.cat([o[:c * 16] for o, c in zip(outs, counts)])
        return full.cpu().numpy().view(trip.dtype) if as_np else full


def wrap_device_u8(ptr, nbytes, device_index):
    """torch uint8 view (no copy) of device memory handed out by the C ABI (cipc_dev_triplets)"""
    import torch

    class _Arr:
        __cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False), "version": 3, "strides": None}

    return torch.as_tensor(_Arr(), device=torch.device("cuda", device_index))


def wrap_device_f64(ptr, n, device_index):
    """torch view (no copy) of n doubles at a raw device pointer handed out by the C ABI
    (cipc_dev_gradient / cipc_dev_scalars), so that NCCL can reduce them in place."""
    import torch

    class _Arr:
        __cuda_array_interface__ = {"shape": (int(n),), "typestr": "<f8", "data": (int(ptr), False), "version": 3, "strides": None}

    return torch.as_tensor(_Arr(), device=torch.device("cuda", device_index))
