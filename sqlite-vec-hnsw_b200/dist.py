"""Rowid-range sharding of one slab over the GPUs of a box.

One process per GPU (torch.distributed, NCCL over NVLink).  Rank g holds the
contiguous rowid range shard_range(n, g, G); every rowid of shard g is below
every rowid of shard g+1, so the global (distance, rowid) order restricted to a
shard is the shard's local order.  A query batch is scanned locally, the G local
top-k lists are exchanged with ONE all-gather per batch (k * 12 bytes per query
per rank) and merged by vecgpu_merge_device.  The reference has no counterpart
(single process); results are identical to a single slab holding all rows.
"""
import numpy as np


def shard_range(n_rows, rank, world):
    """[lo, hi) row positions of `rank`: near-equal contiguous ranges, remainder to the low ranks."""
    base, rem = divmod(int(n_rows), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def pack_local(rowids, dists):
    """[nq,k] i64 + [nq,k] f32 -> one [nq,k,2] i64 tensor (distance bits in the low word) so that a
    single all_gather moves both."""
    import torch

    d = dists.contiguous().view(torch.int32).to(torch.int64) & 0xFFFFFFFF
    return torch.stack([rowids, d], dim=-1).contiguous()


def unpack_gathered(buf):
    """[G,nq,k,2] i64 -> ([G,nq,k] i64 rowids, [G,nq,k] f32 dists)."""
    import torch

    rowids = buf[..., 0].contiguous()
    dists = buf[..., 1].to(torch.int32).contiguous().view(torch.float32)
    return rowids, dists


def all_gather_topk(rowids, dists, group=None):
    """One collective per batch.  Works on NCCL (cuda tensors) and gloo (cpu tensors)."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    local = pack_local(rowids, dists)
    # concatenation along dim 0 (the layout both NCCL and gloo accept), viewed as [world, nq, k, 2]
    out = torch.empty((world * local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, local, group=group)
    return unpack_gathered(out.view((world,) + tuple(local.shape)))


class ShardedSlab:
    """A slab sharded by rowid range across the ranks of the default process group."""

    def __init__(self, vec0, vec_type, dims, n_rows_total, rank, world, device):
        self.vec0 = vec0
        self.rank, self.world, self.device = rank, world, device
        self.n_total = int(n_rows_total)
        self.lo, self.hi = shard_range(self.n_total, rank, world)
        self.slab = vec0.Slab(vec_type, dims, capacity_hint=self.hi - self.lo, device=device)

    def fill_synthetic(self, seed, kind, first_rowid=1):
        # rank g generates exactly its rowid range of the global corpus
        self.slab.fill_synthetic(seed, self.hi - self.lo, first_rowid=first_rowid + self.lo, kind=kind)

    def load_global(self, vectors, first_rowid=1):
        """Every rank passes the same global array; each keeps its range (test helper)."""
        v = np.ascontiguousarray(vectors)[self.lo : self.hi]
        self.slab.load(v, np.arange(first_rowid + self.lo, first_rowid + self.hi, dtype="<i8"))

    def knn_device(self, d_queries, k, metric):
        """Device tensors in/out; global top-k on every rank."""
        r, d = self.slab.knn_device(d_queries, k, metric)
        if self.world == 1:
            return r, d
        gr, gd = all_gather_topk(r, d)
        return self.vec0.merge_device(gr, gd)

    def knn(self, queries_pinned, k, metric):
        """Host (pinned) queries in, host results out — the end-to-end call."""
        import torch

        dq = queries_pinned.to(f"cuda:{self.device}", non_blocking=True)
        r, d = self.knn_device(dq, k, metric)
        return r.cpu(), d.cpu()

    def close(self):
        self.slab.close()
