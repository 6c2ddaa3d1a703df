"""Rowid-range sharding of one slab over the GPUs of a box, one process per GPU.

Rank g holds the contiguous rowid range shard_range(n, g, G); every rowid of shard g is below every rowid of shard
g+1, so the global (distance, rowid) order restricted to a shard is the shard's local order.  A query batch is scanned
locally and the G local top-k lists are exchanged and merged.  Two exchange implementations:

  * "p2p" (default on CUDA): the C entry points vecgpu_shard_knn / vecgpu_shard_knn_device / vecgpu_xchg_merge_device —
    each rank writes its packed top-k straight into every peer's gather buffer over NVLink (CUDA IPC mapping between
    the torchrun processes), the receiving GPU merges when the flags are in (csrc/xchg.cuh).  torch.distributed is only
    used ONCE, to hand the 128-byte endpoint handles around.
  * "nccl" (VECGPU_EXCHANGE=nccl, and the gloo/CPU tests): ONE all_gather_into_tensor of the packed local top-k per
    batch followed by vecgpu_merge_device.  Kept as the reference form of the exchange; both give identical results.

The reference has no counterpart (single process); results are identical to a single slab holding all rows.
The ONE-process form (one handle, all GPUs, no launcher) is vec0.ShardedSlab / vecgpu_sharded_*.
"""
import os

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


def gather_handles(handle, world, device=None, group=None):
    """All ranks' 128-byte endpoint handles, rank order: the one use of torch.distributed on the p2p path."""
    import torch
    import torch.distributed as dist

    t = torch.from_numpy(np.ascontiguousarray(handle, dtype="u1").copy())
    if device is not None:
        t = t.to(device)
    out = torch.empty((world * t.numel(),), dtype=torch.uint8, device=t.device)
    dist.all_gather_into_tensor(out, t, group=group)
    return out.cpu().numpy().reshape(world, -1)


class ShardedSlab:
    """A slab sharded by rowid range across the ranks of the default process group (one process per GPU)."""

    def __init__(self, vec0, vec_type, dims, n_rows_total, rank, world, device, exchange=None, max_queries=1024, max_k=128):
        self.vec0 = vec0
        self.rank, self.world, self.device = rank, world, device
        self.n_total = int(n_rows_total)
        self.lo, self.hi = shard_range(self.n_total, rank, world)
        self.slab = vec0.Slab(vec_type, dims, capacity_hint=self.hi - self.lo, device=device)
        self.exchange = exchange or os.environ.get("VECGPU_EXCHANGE", "p2p")
        self.xchg = None
        if world > 1 and self.exchange == "p2p":
            import torch
            import torch.distributed as dist

            self.xchg = vec0.Exchange(device, rank, world, max_queries=max_queries, max_k=max_k)
            handles = gather_handles(self.xchg.export_handle(), world, device=torch.device("cuda", device))
            self.xchg.attach_ipc(handles)
            dist.barrier()  # every rank has mapped every peer before the first push

    def fill_synthetic(self, seed, kind, first_rowid=1):
        # rank g generates exactly its rowid range of the global corpus
        self.slab.fill_synthetic(seed, self.hi - self.lo, first_rowid=first_rowid + self.lo, kind=kind)

    def load_global(self, vectors, first_rowid=1):
        """Every rank passes the same global array; each keeps its range (test helper)."""
        v = np.ascontiguousarray(vectors)[self.lo : self.hi]
        self.slab.load(v, np.arange(first_rowid + self.lo, first_rowid + self.hi, dtype="<i8"))

    def merge_device(self, r, d, stream=None):
        """This rank's [nq,k] device results -> global top-k on every rank (one exchange)."""
        if self.world == 1:
            return r, d
        if self.xchg is not None:
            return self.xchg.merge_device(r, d, stream=stream)
        gr, gd = all_gather_topk(r, d)
        return self.vec0.merge_device(gr, gd, stream=stream)

    def knn_device(self, d_queries, k, metric, stream=None):
        """Device tensors in/out; global top-k on every rank."""
        if self.world > 1 and self.xchg is not None:
            return self.xchg.shard_knn_device(self.slab, d_queries, k, metric, stream=stream)
        r, d = self.slab.knn_device(d_queries, k, metric, stream=stream)
        return self.merge_device(r, d, stream=stream)

    def knn(self, queries, k, metric):
        """Host queries in, host results out — the end-to-end call.  p2p: ONE C call (vecgpu_shard_knn: pinned H2D, scan,
        push + merge over NVLink, D2H).  nccl: torch glue around the all-gather."""
        if self.world == 1:
            r, d, _ = self.slab.knn(queries.numpy() if hasattr(queries, "numpy") else queries, k, metric)
            return r, d
        if self.xchg is not None:
            r, d, _ = self.xchg.shard_knn(self.slab, queries.numpy() if hasattr(queries, "numpy") else queries, k, metric)
            return r, d
        import torch

        q = queries if hasattr(queries, "to") else torch.from_numpy(np.ascontiguousarray(queries))
        dq = q.to(f"cuda:{self.device}", non_blocking=True)
        r, d = self.knn_device(dq, k, metric)
        return r.cpu(), d.cpu()

    def close(self):
        if self.xchg is not None:
            self.xchg.close()
            self.xchg = None
        self.slab.close()
