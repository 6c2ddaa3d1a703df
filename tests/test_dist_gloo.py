"""N>1 host logic on CPU: world_size-2 gloo processes exercise shard_range, the packed all-gather of
local top-k lists and the global (distance, rowid) merge order.  The local scan is played by the CPU
oracle here (test infrastructure); on the GPU box the same code path runs vecgpu_knn_device + NCCL +
vecgpu_merge_device (tests/test_gpu_parity.py::test_cross_shard_merge_kernel, bench.py --gpus N)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import COSINE, F32, I8, L2, random_rows


def test_shard_range_partitions():
    from sqlite_vec_hnsw_b200.dist import shard_range

    for n in (0, 1, 7, 8, 10_000_000, 500_000_001):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for a, b in zip(spans, spans[1:]):
                assert a[1] == b[0]  # contiguous rowid ranges: shard g entirely below shard g+1
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, elem, dims, metric, k, n, seed, out_dir):
    import oracle
    from sqlite_vec_hnsw_b200 import dist as vdist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    v = random_rows(elem, n, dims, seed=seed, ties=True)
    q = random_rows(elem, 3, dims, seed=seed + 1, ties=True)
    lo, hi = vdist.shard_range(n, rank, world)
    rowids = np.arange(1 + lo, 1 + hi, dtype="<i8")
    r, d, c = oracle.knn(elem, dims, v[lo:hi], q, k, metric, rowids=rowids)
    r[r < 0] = np.iinfo(np.int64).max  # device-API padding convention
    gr, gd = vdist.all_gather_topk(torch.from_numpy(r), torch.from_numpy(d))
    assert gr.shape == (world, 3, k) and gd.dtype == torch.float32
    # this rank's slice must round-trip bit for bit through pack/unpack
    assert torch.equal(gr[rank], torch.from_numpy(r)) and torch.equal(gd[rank].view(torch.int32), torch.from_numpy(d).view(torch.int32))
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), gr=gr.numpy(), gd=gd.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("elem,dims,metric", [(F32, 24, COSINE), (I8, 16, L2)])
def test_two_rank_gather_and_merge_order(tmp_path, orc, elem, dims, metric):
    world, k, n, seed = 2, 12, 501, 7
    mp.start_processes(_worker, args=(world, _free_port(), elem, dims, metric, k, n, seed, str(tmp_path)), nprocs=world,
                       join=True, start_method="spawn")
    v = random_rows(elem, n, dims, seed=seed, ties=True)
    q = random_rows(elem, 3, dims, seed=seed + 1, ties=True)
    er, ed, _ = orc.knn(elem, dims, v, q, k, metric)
    for rank in range(world):
        z = np.load(tmp_path / f"rank{rank}.npz")
        gr, gd = z["gr"], z["gd"]
        for qi in range(3):
            rr, dd = gr[:, qi, :].reshape(-1), gd[:, qi, :].reshape(-1)
            order = np.lexsort((rr, dd))[:k]  # (distance, rowid) ascending == what xmerge_kernel implements
            assert np.array_equal(rr[order], er[qi]) and np.array_equal(dd[order].view("<u4"), ed[qi].view("<u4"))
