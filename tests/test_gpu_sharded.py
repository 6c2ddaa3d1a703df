"""Sharded exact KNN behind the C ABI (include/vecgpu.h "sharded slabs", csrc/xchg.cuh, csrc/xchg_host.inl).

  * vecgpu_sharded_* (ONE process, the form the Rust extension links): result == one slab holding all rows == the
    CPU oracle, bit for bit.  On a single-GPU box the shards are placed on the same device (devices=[0,0,0]) so the
    whole path — worker threads, peer-buffer push, flag wait, merge — still runs in the driver's 1-GPU test pass; with
    2+ GPUs it runs across devices over NVLink.
  * vecgpu_xchg_* + vecgpu_shard_knn (one process per GPU, CUDA IPC): two spawned processes, needs 2 GPUs.
  * BASELINE cfg4 in full: 500 M x bit[1024] Hamming k=10 over 8 GPUs against the CPU scan of all 500 M regenerated
    rows (needs 8 GPUs; skipped elsewhere).
"""
import os
import socket

import numpy as np
import pytest

from helpers import BIT, COSINE, F32, HAMMING, I8, L1, L2, PAIR_IDS, PAIRS, random_rows, same_bits

pytestmark = pytest.mark.gpu
os.environ.setdefault("VECGPU_XCHG_TIMEOUT_MS", "8000")


def _devices(gpu, n_shards):
    return [i % gpu for i in range(n_shards)]


@pytest.mark.parametrize("elem,metric", PAIRS, ids=PAIR_IDS)
@pytest.mark.parametrize("n_shards", [2, 3, 8])
def test_sharded_one_process_equals_single_slab(vg, orc, gpu, elem, metric, n_shards):
    dims = 72 if elem == BIT else 40
    n, k = 20_011, 10
    v = random_rows(elem, n, dims, seed=5, ties=(elem != F32))
    q = random_rows(elem, 7, dims, seed=6, ties=(elem != F32))
    er, ed, ec = orc.knn(elem, dims, v, q, k, metric)
    with vg.ShardedSlab(elem, dims, devices=_devices(gpu, n_shards), max_queries=64, max_k=32) as g:
        assert g.n_shards == n_shards
        g.load(v)
        assert g.count() == (n, n)
        for nq in (1, 7):  # single query (gather to shard 0) and a small batch
            r, d, c = g.knn(q[:nq], k, metric)
            assert np.array_equal(r, er[:nq]) and same_bits(d, ed[:nq]) and np.array_equal(c, ec[:nq])
        # repeated exchanges alternate the buffer halves: results must not change
        for _ in range(5):
            r2, d2, _ = g.knn(q[:1], k, metric)
            assert np.array_equal(r2, er[:1]) and same_bits(d2, ed[:1])


def test_sharded_sparse_rowids_deletes_and_upserts(vg, orc, gpu):
    dims, n, k = 24, 5000, 12
    rng = np.random.default_rng(3)
    v = random_rows(F32, n, dims, seed=31)
    rowids = (np.cumsum(rng.integers(1, 6, size=n)) - 4000).astype("<i8")
    q = random_rows(F32, 4, dims, seed=32)
    skip = np.zeros(n, dtype="u1")
    with vg.ShardedSlab(F32, dims, devices=_devices(gpu, 3), max_queries=16, max_k=16) as g:
        g.load(v, rowids)
        for i in rng.choice(n, size=300, replace=False):
            g.delete(int(rowids[i]))
            skip[i] = 1
        # in-place update of an existing row in every shard
        for i in (10, n // 2, n - 10):
            v[i] = q[0] * 0.5 + v[i] * 0.5
            g.upsert(int(rowids[i]), v[i].tobytes())
            skip[i] = 0
        # append beyond the last rowid (goes to the last shard)
        extra = random_rows(F32, 3, dims, seed=33)
        new_ids = np.array([rowids[-1] + 5, rowids[-1] + 6, rowids[-1] + 100], dtype="<i8")
        for rid, row in zip(new_ids, extra):
            g.upsert(int(rid), row.tobytes())
        vv = np.concatenate([v, extra])
        rr = np.concatenate([rowids, new_ids])
        ss = np.concatenate([skip, np.zeros(3, dtype="u1")])
        r, d, c = g.knn(q, k, COSINE)
        er, ed, ec = orc.knn(F32, dims, vv, q, k, COSINE, rowids=rr, skip=ss)
        assert np.array_equal(r, er) and same_bits(d, ed) and np.array_equal(c, ec)
        assert g.count() == (n + 3, n + 3 - int(ss.sum()))


def test_sharded_k_larger_than_rows_and_empty_shards(vg, orc, gpu):
    dims = 8
    v = random_rows(I8, 5, dims, seed=1)
    q = random_rows(I8, 2, dims, seed=2)
    with vg.ShardedSlab(I8, dims, devices=_devices(gpu, 8), max_queries=8, max_k=16) as g:  # 8 shards, 5 rows: three are empty
        g.load(v)
        r, d, c = g.knn(q, 9, L2)
        er, ed, ec = orc.knn(I8, dims, v, q, 9, L2)
        assert np.array_equal(r, er) and same_bits(d, ed) and np.array_equal(c, ec)
        assert list(c) == [5, 5] and np.all(r[:, 5:] == -1) and np.all(np.isinf(d[:, 5:]))


def test_sharded_batches_larger_than_the_gather_buffer(vg, orc, gpu):
    """nq > max_queries: the exchange runs in pieces (all-gather pacing), k > 32 takes the shared-memory merge."""
    dims, n, k = 32, 9000, 40
    v = random_rows(I8, n, dims, seed=8)
    q = random_rows(I8, 37, dims, seed=9)
    er, ed, ec = orc.knn(I8, dims, v, q, k, L2)
    with vg.ShardedSlab(I8, dims, devices=_devices(gpu, 2), max_queries=8, max_k=40) as g:
        g.load(v)
        r, d, c = g.knn(q, k, L2)
        assert np.array_equal(r, er) and same_bits(d, ed) and np.array_equal(c, ec)


def test_sharded_synthetic_prefix_of_cfg2(vg, orc, gpu):
    n, dims, k = 300_000, 768, 10
    q = orc.synth_rows(F32, 33, 1, 3, dims, 1)
    er, ed, ec = orc.knn_synth(F32, dims, 3, 1, n, 1, q, k, COSINE)
    with vg.ShardedSlab(F32, dims, devices=_devices(gpu, 4), max_queries=16, max_k=16) as g:
        g.fill_synthetic(3, n, kind=1)
        for qi in range(3):
            r, d, c = g.knn(q[qi], k, COSINE)
            assert np.array_equal(r[0], er[qi]) and same_bits(d[0], ed[qi])


# ------------------------------------------------------------------ one process per GPU (torchrun shape), CUDA IPC
def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _ipc_worker(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist

    import sqlite_vec_hnsw_b200 as vg
    from sqlite_vec_hnsw_b200 import dist as vdist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    dims, n, k = 48, 30_001, 10
    v = random_rows(F32, n, dims, seed=77)
    q = random_rows(F32, 5, dims, seed=78)
    res = {}
    for mode in ("p2p", "nccl"):
        if mode == "nccl":
            continue  # the gloo group of this test cannot carry CUDA tensors; NCCL mode is exercised by bench.py --gpus N
        sh = vdist.ShardedSlab(vg, F32, dims, n, rank, world, rank, exchange=mode, max_queries=16, max_k=16)
        sh.load_global(v)
        r, d = sh.knn(q, k, COSINE)  # vecgpu_shard_knn: host in, host out, every rank gets the global top-k
        dq = torch.from_numpy(q).cuda()
        r2, d2 = sh.knn_device(dq, k, COSINE)
        torch.cuda.synchronize()
        sh.xchg.check()
        # local results of two batches, one exchange each, repeated (buffer halves alternate)
        for _ in range(4):
            lr, ld = sh.slab.knn_device(dq, k, COSINE)
            r3, d3 = sh.merge_device(lr, ld)
        torch.cuda.synchronize()
        res[mode] = (r, d, r2.cpu().numpy(), d2.cpu().numpy(), r3.cpu().numpy(), d3.cpu().numpy())
        dist.barrier()
        sh.close()
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), **{f"{m}_{i}": a for m, t in res.items() for i, a in enumerate(t)})
    dist.barrier()
    dist.destroy_process_group()


def test_one_process_per_gpu_ipc_exchange(vg, orc, gpu, tmp_path):
    if gpu < 2:
        pytest.skip("needs 2 GPUs (one process per GPU)")
    import torch.multiprocessing as mp

    world = 2
    mp.start_processes(_ipc_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True, start_method="spawn")
    dims, n, k = 48, 30_001, 10
    v = random_rows(F32, n, dims, seed=77)
    q = random_rows(F32, 5, dims, seed=78)
    er, ed, _ = orc.knn(F32, dims, v, q, k, COSINE)
    for rank in range(world):
        z = np.load(tmp_path / f"rank{rank}.npz")
        for i in (0, 2, 4):
            assert np.array_equal(z[f"p2p_{i}"], er), (rank, i)
            assert same_bits(z[f"p2p_{i + 1}"], ed), (rank, i)


# ------------------------------------------------------------------ BASELINE cfg4 in full (8 GPUs)
def test_cfg4_500m_rows_over_8_gpus_vs_cpu_scan_of_every_row(vg, orc, gpu):
    if gpu < 8:
        pytest.skip("needs 8 GPUs: vec0 bit[1024], 500 M vectors sharded over 8 B200 (BASELINE.json configs[3])")
    n, dims, k, nq = int(os.environ.get("VECGPU_CFG4_ROWS", 500_000_000)), 1024, 10, 8
    q = orc.synth_rows(BIT, 77, 1, nq, dims, 0)
    er, ed, ec = orc.knn_synth(BIT, dims, 5, 1, n, 0, q, k, HAMMING)
    with vg.ShardedSlab(BIT, dims, devices=list(range(8)), capacity_hint=n, max_queries=64, max_k=16) as g:
        g.fill_synthetic(5, n, kind=0)
        assert g.count() == (n, n)
        for qi in range(3):
            r, d, c = g.knn(q[qi], k, HAMMING)
            assert np.array_equal(r[0], er[qi]) and same_bits(d[0], ed[qi])
        r, d, c = g.knn(q, k, HAMMING)
        assert np.array_equal(r, er) and same_bits(d, ed)
