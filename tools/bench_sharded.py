"""BASELINE.json configs 2 (batched), 3 and 4 sharded by rowid range over the ranks of a torchrun launch.
   python -m torch.distributed.run --nproc-per-node N tools/bench_sharded.py
Prints one JSON line per config on rank 0 (device-event timing, max over ranks)."""
import json, os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sqlite_vec_hnsw_b200 as vg
from sqlite_vec_hnsw_b200 import dist as vdist
import oracle

world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); lr = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
scale = float(os.environ.get("VECGPU_SHARD_SCALE", "1.0"))
CFG = [  # name, elem, dims, metric, k, total rows, synth kind, seed, batch sizes
    ("cfg2_f32_768_cos_k10", 0, 768, 2, 10, int(10_000_000 * scale), 1, 3, (1, 1024)),
    ("cfg3_i8_1024_l2_k100", 1, 1024, 0, 100, int(50_000_000 * scale), 0, 4, (1, 1024)),
    ("cfg4_bit_1024_hamming_k10", 2, 1024, 3, 10, int(500_000_000 * scale), 0, 5, (1, 64)),
]
stream = torch.cuda.current_stream(dev)
for name, elem, dims, metric, k, n, kind, seed, batches in CFG:
    if os.environ.get("ONLY") and os.environ["ONLY"] not in name: continue
    sh = vdist.ShardedSlab(vg, elem, dims, n, rank, world, lr)
    sh.fill_synthetic(seed, kind)
    out = {"config": name, "n_gpus": world, "rows_total": n, "rows_per_gpu": sh.hi - sh.lo}
    for nq in batches:
        q = torch.from_numpy(oracle.synth_rows(elem, 99, 1, nq, dims, kind).copy()).to(dev)
        reps = 16 if nq == 1 else 3
        def step():
            res = None
            for _ in range(reps):
                res = sh.knn_device(q, k, metric)
            return res
        step(); torch.cuda.synchronize()
        if world > 1: dist.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream); r, d = step(); b.record(stream); torch.cuda.synchronize()
        t = torch.tensor([a.elapsed_time(b)], device=dev, dtype=torch.float64)
        if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item()) / reps
        key = "single_query" if nq == 1 else f"batch{nq}"
        out[key] = {"ms": ms, "queries_per_s": nq / ms * 1e3, "aggregate_scan_gbs": n * sh.slab.row_bytes / ms / 1e6 if nq == 1 else None,
                    "top1": [int(r[0, 0]), float(d[0, 0])]}
    if rank == 0: print(json.dumps(out), flush=True)
    sh.close(); torch.cuda.empty_cache()
if world > 1: dist.destroy_process_group()
