"""Batched f32 / int8 queries on medium tables: where does the time go besides the tensor-core kernel?"""
import os, signal, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sqlite_vec_hnsw_b200 as vg
signal.alarm(250)
for n in (1_250_000, 100_000):
    s = vg.Slab(0, 768); s.fill_synthetic(seed=3, n=n, kind=1)
    q = torch.randn(1024, 768, device="cuda")
    for nq in (1024, 128):
        best = 1e9
        for rep in range(5):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(); t0 = time.perf_counter()
            a.record(); s.knn_device(q[:nq], 10, 2); b.record(); torch.cuda.synchronize()
            wall = (time.perf_counter() - t0) * 1e3
            best = min(best, a.elapsed_time(b))
        print(f"f32 cos n={n} nq={nq}: {best:7.3f} ms device ({wall:7.3f} ms wall last)  {nq / best * 1e3:9.0f} q/s", flush=True)
    s.close()
