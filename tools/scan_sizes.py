"""Single-query exact scan latency across table sizes (device-resident query, CUDA events)."""
import os, signal, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sqlite_vec_hnsw_b200 as vg
signal.alarm(200)
for dims in (384, 768):
    for n in (10_000, 100_000, 1_000_000, 4_000_000):
        s = vg.Slab(0, dims); s.fill_synthetic(seed=3, n=n, kind=1)
        q = torch.randn(dims, device="cuda")
        s.knn_device(q, 10, 2); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(50): s.knn_device(q, 10, 2)
        b.record(); torch.cuda.synchronize()
        us = a.elapsed_time(b) / 50 * 1e3
        print(f"f32[{dims}] cosine k=10 n={n:8d}: {us:8.1f} us/query  {n * dims * 4 / us / 1e3:7.1f} GB/s  (ideal at 6.9 TB/s: {n * dims * 4 / 6.9e6:7.1f} us)", flush=True)
        s.close()
