"""Per-launch time of knn_device for several k (isolates top-k / merge cost from streaming)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sqlite_vec_hnsw_b200 as vg
from tools.sweep_scan import time_knn
for name, elem, dims, metric, n, kind in [("i8[1024] l2", 1, 1024, 0, 12_000_000, 0), ("f32[768] cos", 0, 768, 2, 4_000_000, 1), ("bit[1024]", 2, 1024, 3, 64_000_000, 0)]:
    slab = vg.Slab(elem, dims); slab.fill_synthetic(seed=7, n=n, kind=kind)
    q = torch.randn(dims, device="cuda") if elem == 0 else torch.randint(0, 255, (slab.row_bytes,), dtype=torch.uint8, device="cuda")
    for k in (1, 10, 32, 100, 300, 1000):
        ms = time_knn(slab, q, k, metric, iters=20)
        print(f"{name:14s} k={k:5d}: {ms:7.3f} ms  {n * slab.row_bytes / ms / 1e6:7.1f} GB/s", flush=True)
    slab.close()
