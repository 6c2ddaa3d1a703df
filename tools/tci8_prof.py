"""Profiling driver for tci8_scan_kernel: N x 1024 int8 rows, 1024-query L2 batches at k = K."""
import os, signal, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sqlite_vec_hnsw_b200 as vg
signal.alarm(int(sys.argv[3]) if len(sys.argv) > 3 else 280)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
K = int(sys.argv[2]) if len(sys.argv) > 2 else 100
s = vg.Slab(1, 1024); s.fill_synthetic(seed=4, n=n, kind=0)
q = torch.randint(-128, 127, (1024, 1024), dtype=torch.int8, device="cuda")
for rep in range(3):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); r, d = s.knn_device(q, K, 0); b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b)
    print(f"i8[1024] l2 k={K} n={n} nq=1024: {ms:9.2f} ms  {1024 / ms * 1e3:9.0f} q/s  {2.0 * 1024 * n * 1024 / ms / 1e9:8.1f} TOP/s", flush=True)
