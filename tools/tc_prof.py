"""Profiling driver for tc_scan_kernel: 10M x 768 f32 cosine, 1024-query batches."""
import os, signal, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sqlite_vec_hnsw_b200 as vg
signal.alarm(int(sys.argv[2]) if len(sys.argv) > 2 else 280)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
s = vg.Slab(0, 768); s.fill_synthetic(seed=3, n=n, kind=1)
q = torch.randn(1024, 768, device="cuda")
for rep in range(3):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); s.knn_device(q, 10, 2); b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b)
    print(f"cos nq=1024 n={n}: {ms:8.2f} ms  {1024 / ms * 1e3:8.0f} q/s  {2.0 * 1024 * n * 768 / ms / 1e9:8.1f} algorithmic TFLOP/s  tc_stats={vg.tc_stats()}", flush=True)
