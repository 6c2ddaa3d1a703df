"""Hamming batched scan (QB = 8 passes) timing: bit[1024], 64 queries."""
import os, signal, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sqlite_vec_hnsw_b200 as vg
signal.alarm(200)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20_000_000
s = vg.Slab(2, 1024); s.fill_synthetic(seed=5, n=n, kind=0)
for nq in (1, 8, 64):
    q = torch.randint(0, 255, (nq, 128), dtype=torch.uint8, device="cuda")
    s.knn_device(q, 10, 3); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(3): s.knn_device(q, 10, 3)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 3
    print(f"bit[1024] hamming n={n} nq={nq}: {ms:8.3f} ms  {nq / ms * 1e3:9.0f} q/s  {n * 128 * ((nq + 7) // 8) / ms / 1e6:7.0f} GB/s streamed  {n * 32.0 * nq / ms / 1e9:7.2f} Tpopc/s", flush=True)
