"""Timing experiments on tci8_scan_kernel via VECGPU_TCI_DEBUG (results are invalid in debug modes)."""
import os, signal, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sqlite_vec_hnsw_b200 as vg
signal.alarm(250)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
s = vg.Slab(1, 1024); s.fill_synthetic(seed=4, n=n, kind=0)
q = torch.randint(-128, 127, (1024, 1024), dtype=torch.int8, device="cuda")
for K in (10, 100):
    for dbg in (0, 1, 2, 6):
        os.environ["VECGPU_TCI_DEBUG"] = str(dbg)
        best = 1e9
        for rep in range(3):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); s.knn_device(q, K, 0); b.record(); torch.cuda.synchronize()
            best = min(best, a.elapsed_time(b))
        print(f"k={K} debug={dbg}: {best:8.2f} ms  {2.0 * 1024 * n * 1024 / best / 1e9:8.1f} TOP/s", flush=True)
