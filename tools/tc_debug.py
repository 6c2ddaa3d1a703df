"""Timing experiments on tc_scan_kernel (f32 batches) via VECGPU_TCI_DEBUG / VECGPU_TC_TERMS (debug results are invalid)."""
import os, signal, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sqlite_vec_hnsw_b200 as vg
signal.alarm(250)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
s = vg.Slab(0, 768); s.fill_synthetic(seed=3, n=n, kind=1)
q = torch.randn(1024, 768, device="cuda")
for terms in ("1", "3"):
    os.environ["VECGPU_TC_TERMS"] = terms
    for dbg in (0, 1):
        os.environ["VECGPU_TCI_DEBUG"] = str(dbg)
        best = 1e9
        for rep in range(3):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); s.knn_device(q, 10, 2); b.record(); torch.cuda.synchronize()
            best = min(best, a.elapsed_time(b))
        print(f"terms={terms} debug={dbg}: {best:8.2f} ms  {2.0 * 1024 * n * 768 / best / 1e9:8.1f} algorithmic TFLOP/s", flush=True)
