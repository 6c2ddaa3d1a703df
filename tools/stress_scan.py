"""Repeat-launch stress for the scan pipeline (hang / fault hunting).  Usage: python tools/stress_scan.py [iters]"""
import os
import signal
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sqlite_vec_hnsw_b200 as vg  # noqa: E402

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 300
signal.alarm(240)
for name, elem, dims, metric, k, n, kind in [
    ("f32[768] cos", 0, 768, 2, 10, 1_000_000, 1),
    ("f32[768] l1", 0, 768, 1, 10, 500_000, 1),
    ("i8[1024] l2", 1, 1024, 0, 100, 2_000_000, 0),
    ("bit[1024]", 2, 1024, 3, 10, 8_000_000, 0),
]:
    slab = vg.Slab(elem, dims)
    slab.fill_synthetic(seed=7, n=n, kind=kind)
    q = torch.randn(dims, device="cuda") if elem == 0 else torch.randint(0, 255, (slab.row_bytes,), dtype=torch.uint8, device="cuda")
    for w in ("4", "6", "8"):
        os.environ["VECGPU_SCAN_WARPS"] = w
        r0 = None
        for i in range(iters):
            r, d = slab.knn_device(q, k, metric)
            if i % 50 == 0:
                torch.cuda.synchronize()
                if r0 is None:
                    r0 = r.clone()
                assert torch.equal(r, r0), "non-deterministic result"
        torch.cuda.synchronize()
        print(f"{name} warps={w}: {iters} launches OK", flush=True)
    slab.close()
print("stress OK")
