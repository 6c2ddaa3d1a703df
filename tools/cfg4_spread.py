#!/usr/bin/env python
"""Why cfg4's single-query time (62.5 M x bit[1024], 8 GB, ~1.2 ms) moves between runs: per-launch device times of repeated scans,
for slabs allocated at different moments of the process.   python tools/cfg4_spread.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle  # noqa: E402
import sqlite_vec_hnsw_b200 as vg  # noqa: E402

dev = torch.device("cuda:0")
n, dims = 62_500_000, 1024
q = torch.from_numpy(oracle.synth_rows(2, 77, 1, 4, dims, 0).copy()).to(dev)


def measure(sl, label):
    stream = torch.cuda.current_stream()
    for _ in range(3):
        sl.knn_device(q[0], 10, 3, stream=stream.cuda_stream)
    torch.cuda.synchronize()
    for rep in range(3):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(31)]
        ev[0].record()
        for i in range(30):
            sl.knn_device(q[0], 10, 3, stream=stream.cuda_stream)
            ev[i + 1].record()
        torch.cuda.synchronize()
        t = np.array([ev[i].elapsed_time(ev[i + 1]) for i in range(30)])
        print(f"{label} set {rep}: mean {t.mean():.3f} ms  min {t.min():.3f}  max {t.max():.3f}  -> {n * 128 / t.mean() / 1e6:.0f} GB/s (best launch {n * 128 / t.min() / 1e6:.0f})", flush=True)


for trial in range(3):
    ballast = None
    if trial == 1:
        ballast = torch.empty(60 * 1024**3, dtype=torch.uint8, device=dev)  # the slab lands elsewhere in HBM
    if trial == 2:
        ballast = torch.empty(120 * 1024**3, dtype=torch.uint8, device=dev)
    sl = vg.Slab(2, dims)
    sl.fill_synthetic(seed=6, n=n, kind=0)
    measure(sl, f"trial {trial} (ballast {0 if ballast is None else ballast.numel() >> 30} GiB)")
    sl.close()
    del ballast
    torch.cuda.empty_cache()

# ... and right behind the int8 tensor-core batches of cfg3, as in bench.py
import subprocess  # noqa: E402
s8 = vg.Slab(1, 1024)
s8.fill_synthetic(seed=5, n=50_000_000, kind=0)
qb8 = torch.from_numpy(oracle.synth_rows(1, 78, 1, 1024, 1024, 0).copy()).to(dev)
sl = vg.Slab(2, dims)
sl.fill_synthetic(seed=6, n=n, kind=0)
measure(sl, "before the int8 batches")
for _ in range(4):
    s8.knn_device(qb8, 100, 0, stream=torch.cuda.current_stream().cuda_stream)
torch.cuda.synchronize()
print("clocks right after the batches:", subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,clocks_throttle_reasons.sw_power_cap", "--format=csv,noheader"],
                                                       capture_output=True, text=True).stdout.strip(), flush=True)
measure(sl, "right after 4 int8 1024-query batches")
sl.close()
s8.close()

