"""Throughput of the batched (1024-query) paths on the 10M x 768 corpus: tensor cores vs CUDA-core exact."""
import os, signal, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sqlite_vec_hnsw_b200 as vg
signal.alarm(280)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
dims = 768
s = vg.Slab(0, dims); s.fill_synthetic(seed=3, n=n, kind=1)
for metric, name in ((2, "cos"), (0, "l2")):
    for nq in (128, 1024):
        q = torch.randn(nq, dims, device="cuda")
        for tc in ("1", "0"):
            if tc == "0" and nq > 128: continue
            os.environ["VECGPU_TC"] = tc
            r, d = s.knn_device(q, 10, metric); torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            iters = 3 if tc == "1" else 1
            a.record()
            for _ in range(iters): r, d = s.knn_device(q, 10, metric)
            b.record(); torch.cuda.synchronize()
            ms = a.elapsed_time(b) / iters
            flops = 2.0 * nq * n * dims
            print(f"{name} nq={nq:5d} tc={tc}: {ms:9.2f} ms  {nq / ms * 1e3:9.0f} q/s  {flops / ms / 1e9:8.1f} algorithmic TFLOP/s (x3 executed)  tc_stats={vg.tc_stats()}", flush=True)
