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

# int8 L2 batches (config 3 shape): tensor cores (exact) vs CUDA-core dp4a batches
s.close()
n8 = int(sys.argv[2]) if len(sys.argv) > 2 else 20_000_000
s = vg.Slab(1, 1024); s.fill_synthetic(seed=4, n=n8, kind=0)
for nq in (128, 1024):
    q = torch.randint(-128, 127, (nq, 1024), dtype=torch.int8, device="cuda")
    for tc in ("1", "0"):
        if tc == "0" and nq > 128: continue
        os.environ["VECGPU_TC"] = tc
        K8 = int(sys.argv[3]) if len(sys.argv) > 3 else 100
        r, d = s.knn_device(q, K8, 0); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); r, d = s.knn_device(q, K8, 0); b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b)
        print(f"i8[1024] l2 k={K8} n={n8} nq={nq:5d} tc={tc}: {ms:9.2f} ms  {nq / ms * 1e3:9.0f} q/s  {2.0 * nq * n8 * 1024 / ms / 1e9:8.1f} TOP/s", flush=True)
