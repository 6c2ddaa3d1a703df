#!/usr/bin/env python
"""Where does a single query's end-to-end time go?  For a few table sizes: device back-to-back time per query
(vecgpu_knn_device, no sync between queries) vs end-to-end time per query (vecgpu_knn: host query in, host top-k out),
with the fused scan tail on and off.   python tools/e2e_probe.py"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sqlite_vec_hnsw_b200 as vg  # noqa: E402

vg.load_library()
dev = torch.device("cuda", 0)
stream = torch.cuda.current_stream(dev)
for name, elem, dims, metric, k, n in [("cfg1 10k x f32[384] L2", 0, 384, 0, 10, 10_000), ("shard 1.25M x f32[768] cos", 0, 768, 2, 10, 1_250_000),
                                       ("cfg2 10M x f32[768] cos", 0, 768, 2, 10, 10_000_000), ("cfg4 62.5M x bit[1024]", 2, 1024, 3, 10, 62_500_000)]:
    sl = vg.Slab(elem, dims)
    sl.fill_synthetic(seed=3, n=n, kind=1 if elem == 0 else 0)
    rb = sl.row_bytes
    qh = np.random.default_rng(1).integers(0, 255, size=(64, rb), dtype=np.uint8)
    if elem == 0:
        qh = np.random.default_rng(1).standard_normal((64, dims)).astype("<f4").view(np.uint8).reshape(64, rb)
    qd = torch.from_numpy(qh).to(dev)
    for fuse in ("0", "1", "0", "1", "0", "1"):
        os.environ["VECGPU_FUSE_MERGE"] = fuse
        iters = 200 if n <= 2_000_000 else 30
        for _ in range(5):
            sl.knn_device(qd[0], k, metric, stream=stream.cuda_stream)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for i in range(iters):
            sl.knn_device(qd[i % 64], k, metric, stream=stream.cuda_stream)
        b.record(stream)
        torch.cuda.synchronize()
        t_dev = a.elapsed_time(b) / iters * 1e3
        for _ in range(5):
            sl.knn(qh[0], k, metric)
        t0 = time.perf_counter()
        for i in range(iters):
            sl.knn(qh[i % 64], k, metric)
        t_e2e = (time.perf_counter() - t0) / iters * 1e6
        gbs = n * rb / (t_dev * 1e-6) / 1e9
        print(f"{name:32s} fuse={fuse}: device {t_dev:9.1f} us/query ({gbs:7.1f} GB/s)   e2e {t_e2e:9.1f} us/query   overhead {t_e2e - t_dev:6.1f} us", flush=True)
    sl.close()
    torch.cuda.empty_cache()
