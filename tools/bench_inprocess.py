#!/usr/bin/env python
"""ONE-process sharded slab (vecgpu_sharded_*: the form the Rust extension links) over every visible GPU: end-to-end
single-query throughput through ONE C call per query (host query in, host top-k out), for BASELINE cfg2 and cfg4.
    python tools/bench_inprocess.py [n_gpus]"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle  # noqa: E402  (query generation only)
import sqlite_vec_hnsw_b200 as vg  # noqa: E402

lib = vg.load_library()
ndev = int(sys.argv[1]) if len(sys.argv) > 1 else lib.vecgpu_device_count()
for name, elem, dims, metric, k, n, seed, kind in [("cfg2 10M x f32[768] cosine k=10", 0, 768, 2, 10, 10_000_000, 3, 1),
                                                    ("cfg4 500M x bit[1024] hamming k=10", 2, 1024, 3, 10, 500_000_000 if ndev >= 8 else 62_500_000 * ndev, 5, 0)]:
    g = vg.ShardedSlab(elem, dims, devices=list(range(ndev)), capacity_hint=n, max_queries=64, max_k=16)
    g.fill_synthetic(seed, n, kind=kind)
    q = oracle.synth_rows(elem, 33, 1, 64, dims, kind)
    for i in range(8):
        g.knn(q[i], k, metric)
    iters = 200
    t0 = time.perf_counter()
    for i in range(iters):
        r, d, c = g.knn(q[i % 64], k, metric)
    el = time.perf_counter() - t0
    bytes_total = n * g.row_bytes
    print(json.dumps({"workload": name, "api": "vecgpu_sharded_knn (one process, one call per query, host buffers)", "gpus": ndev,
                      "queries_per_s": iters / el, "us_per_query": el / iters * 1e6, "aggregate_gbs": bytes_total * iters / el / 1e9}), flush=True)
    g.close()
