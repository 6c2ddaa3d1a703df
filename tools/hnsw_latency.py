#!/usr/bin/env python
"""Single-query HNSW latency on BASELINE cfg5's graph (1 M x f32[384], L2, M=16, efc=200, ef_search=200): one CTA per query
(hnsw_search_cta_kernel) against one warp per query (hnsw_search_kernel), through vecgpu_hnsw_search (host query in, host
top-k out).   python tools/hnsw_latency.py [rows]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle  # noqa: E402
import sqlite_vec_hnsw_b200 as vg  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
dims, k = 384, 10
s = vg.Slab(0, dims)
s.fill_synthetic(seed=6, n=n, kind=1)
idx = vg.HnswIndex(s, 0, M=16, ef_construction=200, seed=1)
t0 = time.time()
idx.rebuild()
print(f"rebuild {n} x {dims}: {time.time() - t0:.2f} s", flush=True)
q = oracle.synth_rows(0, 67, 1, 256, dims, 1)
for ef in (50, 200, 500):
    for mode, label in (("64", "CTA per query"), ("0", "warp per query")):
        os.environ["VECGPU_HNSW_CTA_MAX_NQ"] = mode
        for nq in (1, 16, 64):
            idx.search(q[:nq], k, ef_search=ef)
            sc0 = idx.stats()["distances_scored"]
            reps = 50 if nq == 1 else 10
            t0 = time.perf_counter()
            for i in range(reps):
                idx.search(q[(i * nq) % 192:(i * nq) % 192 + nq], k, ef_search=ef)
            el = (time.perf_counter() - t0) / reps
            sc = (idx.stats()["distances_scored"] - sc0) / (reps * nq)
            print(f"ef={ef:4d} {label:15s} nq={nq:3d}: {el * 1e3:8.3f} ms per call, {el / nq * 1e3:8.3f} ms per query, {sc:7.0f} distances/query", flush=True)
idx.close()
s.close()
