#!/usr/bin/env python
"""Where a single-query HNSW search spends its cycles (hnsw_search_cta_kernel, VECGPU_HNSW_TIMING phase counters of thread 0):
row fetch (incl. waiting for the control warp), scoring, admission into the sorted array, pop + adjacency + visited filter.
   python tools/hnsw_phase_probe.py [rows] [ef]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle  # noqa: E402
import sqlite_vec_hnsw_b200 as vg  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
ef = int(sys.argv[2]) if len(sys.argv) > 2 else 200
dims, k = 384, 10
s = vg.Slab(0, dims)
s.fill_synthetic(seed=6, n=n, kind=1)
idx = vg.HnswIndex(s, 0, M=16, ef_construction=200, seed=1)
idx.rebuild()
q = oracle.synth_rows(0, 67, 1, 64, dims, 1)
for i in range(3):
    idx.search(q[i:i + 1], k, ef_search=ef)
os.environ["VECGPU_HNSW_TIMING"] = "1"
for i in range(4):
    sc0 = idx.stats()["distances_scored"]
    t0 = time.perf_counter()
    idx.search(q[8 + i:9 + i], k, ef_search=ef)
    el = time.perf_counter() - t0
    print(f"query {i}: {el * 1e3:.3f} ms, {idx.stats()['distances_scored'] - sc0} distances, histogram {idx.batch_histogram()}", flush=True)
del os.environ["VECGPU_HNSW_TIMING"]
idx.close()
s.close()
