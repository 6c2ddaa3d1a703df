#!/usr/bin/env python
"""Where the one-CTA-per-query HNSW walk stops paying: calls of nq queries through vecgpu_hnsw_search on cfg5's graph
(1 M x f32[384], M=16, efc=200), both kernels.   python tools/hnsw_crossover.py [rows] [ef]"""
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
q = oracle.synth_rows(0, 67, 1, 4096, dims, 1)
for nq in (64, 128, 148, 256, 296, 444, 592, 740, 888, 1036, 1184, 1480, 2048, 4096):
    line = f"nq={nq:5d}:"
    for mode, label in (("100000", "CTA"), ("0", "warp")):
        os.environ["VECGPU_HNSW_CTA_MAX_NQ"] = mode
        idx.search(q[:nq], k, ef_search=ef)
        reps = 5
        t0 = time.perf_counter()
        for _ in range(reps):
            idx.search(q[:nq], k, ef_search=ef)
        el = (time.perf_counter() - t0) / reps
        line += f"  {label} {el * 1e3:7.3f} ms"
    print(line, flush=True)
idx.close()
s.close()
