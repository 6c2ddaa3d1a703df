#!/usr/bin/env python
"""Single-query latency of small tables against the number of CTAs of the scan (VECGPU_SCAN_GX caps it): fewer CTAs leave fewer
partial lists for the fused final merge but stream more rows each.   python tools/small_table_gx.py"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle
import sqlite_vec_hnsw_b200 as vg
CASES = ((10_000, 384, 0, 10), (50_000, 384, 0, 10), (10_000, 384, 0, 100), (10_000, 384, 0, 32), (10_000, 384, 0, 64), (100_000, 384, 0, 100),
         (1_000_000, 384, 0, 100), (2_000, 128, 0, 50))
NQ = int(sys.argv[2]) if len(sys.argv) > 2 else 1   # queries per call (a small batch shares one pass of the scan)
if len(sys.argv) > 1 and sys.argv[1] == "long":   # result lists beyond the fused merge's reach at one CTA per SM (148 k > 16384)
    CASES = ((10_000, 384, 0, 200), (10_000, 384, 0, 500), (10_000, 384, 0, 1000), (100_000, 384, 0, 200), (100_000, 384, 0, 1000), (1_000_000, 384, 0, 500))
for n, dims, metric, k in CASES:
    with vg.Slab(0, dims) as s:
        s.fill_synthetic(seed=1, n=n, kind=1)
        q = oracle.synth_rows(0, 2, 1, NQ, dims, 1)
        line = []
        ref = None
        for cap in (0, 8, 16, 24, 32, 48, 64, 96, 148):
            if cap: os.environ["VECGPU_SCAN_GX"] = str(cap)
            else: os.environ.pop("VECGPU_SCAN_GX", None)
            for _ in range(200): r = s.knn(q, k, metric)
            t0 = time.perf_counter()
            for _ in range(2000): r = s.knn(q, k, metric)
            us = (time.perf_counter() - t0) / 2000 * 1e6
            if ref is None: ref = r
            assert all(np.array_equal(a.view("u1"), b.view("u1")) for a, b in zip(r, ref))
            line.append(f"{cap or 'auto'}: {us:.1f}")
        print(f"{n} x f32[{dims}] metric {metric} k={k} nq={NQ}  us per call by CTA cap  " + "  ".join(line), flush=True)
os.environ.pop("VECGPU_SCAN_GX", None)
