#!/usr/bin/env python
"""Throughput of vecgpu_score (K5, the candidate scoring of search_layer): one call scoring NQ queries x CAND random candidates each
over N x f32[384] rows (host rowids in, host distances out) and the device time of its pair_kernel.   python tools/score_rate.py"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle  # noqa: E402
import sqlite_vec_hnsw_b200 as vg  # noqa: E402

n, dims = 1_000_000, 384
s = vg.Slab(0, dims)
s.fill_synthetic(seed=6, n=n, kind=1)
rng = np.random.default_rng(1)
for nq, cand in ((1, 32), (1, 64), (256, 32), (4096, 32), (16384, 64)):
    q = oracle.synth_rows(0, 67, 1, nq, dims, 1)
    ids = rng.integers(1, n + 1, size=(nq, cand)).astype("<i8").reshape(-1)
    offs = (np.arange(nq + 1, dtype="<u4") * cand).astype("<u4")
    s.score(q, ids, offs, 0)
    reps = 20 if nq <= 256 else 5
    t0 = time.perf_counter()
    for _ in range(reps):
        s.score(q, ids, offs, 0)
    el = (time.perf_counter() - t0) / reps
    print(f"nq={nq:6d} x {cand} candidates: {el * 1e3:8.3f} ms per call, {nq * cand / el / 1e6:8.1f} M pairs/s, {nq * cand * dims * 4 / el / 1e9:8.1f} GB/s gathered", flush=True)
s.close()
