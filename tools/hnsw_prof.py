"""Profiling driver for the on-device HNSW search kernel: build N x 384 (device path), then ONE search launch of NQ queries."""
import os, signal, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sqlite_vec_hnsw_b200 as vg
import oracle
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
nq = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
signal.alarm(int(sys.argv[3]) if len(sys.argv) > 3 else 600)
dims, k = 384, 10
s = vg.Slab(0, dims); s.fill_synthetic(seed=6, n=n, kind=1)
idx = vg.HnswIndex(s, vg.DistanceMetric.L2, M=16, ef_construction=200, seed=1)
t0 = time.time(); idx.rebuild(); print(f"build {time.time() - t0:.2f} s launches={idx.device_stats()['launches']}", flush=True)
q = oracle.synth_rows(0, 67, 1, nq, dims, 1)
for rep in range(2):
    t0 = time.time(); idx.search(q, k, ef_search=200); dt = time.time() - t0
    print(f"search {nq} queries: {dt * 1e3:.1f} ms  {nq / dt:.0f} q/s", flush=True)
