"""HNSW recall on clustered vs structureless data, same parameters (M=16, efc=200), batched build."""
import os, signal, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sqlite_vec_hnsw_b200 as vg
signal.alarm(500)
rng = np.random.default_rng(0)
n, dims, nq = 200_000, 384, 500
centers = rng.standard_normal((2000, dims)).astype("<f4")
for name, data_fn in [
    ("clustered (2000 centres + 0.35*noise)", lambda m: (centers[rng.integers(0, 2000, m)] + 0.35 * rng.standard_normal((m, dims))).astype("<f4")),
    ("iid N(0,1)", lambda m: rng.standard_normal((m, dims)).astype("<f4")),
]:
    v = data_fn(n); q = data_fn(nq)
    s = vg.Slab(0, dims); s.load(v)
    er, ed, ec = s.knn(q, 10, 0)
    idx = vg.HnswIndex(s, vg.DistanceMetric.L2, M=16, ef_construction=200, seed=1)
    t0 = time.time(); idx.rebuild(batch=4096); t1 = time.time()
    out = []
    for ef in (50, 200, 800, 3200):
        r, d, c = idx.search(q, 10, ef_search=ef)
        out.append(sum(len(set(a.tolist()) & set(b.tolist())) for a, b in zip(r, er)) / er.size)
    print(f"{name:40s} n={n} D={dims}: build {t1 - t0:5.1f} s; recall@10 at ef=50/200/800/3200: " + " ".join(f"{x:.3f}" for x in out), flush=True)
    idx.close(); s.close()
