import os, signal, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sqlite_vec_hnsw_b200 as vg
signal.alarm(500)
rng = np.random.default_rng(0)
n, dims, nq = 6000, 64, 300
centers = rng.standard_normal((60, dims)).astype("<f4")
v = (centers[rng.integers(0, 60, n)] + 0.35 * rng.standard_normal((n, dims))).astype("<f4")
q = (centers[rng.integers(0, 60, nq)] + 0.35 * rng.standard_normal((nq, dims))).astype("<f4")
s = vg.Slab(0, dims); s.load(v)
er, _, _ = s.knn(q, 10, 0)
for M, efc in ((16, 200), (32, 400)):
    for b in (1, 64, 1024):
        idx = vg.HnswIndex(s, vg.DistanceMetric.L2, M=M, ef_construction=efc, seed=1)
        t0 = time.time(); idx.rebuild(batch=b); t1 = time.time()
        out = []
        for ef in (50, 200, 800):
            r, d, c = idx.search(q, 10, ef_search=ef)
            out.append(sum(len(set(a.tolist()) & set(bb.tolist())) for a, bb in zip(r, er)) / er.size)
        print(f"clustered n={n} D={dims} M={M} efc={efc} batch={b:5d}: build {t1-t0:5.1f}s recall ef50/200/800 = " + " ".join(f"{x:.3f}" for x in out), flush=True)
        idx.close()
