"""Recall vs insert batch size (1 = sequential like the reference)."""
import os, signal, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sqlite_vec_hnsw_b200 as vg
import oracle
signal.alarm(500)
for n, dims, M, efc, batches in [(4000, 64, 16, 200, (1, 64, 1024)), (20000, 64, 16, 200, (64, 1024, 4096)), (30000, 384, 16, 200, (64, 4096)), (30000, 384, 32, 400, (4096,))]:
    s = vg.Slab(0, dims); s.fill_synthetic(seed=6, n=n, kind=1)
    q = oracle.synth_rows(0, 66, 1, 500, dims, 1)
    er, ed, ec = s.knn(q, 10, 0)
    for b in batches:
        idx = vg.HnswIndex(s, vg.DistanceMetric.L2, M=M, ef_construction=efc, seed=1)
        t0 = time.time(); idx.rebuild(batch=b); t1 = time.time()
        out = []
        for ef in (50, 200, 800):
            r, d, c = idx.search(q, 10, ef_search=ef)
            out.append(sum(len(set(a.tolist()) & set(bb.tolist())) for a, bb in zip(r, er)) / er.size)
        st = idx.stats()
        print(f"n={n} D={dims} M={M} efc={efc} batch={b:5d}: build {t1 - t0:6.1f} s  recall@10 ef50/200/800 = {out[0]:.3f} {out[1]:.3f} {out[2]:.3f}  avg deg0={st['edges'] / n:.1f}", flush=True)
        idx.close()
    s.close()
