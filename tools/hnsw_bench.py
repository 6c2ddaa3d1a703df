"""Config 5: vec_rebuild_hnsw on N x 384 f32 L2, M=16, ef_construction=200 — build time, distances/s, recall@10."""
import os, signal, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sqlite_vec_hnsw_b200 as vg
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
signal.alarm(int(sys.argv[3]) if len(sys.argv) > 3 else 900)
dims, nq, k = 384, 1000, 10
s = vg.Slab(0, dims); s.fill_synthetic(seed=6, n=n, kind=1)
idx = vg.HnswIndex(s, vg.DistanceMetric.L2, M=16, ef_construction=200, seed=1)
t0 = time.time(); idx.rebuild(batch=batch); t1 = time.time()
st = idx.stats()
print(f"build: n={n} batch={batch}: {t1 - t0:.1f} s  {n / (t1 - t0):.0f} vec/s  distances={st['distances_scored']:.3e} "
      f"({st['distances_scored'] / (t1 - t0) / 1e6:.1f} M/s, {st['distances_scored'] * dims * 4 / (t1 - t0) / 1e9:.1f} GB/s gathered)  rounds={st['rounds']} edges={st['edges']} entry_level={st['entry_level']}", flush=True)
import oracle
q = oracle.synth_rows(0, 66, 1, nq, dims, 1)
t0 = time.time(); r, d, c = idx.search(q, k, ef_search=200); t1 = time.time()
er, ed, ec = s.knn(q, k, 0)  # exact ground truth from the scan (K1/K2)
hit = sum(len(set(a.tolist()) & set(b.tolist())) for a, b in zip(r, er))
print(f"search: {nq} queries ef=200 in {t1 - t0:.2f} s ({nq / (t1 - t0):.0f} q/s)  recall@10 = {hit / er.size:.4f}", flush=True)
