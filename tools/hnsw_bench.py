"""Config 5: vec_rebuild_hnsw on N x 384 f32 L2, M=16, ef_construction=200 — build time, distances/s, recall@10,
for the on-device search kernel (K6) and the lockstep driver (VECGPU_HNSW_DEVICE=0)."""
import os, signal, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sqlite_vec_hnsw_b200 as vg
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
signal.alarm(int(sys.argv[3]) if len(sys.argv) > 3 else 900)
modes = sys.argv[4].split(",") if len(sys.argv) > 4 else ["1", "0"]
dims, nq, k = 384, 1000, 10
s = vg.Slab(0, dims); s.fill_synthetic(seed=6, n=n, kind=1)
import oracle
q = oracle.synth_rows(0, 66, 1, nq, dims, 1)
qbig = oracle.synth_rows(0, 67, 1, 20000, dims, 1)
er, ed, ec = s.knn(q, k, 0)  # exact ground truth from the scan (K1/K2)
edges = {}
for mode in modes:
    os.environ["VECGPU_HNSW_DEVICE"] = mode
    name = "device-search" if mode == "1" else "lockstep"
    idx = vg.HnswIndex(s, vg.DistanceMetric.L2, M=16, ef_construction=200, seed=1)
    t0 = time.time(); idx.rebuild(batch=batch); t1 = time.time()
    st = idx.stats()
    print(f"[{name}] build: n={n} batch={batch}: {t1 - t0:.1f} s  {n / (t1 - t0):.0f} vec/s  distances={st['distances_scored']:.3e} "
          f"({st['distances_scored'] / (t1 - t0) / 1e6:.1f} M/s, {st['distances_scored'] * dims * 4 / (t1 - t0) / 1e9:.1f} GB/s gathered)  "
          f"launches/rounds={st['rounds']} edges={st['edges']} entry_level={st['entry_level']} dev={idx.device_stats()}", flush=True)
    for qs, label in ((q, "1000"), (q[:1], "1"), (qbig, "20000")):
        if mode == "0" and len(qs) > 5000:
            continue
        idx.search(qs, k, ef_search=200)  # warm-up of the same shape (workspaces are sized by the launch)
        sc0 = idx.stats()["distances_scored"]
        t0 = time.time(); r, d, c = idx.search(qs, k, ef_search=200); t1 = time.time()
        sc = idx.stats()["distances_scored"] - sc0
        extra = ""
        if label == "1000":
            hit = sum(len(set(a.tolist()) & set(b.tolist())) for a, b in zip(r, er))
            extra = f"  recall@10 = {hit / er.size:.4f}"
        print(f"[{name}] search: {len(qs)} queries ef=200 in {(t1 - t0) * 1e3:.2f} ms ({len(qs) / (t1 - t0):.0f} q/s, "
              f"{sc / max(1, len(qs)):.0f} distances/query, {sc * dims * 4 / (t1 - t0) / 1e9:.1f} GB/s gathered){extra}", flush=True)
    edges[mode] = idx.export_edges() if n <= 200000 else None
    res = idx.search(q, k, ef_search=200)
    edges["r" + mode] = res
    idx.close()
if len(modes) == 2:
    if edges[modes[0]] is not None:
        print("graphs identical:", all(np.array_equal(a.view("u1"), b.view("u1")) for a, b in zip(edges[modes[0]], edges[modes[1]])))
    print("search results identical:", all(np.array_equal(a.view("u1"), b.view("u1")) for a, b in zip(edges["r" + modes[0]], edges["r" + modes[1]])))
