import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle, sqlite_vec_hnsw_b200 as vg
n, dims, M, efc = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
v = oracle.synth_rows(0, 6, 1, n, dims, 1); q = oracle.synth_rows(0, 7, 1, 12, dims, 1)
s = vg.Slab(0, dims); s.load(v)
idx = vg.HnswIndex(s, 0, M=M, ef_construction=efc, seed=1); idx.rebuild()
print("built", idx.stats(), flush=True)
os.environ["VECGPU_HNSW_CTA_MAX_NQ"] = "0"
wr, wd, wc = idx.search(q, 10, ef_search=64); print("warp ok", flush=True)
os.environ["VECGPU_HNSW_CTA_MAX_NQ"] = "64"
for nq in (1, 2, 12):
    t0 = time.time(); cr, cd, cc = idx.search(q[:nq], 10, ef_search=64); print("cta nq", nq, time.time() - t0, np.array_equal(cr, wr[:nq]), flush=True)
