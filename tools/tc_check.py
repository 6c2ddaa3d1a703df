"""Tensor-core batched path vs the CPU oracle (rowids + distances must be bit-exact).  Guarded by alarms."""
import os, signal, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle
import sqlite_vec_hnsw_b200 as vg

signal.alarm(200)
ok = True
for name, n, dims, nq, k, metric, kind in [
    ("L2 small", 20000, 96, 64, 10, 0, 0), ("cos small", 20000, 96, 64, 10, 2, 1), ("cos 768", 60000, 768, 130, 10, 2, 1),
    ("L2 384 ragged nq", 50001, 384, 77, 20, 0, 0), ("cos dims=100", 30000, 100, 33, 5, 2, 1), ("L2 k=96", 40000, 64, 40, 96, 0, 0),
]:
    s = vg.Slab(0, dims)
    s.fill_synthetic(seed=11, n=n, kind=kind)
    cpu = oracle.synth_rows(0, 11, 1, n, dims, kind)
    q = oracle.synth_rows(0, 12, 1, nq, dims, kind)
    os.environ["VECGPU_TC"] = "1"
    t0 = time.time(); r, d, c = s.knn(q, k, metric); t1 = time.time()
    os.environ["VECGPU_TC"] = "0"
    r0, d0, c0 = s.knn(q, k, metric)
    er, ed, ec = oracle.knn(0, dims, cpu, q, k, metric)
    same = np.array_equal(r, er) and np.array_equal(d.view("<u4"), ed.view("<u4")) and np.array_equal(c, ec)
    same0 = np.array_equal(r0, er) and np.array_equal(d0.view("<u4"), ed.view("<u4"))
    nbad = int((r != er).any(axis=1).sum())
    print(f"{name:18s} n={n} D={dims} nq={nq} k={k}: tc {'OK' if same else 'MISMATCH (%d queries)' % nbad}  exact-path {'OK' if same0 else 'MISMATCH'}  ({(t1 - t0) * 1e3:.1f} ms first call)", flush=True)
    ok &= same and same0
    s.close()
print("TC_CHECK", "PASS" if ok else "FAIL")
