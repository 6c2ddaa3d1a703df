"""Differential fuzz on the GPU: for random shapes, every fast path must return exactly what the plain exact scan returns
(tensor-core f32 / int8 batches, lane-per-query Hamming, small-merge fast path, HNSW device walks — one CTA per query and one warp per query — vs lockstep driver)."""
import os, signal, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sqlite_vec_hnsw_b200 as vg
budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
signal.alarm(int(budget) + 120)
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1234)
KNOBS = ("VECGPU_TC", "VECGPU_HAM_BATCH", "VECGPU_MERGE_SMALL", "VECGPU_TC_MIN_WORK", "VECGPU_HNSW_DEVICE", "VECGPU_TC_TERMS", "VECGPU_TCI_SAMPLE",
         "VECGPU_HNSW_CTA_MAX_NQ")
def env(**kw):
    for k in KNOBS: os.environ.pop(k, None)
    for k, v in kw.items(): os.environ[k] = str(v)
def rows(elem, n, dims, ties):
    if elem == 0:
        return (rng.integers(-2, 3, size=(n, dims)) if ties else rng.standard_normal((n, dims))).astype("<f4")
    if elem == 1:
        return (rng.integers(-2, 3, size=(n, dims)) if ties else rng.integers(-128, 128, size=(n, dims))).astype("i1")
    nb = (dims + 7) // 8
    r = rng.integers(0, 256, size=(n, nb)).astype("u1")
    if dims % 8: r[:, -1] &= (1 << (dims % 8)) - 1
    if ties: r[:, : max(0, nb - 1)] = 0
    return r
t0, cases, fails = time.time(), 0, 0
while time.time() - t0 < budget:
    kind = rng.integers(0, 4)
    ties = bool(rng.integers(0, 4) == 0)
    if kind == 0:   # f32 batches
        elem, metric, dims = 0, int(rng.choice([0, 2])), int(rng.integers(16, 300))
        n, nq, k = int(rng.integers(8192, 60000)), int(rng.integers(16, 300)), int(rng.integers(1, 97))
        fast = dict(VECGPU_TC_MIN_WORK=0, VECGPU_TC_TERMS=int(rng.choice([1, 3])))
    elif kind == 1:  # int8 batches
        elem, metric, dims = 1, 0, int(rng.integers(16, 600))
        n, nq, k = int(rng.integers(8192, 60000)), int(rng.integers(16, 1100)), int(rng.integers(1, 193))
        fast = dict(VECGPU_TCI_SAMPLE=int(rng.choice([0, 20, 40])))
    elif kind == 2:  # Hamming batches
        elem, metric, dims = 2, 3, int(rng.integers(8, 1025))
        n, nq, k = int(rng.integers(4096, 80000)), int(rng.integers(16, 200)), int(rng.integers(1, 33))
        fast = dict()
    else:            # HNSW: device walk + device link vs lockstep
        elem, metric, dims = int(rng.choice([0, 1])), 0, int(rng.integers(4, 64))
        n, nq, k = int(rng.integers(500, 6000)), 40, int(rng.integers(1, 20))
        M, efc, ef = int(rng.integers(4, 24)), int(rng.integers(20, 150)), int(rng.integers(1, 200))
    v = rows(elem, n, dims, ties); q = rows(elem, nq, dims, ties)
    rowids = np.sort(rng.choice(np.arange(1, 4 * n), size=n, replace=False)).astype("<i8") if rng.integers(0, 2) else None
    with vg.Slab(elem, dims) as s:
        s.load(v, rowids)
        for p in rng.choice(n, size=int(rng.integers(0, 20)), replace=False):
            s.delete(int(rowids[p]) if rowids is not None else int(p) + 1)
        if kind < 3:
            env(**fast); a = s.knn(q, k, metric)
            env(VECGPU_TC=0, VECGPU_HAM_BATCH=0, VECGPU_MERGE_SMALL=0); b = s.knn(q, k, metric)
            ok = all(np.array_equal(x.view("u1"), y.view("u1")) for x, y in zip(a, b))
            desc = f"kind={kind} elem={elem} metric={metric} dims={dims} n={n} nq={nq} k={k} ties={ties} sparse={rowids is not None} {fast}"
        else:
            bsz = int(rng.choice([64, 512, 0]))
            env(VECGPU_HNSW_DEVICE=1); i1 = vg.HnswIndex(s, metric, M=M, ef_construction=efc, seed=3); i1.rebuild(batch=bsz)
            e1 = i1.export_edges(); r1 = i1.search(q, k, ef_search=ef)
            env(VECGPU_HNSW_DEVICE=0); r2 = i1.search(q, k, ef_search=ef)       # lockstep walk of the device-built graph
            ok = all(np.array_equal(x.view("u1"), y.view("u1")) for x, y in zip(r1, r2))   # r1: 40 queries -> one CTA per query
            env(VECGPU_HNSW_DEVICE=1, VECGPU_HNSW_CTA_MAX_NQ=0); r3 = i1.search(q, k, ef_search=ef)   # one warp per query
            ok = ok and all(np.array_equal(x.view("u1"), y.view("u1")) for x, y in zip(r1, r3))
            os.environ["VECGPU_HNSW_DEVICE"] = "1"; os.environ["VECGPU_HNSW_LINK"] = "host"
            i2 = vg.HnswIndex(s, metric, M=M, ef_construction=efc, seed=3); i2.rebuild(batch=bsz)   # device walk, host linking
            os.environ.pop("VECGPU_HNSW_LINK")
            ok = ok and all(np.array_equal(x.view("u1"), y.view("u1")) for x, y in zip(e1, i2.export_edges()))
            if rng.integers(0, 3) == 0:
                env(VECGPU_HNSW_DEVICE=0); i3 = vg.HnswIndex(s, metric, M=M, ef_construction=efc, seed=3); i3.rebuild(batch=bsz)  # all lockstep
                ok = ok and all(np.array_equal(x.view("u1"), y.view("u1")) for x, y in zip(e1, i3.export_edges()))
                i3.close()
            desc = f"hnsw elem={elem} dims={dims} n={n} M={M} efc={efc} ef={ef} k={k} ties={ties} batch={bsz}"
            i1.close(); i2.close()
    cases += 1
    if not ok:
        fails += 1
        print("MISMATCH", desc, flush=True)
env()
print(f"fuzz_paths: {cases} cases, {fails} mismatches, {time.time() - t0:.0f} s")
sys.exit(1 if fails else 0)
