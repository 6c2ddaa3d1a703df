"""Differential fuzz of the slab's write paths on the GPU: random appends, in-place updates, deletes, inserts between existing
rowids, re-used rowids, emptied rows and compactions, interleaved with queries through the batched fast paths (tensor-core
f32 / int8 with their cached row norms, lane-per-query Hamming) and the streaming scan.  Every answer must equal the CPU
oracle's scan of a plain {rowid: vector} model of the table, rowids and distance bits.   python tools/fuzz_writes.py [seconds] [seed]"""
import os, signal, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle
import sqlite_vec_hnsw_b200 as vg
budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
signal.alarm(int(budget) + 120)
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 77)
os.environ["VECGPU_TC_MIN_WORK"] = "0"        # small tables take the tensor-core route too
os.environ["VECGPU_SHIFT_CHUNK_MB"] = "1"     # several chunks per out-of-order insert
def rows(elem, n, dims):
    if elem == 0: return rng.standard_normal((n, dims)).astype("<f4")
    if elem == 1: return rng.integers(-128, 128, size=(n, dims)).astype("i1")
    nb = (dims + 7) // 8
    r = rng.integers(0, 256, size=(n, nb)).astype("u1")
    if dims % 8: r[:, -1] &= (1 << (dims % 8)) - 1
    return r
t0, cases, checks, fails, ops_done = time.time(), 0, 0, 0, {}
while time.time() - t0 < budget:
    elem = int(rng.integers(0, 3))
    metric = [int(rng.choice([0, 2])), 0, 3][elem]
    dims = int(rng.integers(16, 200)) if elem < 2 else int(rng.integers(8, 513))
    n, nq, k = int(rng.integers(3000, 30000)), int(rng.integers(16, 64)), int(rng.integers(1, 40))
    v = rows(elem, n, dims)
    ids = np.sort(rng.choice(np.arange(1, 4 * n), size=n, replace=False)).astype("<i8")
    model = {int(r): v[i] for i, r in enumerate(ids)}
    dead = []
    with vg.Slab(elem, dims) as s:
        s.load(v, ids)
        q = rows(elem, nq, dims)
        if elem == 0 and metric == 2 and rng.integers(0, 2):   # cosine: a zero row sits on the always-re-ranked list
            z = int(ids[rng.integers(n)])
            model[z] = np.zeros(dims, dtype="<f4"); s.upsert(z, model[z].tobytes())
        for step in range(int(rng.integers(20, 60))):
            op = rng.choice(["append", "between", "update", "delete", "reuse", "empty", "compact", "query"],
                            p=[0.12, 0.25, 0.12, 0.1, 0.06, 0.03, 0.02, 0.3])
            keys = list(model)
            if op == "append":
                rid = max(max(keys), max(dead, default=0)) + int(rng.integers(1, 5)); model[rid] = rows(elem, 1, dims)[0]; s.upsert(rid, model[rid].tobytes())
            elif op == "between":
                rid = int(rng.integers(1, max(keys)))
                if rid in model or rid in dead: continue
                model[rid] = rows(elem, 1, dims)[0]; s.upsert(rid, model[rid].tobytes())
            elif op == "update":
                rid = int(rng.choice(keys)); model[rid] = rows(elem, 1, dims)[0]; s.upsert(rid, model[rid].tobytes())
            elif op == "delete":
                rid = int(rng.choice(keys)); del model[rid]; dead.append(rid); s.delete(rid)
            elif op == "reuse":
                if not dead: continue
                rid = dead.pop(int(rng.integers(len(dead)))); model[rid] = rows(elem, 1, dims)[0]; s.upsert(rid, model[rid].tobytes())
            elif op == "empty":
                rid = int(rng.choice(keys)); del model[rid]; dead.append(rid); s.upsert(rid, b"")
            elif op == "compact":
                s.compact(); dead = []
            else:
                rr = np.array(sorted(model), dtype="<i8")
                vv = np.stack([model[int(x)] for x in rr])
                batch = bool(rng.integers(0, 2))
                qq = q if batch else q[:1]
                r, d, c = s.knn(qq, k, metric)
                er, ed, ec = oracle.knn(elem, dims, vv, qq, k, metric, rowids=rr)
                ok = np.array_equal(r, er) and np.array_equal(d.view("<u4"), ed.view("<u4")) and np.array_equal(c, ec)
                checks += 1
                if not ok:
                    fails += 1
                    print(f"MISMATCH elem={elem} metric={metric} dims={dims} n={n} nq={len(qq)} k={k} step={step}", flush=True)
                    break
            ops_done[str(op)] = ops_done.get(str(op), 0) + 1
    cases += 1
print(f"fuzz_writes: {cases} tables, {checks} checked queries, {fails} mismatches, {time.time() - t0:.0f} s; operations {dict(sorted(ops_done.items()))}; "
      f"tensor-core queries {vg.tc_stats()[0]}")
sys.exit(1 if fails else 0)
