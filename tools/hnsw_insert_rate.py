#!/usr/bin/env python
"""Incremental HNSW inserts into a resident graph (vecgpu_hnsw_insert_appended) on cfg5's shape: N x f32[384] built, then rows
appended one by one (the reference's INSERT shape: 24-162 vec/s published) and in batches.   python tools/hnsw_insert_rate.py [rows]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle  # noqa: E402
import sqlite_vec_hnsw_b200 as vg  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
dims = 384
s = vg.Slab(0, dims)
s.fill_synthetic(seed=6, n=n, kind=1)
idx = vg.HnswIndex(s, 0, M=16, ef_construction=200, seed=1)
t0 = time.perf_counter()
idx.rebuild()
print(f"rebuild {n} x {dims}: {time.perf_counter() - t0:.2f} s  {idx.device_stats()}", flush=True)
new = oracle.synth_rows(0, 99, 1, 12_200, dims, 1)
pos = 0
for count, label in ((200, "one row per call"),):
    s.append(new[pos:pos + 1])            # (the first append past the loaded rows grows the slab's allocation: not timed)
    assert idx.insert_appended(batch=1) == 1
    pos += 1
    t0 = time.perf_counter()
    for i in range(count):
        s.append(new[pos + i:pos + i + 1])
        assert idx.insert_appended(batch=1) == 1
    el = time.perf_counter() - t0
    pos += count
    print(f"{label}: {count} inserts in {el:.3f} s = {count / el:.0f} vec/s ({el / count * 1e3:.2f} ms each)", flush=True)
for bsz in (100, 1000, 10_000):
    t0 = time.perf_counter()
    s.append(new[pos:pos + bsz])
    t1 = time.perf_counter()
    assert idx.insert_appended() == bsz
    el = time.perf_counter() - t1
    pos += bsz
    print(f"{bsz} rows per call: slab append {(t1 - t0) * 1e3:.1f} ms, graph insert {el * 1e3:.1f} ms = {bsz / el:.0f} vec/s  {idx.device_stats()}", flush=True)
t0 = time.perf_counter()
for i in range(20):                      # Vec0Tab::update: the row gets a new vector, its node is deleted and inserted again
    rid = 1000 + 37 * i
    s.upsert(rid, new[12_000 + i].tobytes())
    idx.reinsert(rid)
el = time.perf_counter() - t0
print(f"update + reinsert: 20 rows in {el:.3f} s = {el / 20 * 1e3:.1f} ms each", flush=True)
q = new[:200]
r, d, c = idx.search(q, 1, ef_search=64)
print(f"the first 200 inserted rows find themselves: {(r[:, 0] == n + 1 + np.arange(200)).mean():.3f}; nodes {idx.stats()['nodes']}", flush=True)
idx.close()
s.close()

# ---- rows inserted OUT of rowid order (vecgpu_slab_upsert between existing rows + vecgpu_hnsw_insert_at): every later row and
#      node id moves up by one, on the device.  Needs rowids with gaps, so this table is loaded with even rowids.
v = oracle.synth_rows(0, 6, 1, n, dims, 1)
s = vg.Slab(0, dims)
s.load(v, np.arange(1, n + 1, dtype="<i8") * 2)
del v
idx = vg.HnswIndex(s, 0, M=16, ef_construction=200, seed=1)
idx.rebuild()
s.upsert(2 * n + 2, new[0].tobytes())     # grows the allocations once (not timed)
idx.insert_appended(batch=1)
rng = np.random.default_rng(1)
rids = (rng.choice(n - 2, 40, replace=False).astype(np.int64) + 1) * 2 + 1
for lo, hi, label in ((0, 20, "anywhere in the table"),):
    t_slab = t_idx = 0.0
    for i in range(lo, hi):
        t0 = time.perf_counter()
        s.upsert(int(rids[i]), new[100 + i].tobytes())
        t1 = time.perf_counter()
        idx.insert_at(int(rids[i]))
        t2 = time.perf_counter()
        t_slab += t1 - t0
        t_idx += t2 - t1
    print(f"out-of-order insert ({label}): {hi - lo} rows, slab upsert {t_slab / (hi - lo) * 1e3:.2f} ms + graph insert_at {t_idx / (hi - lo) * 1e3:.2f} ms each", flush=True)
r, d, c = idx.search(new[100:120], 1, ef_search=64)
print(f"the 20 rows inserted out of order find themselves: {(r[:, 0] == rids[:20]).mean():.3f}; nodes {idx.stats()['nodes']}", flush=True)
idx.close()
s.close()
