"""CPU baselines for every BASELINE.json config: the oracle port of the reference path on the host cores of
the box, vectors contiguous in RAM, plus a "reference-shaped" figure for config 1 that pays one SQLite lookup per row as
src/vtab.rs:2594-2616 does.  Reported numbers, not targets."""
import os, sqlite3, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle

T = oracle.num_threads()
def timed(fn, min_s=2.0):
    fn(); n = 0; t0 = time.perf_counter()
    while time.perf_counter() - t0 < min_s:
        fn(); n += 1
    return (time.perf_counter() - t0) / n

print(f"host threads used by the oracle: {T}")
# ---- config 1: 10k x f32[384], L2 k=10, 100 queries, in full
v = oracle.synth_rows(0, 1, 1, 10_000, 384, 0); q = oracle.synth_rows(0, 2, 1, 100, 384, 0)
dt = timed(lambda: oracle.knn(0, 384, v, q, 10, 0))
print(f"cfg1 10k x f32[384] L2 k=10, 100 queries (contiguous RAM): {100 / dt:9.0f} q/s  ({10_000 * 1536 * 100 / dt / 1e9:6.1f} GB/s)")
db = sqlite3.connect(":memory:")
db.execute("CREATE TABLE t_data(rowid INTEGER PRIMARY KEY, vec00 BLOB)")
db.executemany("INSERT INTO t_data VALUES (?, ?)", [(i + 1, v[i].tobytes()) for i in range(len(v))])
def ref_shaped(nq=3):
    for qi in range(nq):
        rowids = [r[0] for r in db.execute("SELECT rowid FROM t_data ORDER BY rowid")]          # shadow.rs:853-868
        rows = np.empty((len(rowids), 384), dtype="<f4")
        for j, rid in enumerate(rowids):                                                        # one lookup per row, vtab.rs:2594-2616
            rows[j] = np.frombuffer(db.execute("SELECT vec00 FROM t_data WHERE rowid = ?", (rid,)).fetchone()[0], dtype="<f4")
        oracle.knn(0, 384, rows, q[qi : qi + 1], 10, 0)
t0 = time.perf_counter(); ref_shaped(); dt = (time.perf_counter() - t0) / 3
print(f"cfg1 reference-shaped (python sqlite3: rowid list + one SELECT per row + scan): {1 / dt:9.1f} q/s  -> the per-row lookups, not the arithmetic, are the reference's cost")
# ---- configs 2-4 on a prefix, bytes/s so it extrapolates
for name, elem, dims, metric, k, n, kind, full in (("cfg2 f32[768] cosine k=10", 0, 768, 2, 10, 400_000, 1, 10_000_000),
                                                    ("cfg3 i8[1024] L2 k=100", 1, 1024, 0, 100, 1_000_000, 0, 50_000_000),
                                                    ("cfg4 bit[1024] Hamming k=10", 2, 1024, 3, 10, 4_000_000, 0, 500_000_000)):
    v = oracle.synth_rows(elem, 3, 1, n, dims, kind); q = oracle.synth_rows(elem, 4, 1, 4, dims, kind)
    dt = timed(lambda: oracle.knn(elem, dims, v, q[:1], k, metric))
    rb = oracle.row_bytes(elem, dims)
    print(f"{name}: {1 / dt:8.2f} q/s on a {n}-row prefix = {n * rb / dt / 1e9:6.1f} GB/s -> {1 / dt * n / full:8.3f} q/s at the full {full} rows (single query)")
    dt = timed(lambda: oracle.knn(elem, dims, v, q, k, metric))
    print(f"{'':{len(name)}}  4 queries per call: {4 / dt:8.2f} q/s on the prefix")
    del v
# ---- config 5: distances/s of the candidate scoring arithmetic (384-d f32 L2)
v = oracle.synth_rows(0, 6, 1, 1_000_000, 384, 1); q = oracle.synth_rows(0, 7, 1, 1, 384, 1)
dt = timed(lambda: oracle.distances(0, 384, v, q[0], 0))
print(f"cfg5 candidate scoring, f32[384] L2: {1_000_000 / dt / 1e6:7.1f} M distances/s ({1_000_000 * 1536 / dt / 1e9:5.1f} GB/s) on contiguous rows, all threads;")
idx = np.random.default_rng(1).integers(0, 1_000_000, size=32)
dt = timed(lambda: oracle.distances(0, 384, v[idx], q[0], 0), 1.0)
print(f"     one expansion step as the reference does it (gather 32 rows + 32 distances): {32 / dt / 1e6:7.3f} M distances/s")
