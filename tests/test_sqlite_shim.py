"""SQLite-side shim: the reference's on-disk format read / written with Python's sqlite3 and served from a slab."""
import sqlite3

import numpy as np
import pytest

from helpers import COSINE, F32, I8, L2, random_rows


def _shim():
    from sqlite_vec_hnsw_b200 import sqlite_shim

    return sqlite_shim


def test_shadow_ddl_matches_reference():
    sh = _shim()
    # src/shadow.rs:111-129
    assert sh.data_table_ddl("t") == 'CREATE TABLE "main"."t_data" (rowid INTEGER PRIMARY KEY, vec00 BLOB);'
    assert (sh.data_table_ddl("docs", 2, ["TEXT", "INTEGER"])
            == 'CREATE TABLE "main"."docs_data" (rowid INTEGER PRIMARY KEY, vec00 BLOB, vec01 BLOB, col00 TEXT, col01 INTEGER);')
    conn = sqlite3.connect(":memory:")
    sh.create_shadow_tables(conn, "t", 1, ["TEXT"])
    info = dict(conn.execute('SELECT key, value FROM "t_info"').fetchall())
    assert info["STORAGE_SCHEMA"] == "unified" and info["CREATE_VERSION"] == "0.2.0"  # src/shadow.rs:141-181
    assert sh.storage_schema(conn, "t") == "unified"


def test_read_column_order_and_skips():
    sh = _shim()
    conn = sqlite3.connect(":memory:")
    sh.create_shadow_tables(conn, "t")
    v = random_rows(F32, 6, 4, seed=1)
    for rid, blob in [(7, v[0].tobytes()), (2, v[1].tobytes()), (9, None), (4, b""), (5, v[2].tobytes()[:8]), (11, v[3].tobytes())]:
        conn.execute('INSERT INTO "t_data" (rowid, vec00) VALUES (?, ?)', (rid, blob))
    rowids, vec, skip = sh.read_column(conn, "t", 0, 16)
    assert list(rowids) == [2, 4, 5, 7, 9, 11]  # ORDER BY rowid (src/shadow.rs:853-868)
    assert list(skip) == [0, 1, 1, 0, 1, 0]  # empty / wrong length / NULL are skipped (src/vtab.rs:2596-2613)
    assert vec[0].tobytes() == v[1].tobytes() and vec[3].tobytes() == v[0].tobytes() and not vec[1].any()


class _RecorderSlab:
    """Stands in for the HBM slab on CPU-only hosts: records the staging calls."""

    def __init__(self):
        self.calls = []

    def load(self, vec, rowids):
        self.calls.append(("load", len(rowids)))

    def upsert(self, rowid, blob):
        self.calls.append(("upsert", rowid, len(blob)))

    def delete(self, rowid):
        self.calls.append(("delete", rowid))

    def close(self):
        pass


def test_staging_and_hooks_drive_the_slab():
    sh = _shim()
    conn = sqlite3.connect(":memory:")
    sh.create_shadow_tables(conn, "t")
    rec = _RecorderSlab()
    t = sh.Vec0Table(conn, "t", F32, 4, slab_factory=lambda: rec)
    assert t.is_stale()
    v = random_rows(F32, 3, 4, seed=2)
    conn.execute('INSERT INTO "t_data" (rowid, vec00) VALUES (1, ?)', (v[0].tobytes(),))
    conn.execute('INSERT INTO "t_data" (rowid, vec00) VALUES (2, ?)', (b"",))
    assert t.stage() == 2 and rec.calls == [("load", 2), ("upsert", 2, 0)] and not t.is_stale()
    assert t.insert(v[1].tobytes()) == 3  # auto rowid = MAX+1 (src/shadow.rs:888-900)
    t.update(1, v[2].tobytes())
    t.delete(3)
    assert rec.calls[2:] == [("upsert", 3, 16), ("upsert", 1, 16), ("delete", 3)] and not t.is_stale()
    conn.execute('INSERT INTO "t_data" (rowid, vec00) VALUES (50, ?)', (v[0].tobytes(),))  # a writer that bypasses the hooks
    assert t.is_stale()
    with pytest.raises(Exception):
        t.insert(b"\x00" * 5)  # wrong-length vectors are rejected on insert (src/vtab.rs:1474-1498)


@pytest.mark.gpu
def test_end_to_end_knn_from_sqlite(vg, orc, gpu, tmp_path):
    sh = _shim()
    conn = sqlite3.connect(str(tmp_path / "vec.db"))  # a file database, like tests/test_disk_persistence.rs
    sh.create_shadow_tables(conn, "items", 1, ["TEXT"])
    n, dims = 3000, 24
    v = random_rows(F32, n, dims, seed=3)
    rowids = np.arange(n) * 2 + 1
    conn.executemany('INSERT INTO "items_data" (rowid, vec00, col00) VALUES (?, ?, ?)',
                     [(int(r), v[i].tobytes(), f"doc{i}") for i, r in enumerate(rowids)])
    conn.execute('UPDATE "items_data" SET vec00 = NULL WHERE rowid = 5')
    conn.commit()
    t = sh.Vec0Table(conn, "items", F32, dims)  # default metric: cosine
    skip = np.zeros(n, dtype="u1")
    skip[2] = 1
    res = t.knn(v[10].tobytes(), 5)
    er, ed, _ = orc.knn(F32, dims, v, v[10], 5, COSINE, rowids=rowids, skip=skip)
    assert [r for r, _ in res] == list(er[0]) and np.array_equal(np.array([d for _, d in res], dtype="<f4").view("<u4"), ed[0].view("<u4"))
    # JSON query text == blob (tests/integration_test.rs:1076-1128)
    js = "[" + ",".join(repr(float(x)) for x in v[10]) + "]"
    assert t.knn(js, 5) == res
    # write through the hooks, then query again
    nv = random_rows(F32, 2, dims, seed=4)
    new_id = t.insert(nv[0].tobytes())
    t.update(int(rowids[0]), nv[1].tobytes())
    t.delete(int(rowids[10]))
    v2 = np.concatenate([v, nv[:1]])
    v2[0] = nv[1]
    rowids2 = np.concatenate([rowids, [new_id]])
    skip2 = np.concatenate([skip, [0]]).astype("u1")
    skip2[10] = 1
    res = t.knn(nv[0].tobytes(), 4, metric=L2)
    er, ed, _ = orc.knn(F32, dims, v2, nv[0], 4, L2, rowids=rowids2, skip=skip2)
    assert [r for r, _ in res] == list(er[0]) and res[0] == (new_id, 0.0)
    # compaction of the resident slab keeps every answer (the deleted row and the NULL blob are physically gone)
    assert t.compact(min_dead_fraction=0.5) == 0 and t.compact() == 2
    assert t.knn(nv[0].tobytes(), 4, metric=L2) == res and t.slab.count()[0] == t.slab.count()[1]
    # an external writer is noticed and the slab re-staged
    conn.execute('DELETE FROM "items_data" WHERE rowid = ?', (new_id,))
    res = t.knn(nv[0].tobytes(), 1, metric=L2)
    assert res[0][0] != new_id
    t.close()


@pytest.mark.gpu
@pytest.mark.parametrize("metric_s", ["l2", "cosine"])
def test_rebuild_hnsw_writes_the_reference_shadow_tables(vg, orc, gpu, tmp_path, metric_s):
    # vec_rebuild_hnsw (src/sql_functions.rs:436-534): graph built on the GPU, written back into the reference's
    # "{t}_{c}_hnsw_nodes / _hnsw_edges / _hnsw_meta" tables (src/shadow.rs:407-500); a literal restatement of search_hnsw
    # that reads ONLY those tables (neighbours in primary-key order, as the reference's SELECT returns them) must find
    # what the resident index finds
    from oracle import hnsw_ref

    sh = _shim()
    conn = sqlite3.connect(str(tmp_path / "hnsw.db"))
    sh.create_shadow_tables(conn, "docs", 1, [])
    n, dims = 2500, 20
    v = random_rows(F32, n, dims, seed=7)
    conn.executemany('INSERT INTO "docs_data" (rowid, vec00) VALUES (?, ?)', [(i + 1, v[i].tobytes()) for i in range(n)])
    conn.execute('UPDATE "docs_data" SET vec00 = NULL WHERE rowid = 17')
    sh.create_hnsw_shadow_tables(conn, "docs", "emb", dims, "float32", metric_s, m=12, ef_construction=80)
    metric = COSINE if metric_s == "cosine" else L2
    t = sh.Vec0Table(conn, "docs", F32, dims, distance_metric=metric)
    assert t.rebuild_hnsw("emb") == n - 1
    m, max_m0, efc, ep, epl, num_nodes = conn.execute(
        'SELECT m, max_m0, ef_construction, entry_point_rowid, entry_point_level, num_nodes FROM "docs_emb_hnsw_meta"').fetchone()
    assert (m, max_m0, efc, num_nodes) == (12, 24, 80, n - 1) and ep >= 1 and epl >= 0
    assert conn.execute('SELECT COUNT(*), MIN(level), COUNT(vector) FROM "docs_emb_hnsw_nodes"').fetchone()[::2] == (n - 1, n - 1)
    assert conn.execute('SELECT COUNT(*) FROM "docs_emb_hnsw_nodes" WHERE rowid = 17').fetchone()[0] == 0
    assert conn.execute('SELECT level FROM "docs_emb_hnsw_nodes" WHERE rowid = ?', (ep,)).fetchone()[0] == epl
    deg = conn.execute('SELECT MAX(c) FROM (SELECT COUNT(*) AS c FROM "docs_emb_hnsw_edges" WHERE level = 0 GROUP BY from_rowid)').fetchone()[0]
    assert deg <= max_m0
    # stored node vectors: the column's blobs for L2, unit vectors for cosine (src/hnsw/insert.rs:300-322)
    blob = conn.execute('SELECT vector FROM "docs_emb_hnsw_nodes" WHERE rowid = 5').fetchone()[0]
    want = vg.normalize(v[4:5])[0] if metric_s == "cosine" else v[4]
    assert np.array_equal(np.frombuffer(blob, dtype="<f4").view("<u4"), want.view("<u4"))
    # walk the graph from SQLite alone
    stored = {r: np.frombuffer(b, dtype="<f4") for r, b in conn.execute('SELECT rowid, vector FROM "docs_emb_hnsw_nodes"')}
    def nbrs(node, level):
        return [r[0] for r in conn.execute('SELECT to_rowid FROM "docs_emb_hnsw_edges" WHERE from_rowid = ? AND level = ? ORDER BY to_rowid', (node, level))]
    q = random_rows(F32, 5, dims, seed=8)
    for qi in range(len(q)):
        qs = vg.normalize(q[qi : qi + 1])[0] if metric_s == "cosine" else q[qi]
        walk = hnsw_ref.search_hnsw(lambda rid: float(orc.distance(F32, qs, stored[rid], L2)), nbrs, ep, epl, 8, 60)
        got = t.hnsw_knn(q[qi].tobytes(), 8, ef_search=60)
        assert [r for r, _ in got] == [w[0] for w in walk]
        want_d = [orc.convert_cosine_output(w[1]) if metric_s == "cosine" else w[1] for w in walk]
        assert np.allclose([d for _, d in got], want_d, rtol=1e-6, atol=1e-7)
    t.close()


@pytest.mark.gpu
def test_hnsw_hooks_after_rebuild_delete_insert_update(vg, orc, gpu, tmp_path):
    """Vec0Tab::delete (src/vtab.rs:1340-1397) removes the node, its edges in both directions and fixes the meta row;
    the resident index never returns a deleted rowid; rows inserted in rowid order and updated rows are (re)inserted into the
    resident graph at once; a row inserted out of rowid order makes it stale until it is rebuilt."""
    sh = _shim()
    conn = sqlite3.connect(str(tmp_path / "hooks.db"))
    sh.create_shadow_tables(conn, "docs", 1, [])
    n, dims = 1200, 16
    v = random_rows(F32, n, dims, seed=17)
    conn.executemany('INSERT INTO "docs_data" (rowid, vec00) VALUES (?, ?)', [(i + 1, v[i].tobytes()) for i in range(n)])
    sh.create_hnsw_shadow_tables(conn, "docs", "emb", dims, "float32", "cosine", m=8, ef_construction=60)
    t = sh.Vec0Table(conn, "docs", F32, dims, distance_metric=COSINE)
    t.rebuild_hnsw("emb")
    ep, num0, ver0 = conn.execute('SELECT entry_point_rowid, num_nodes, hnsw_version FROM "docs_emb_hnsw_meta"').fetchone()
    assert t.hnsw_knn(v[99].tobytes(), 3)[0][0] == 100
    # delete an ordinary node and the entry point itself
    for rid in (100, ep):
        t.delete(rid)
        assert conn.execute('SELECT COUNT(*) FROM "docs_emb_hnsw_nodes" WHERE rowid = ?', (rid,)).fetchone()[0] == 0
        assert conn.execute('SELECT COUNT(*) FROM "docs_emb_hnsw_edges" WHERE from_rowid = ? OR to_rowid = ?', (rid, rid)).fetchone()[0] == 0
    ep2, epl2, num1, ver1 = conn.execute('SELECT entry_point_rowid, entry_point_level, num_nodes, hnsw_version FROM "docs_emb_hnsw_meta"').fetchone()
    assert num1 == num0 - 2 and ver1 == ver0 + 2 and ep2 not in (ep, 100)
    assert conn.execute('SELECT MAX(level) FROM "docs_emb_hnsw_nodes"').fetchone()[0] == epl2
    got = [r for r, _ in t.hnsw_knn(v[99].tobytes(), 10, ef_search=80)]
    assert 100 not in got and ep not in got and len(got) == 10
    # the resident graph lost the two nodes as well, and its entry point is a live node again
    assert t._hnsw.stats()["nodes"] == n - 2 and t._hnsw.entry_point()[0] not in (ep, 100)
    fr, to, _, _ = t._hnsw.export_edges()
    assert not np.isin(fr, [ep, 100]).any() and not np.isin(to, [ep, 100]).any()
    # the exact scan agrees that both are gone
    exact = [r for r, _ in t.knn(v[99].tobytes(), 10)]
    assert 100 not in exact and ep not in exact
    # rows that arrive in rowid order are inserted into the resident graph at once (Vec0Tab::insert -> insert_hnsw,
    # src/vtab.rs:1409); the shadow tables follow at the next flush
    new = random_rows(F32, 3, dims, seed=18)
    ids = [t.insert(new[i].tobytes()) for i in range(3)]
    assert ids == [n + 1, n + 2, n + 3]
    for i in range(3):
        assert t.hnsw_knn(new[i].tobytes(), 3)[0][0] == ids[i]
    assert conn.execute('SELECT COUNT(*) FROM "docs_emb_hnsw_nodes" WHERE rowid > ?', (n,)).fetchone()[0] == 0
    assert t.flush_hnsw_shadow() == n - 2 + 3
    assert conn.execute('SELECT COUNT(*) FROM "docs_emb_hnsw_nodes" WHERE rowid > ?', (n,)).fetchone()[0] == 3
    assert conn.execute('SELECT COUNT(*) FROM "docs_emb_hnsw_edges" WHERE from_rowid = ?', (ids[0],)).fetchone()[0] > 0
    # a re-used rowid (100 was deleted above; its tombstoned row still sits in the slab): no row moves, the node goes straight in
    old_id = t.insert(new[0].tobytes(), rowid=100)
    assert {r for r, _ in t.hnsw_knn(new[0].tobytes(), 3)[:2]} == {old_id, ids[0]}
    assert t._hnsw.stats()["nodes"] == n - 2 + 3 + 1
    # Vec0Tab::update (src/vtab.rs:1860-1895): the node is deleted and inserted again with the new vector
    t.update(5, new[2].tobytes())
    got5 = t.hnsw_knn(new[2].tobytes(), 3)
    assert {got5[0][0], got5[1][0]} == {5, ids[2]} and got5[0][1] < 1e-6 and got5[1][1] < 1e-6
    t.update(6, b"")                                 # emptied: leaves the index
    assert 6 not in [r for r, _ in t.hnsw_knn(v[5].tobytes(), 10)]
    t.update(5, new[1].tobytes())
    got5 = t.hnsw_knn(new[1].tobytes(), 3)
    assert {got5[0][0], got5[1][0]} == {5, ids[1]} and got5[1][1] < 1e-6
    t.close()


@pytest.mark.gpu
@pytest.mark.parametrize("quant", [None, "int8"])
def test_hnsw_hook_for_a_row_inserted_out_of_rowid_order(vg, orc, gpu, tmp_path, quant):
    """Vec0Tab::insert with an explicit rowid BETWEEN existing ones (src/vtab.rs:1409-1682): the slab keeps rowid order, so every
    later row moves up one position; the resident graph is renumbered on the device (vecgpu_hnsw_insert_at) and answers at
    once — no rebuild — for a plain cosine column (stored = normalised) and an int8-quantised one (stored slab of its own)."""
    sh = _shim()
    conn = sqlite3.connect(str(tmp_path / "gap.db"))
    sh.create_shadow_tables(conn, "docs", 1, [])
    n, dims = 900, 32
    v = random_rows(F32, n + 6, dims, seed=23)
    conn.executemany('INSERT INTO "docs_data" (rowid, vec00) VALUES (?, ?)', [(10 * (i + 1), v[i].tobytes()) for i in range(n)])
    kw = {} if quant is None else {"index_quantization": quant}
    sh.create_hnsw_shadow_tables(conn, "docs", "emb", dims, "float32", "cosine", m=8, ef_construction=60, **kw)
    t = sh.Vec0Table(conn, "docs", F32, dims, distance_metric=COSINE)
    assert t.rebuild_hnsw("emb") == n
    for j, rid in enumerate([5, 4567, 8995, 15, 4568, 10 * n + 10]):   # first, middle, last gap, ..., and one appended row
        assert t.insert(v[n + j].tobytes(), rowid=rid) == rid
        got = t.hnsw_knn(v[n + j].tobytes(), 3, ef_search=80)
        assert got[0][0] == rid and got[0][1] < 1e-3, (rid, got)
        assert t.knn(v[n + j].tobytes(), 1)[0][0] == rid
    assert not getattr(t, "_hnsw_stale", False) and t._hnsw.stats()["nodes"] == n + 6
    # the old rows are still found where they are now (every position behind rowid 5 moved six times)
    for i in (0, 1, 455, 456, 457, n - 1):
        assert t.hnsw_knn(v[i].tobytes(), 1, ef_search=80)[0][0] == 10 * (i + 1)
    assert t.flush_hnsw_shadow() == n + 6
    assert conn.execute('SELECT COUNT(*) FROM "docs_emb_hnsw_nodes" WHERE rowid IN (5, 15, 4567, 4568, 8995)').fetchone()[0] == 5
    assert conn.execute('SELECT COUNT(*) FROM "docs_emb_hnsw_edges" WHERE from_rowid = 4567').fetchone()[0] > 0
    t.close()


@pytest.mark.gpu
def test_rebuild_hnsw_with_int8_index_quantization(vg, orc, gpu, tmp_path):
    """index_quantization=int8 on a cosine float32 column: stored node vectors are quantize_int8_for_index(normalize(v))
    (src/hnsw/insert.rs:300-322), the query likewise (src/hnsw/search.rs:285-302), the graph is walked with int8 L2; the
    reference's bar is recall@10 >= 90 % against the float32 ground truth (tests/test_quantization_perf.rs:194-289)."""
    sh = _shim()
    conn = sqlite3.connect(str(tmp_path / "q8.db"))
    sh.create_shadow_tables(conn, "docs", 1, [])
    n, dims, nq = 5000, 128, 20
    rng = np.random.default_rng(5)
    v = rng.standard_normal((n, dims)).astype("<f4")
    q = v[rng.choice(n, nq, replace=False)] + 0.05 * rng.standard_normal((nq, dims)).astype("<f4")
    conn.executemany('INSERT INTO "docs_data" (rowid, vec00) VALUES (?, ?)', [(i + 1, v[i].tobytes()) for i in range(n)])
    sh.create_hnsw_shadow_tables(conn, "docs", "emb", dims, "float32", "cosine", m=32, ef_construction=400, index_quantization="int8")
    t = sh.Vec0Table(conn, "docs", F32, dims, distance_metric=COSINE)
    assert t.rebuild_hnsw("emb") == n
    # stored blobs are int8[dims], bit-identical to the reference pipeline restated by the oracle
    blob = conn.execute('SELECT vector FROM "docs_emb_hnsw_nodes" WHERE rowid = 7').fetchone()[0]
    assert len(blob) == dims
    assert np.array_equal(np.frombuffer(blob, dtype="i1"), orc.quantize_int8_for_index(orc.normalize(v[6:7]))[0])
    er, _, _ = orc.knn(F32, dims, v, q, 10, COSINE)
    hit = 0
    for qi in range(nq):
        got = [r for r, _ in t.hnsw_knn(q[qi].tobytes(), 10, ef_search=200)]
        hit += len(set(got) & set(er[qi].tolist()))
    assert hit / (10 * nq) >= 0.90
    # the walk equals the sequential oracle's walk over the same stored vectors when the graph is built one insert at a time
    t.close()
