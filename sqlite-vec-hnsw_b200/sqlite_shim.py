"""vec0-compatible SQLite reader that exercises the C-ABI boundary end to end.

The reference stores a vec0 table as plain SQLite shadow tables (src/shadow.rs:105-184):

    "{t}_data"(rowid INTEGER PRIMARY KEY, vec00 BLOB, vec01 BLOB, ..., col00 T, ...)   one BLOB per row and vector column
    "{t}_info"(key TEXT PRIMARY KEY, value)                                            CREATE_VERSION*, STORAGE_SCHEMA='unified'

A BLOB is the raw little-endian element array (src/vector.rs:223-242, 592-600).  This module reads and writes that
format with Python's sqlite3 — no Rust extension needed — and keeps an HBM slab in step with it:

    stage()                      SELECT rowid, vecNN FROM "{t}_data" ORDER BY rowid  ->  vecgpu_slab_load
                                 (src/shadow.rs:853-868 order; NULL / empty / wrong-length blobs become skipped rows,
                                 src/vtab.rs:2596-2613)
    knn(query_blob, k)           the body of brute_force_search (src/vtab.rs:2573-2623) on the GPU
    insert / update / delete     the hooks of Vec0Tab::insert/update/delete (src/vtab.rs:1409, 1684, 1326): SQLite first,
                                 then vecgpu_slab_upsert / vecgpu_slab_delete
    is_stale()                   (COUNT(*), MAX(rowid), total_changes) heuristic of SURVEY H5 for writers that bypass us
"""
import numpy as np

from . import vec0


def data_table_ddl(table, n_vector_columns=1, data_columns=()):
    """The CREATE TABLE statement of src/shadow.rs:111-129 (schema "main")."""
    sql = f'CREATE TABLE "main"."{table}_data" (rowid INTEGER PRIMARY KEY'
    for i in range(n_vector_columns):
        sql += f", vec{i:02d} BLOB"
    for i, col_type in enumerate(data_columns):
        sql += f", col{i:02d} {col_type}"
    return sql + ");"


INFO_ROWS = [  # src/shadow.rs:141-181
    ("CREATE_VERSION", "0.2.0"), ("CREATE_VERSION_MAJOR", 0), ("CREATE_VERSION_MINOR", 2), ("CREATE_VERSION_PATCH", 0),
    ("STORAGE_SCHEMA", "unified"),
]


def create_shadow_tables(conn, table, n_vector_columns=1, data_columns=()):
    """Create "{t}_data" and "{t}_info" exactly as the reference does (src/shadow.rs:105-184)."""
    conn.execute(data_table_ddl(table, n_vector_columns, data_columns))
    conn.execute(f'CREATE TABLE "main"."{table}_info" (key TEXT PRIMARY KEY, value);')
    conn.executemany(f'INSERT INTO "main"."{table}_info" (key, value) VALUES (?, ?)', INFO_ROWS)


def create_hnsw_shadow_tables(conn, table, column, dims, element_type="float32", distance_metric="l2", m=32, ef_construction=400,
                              index_quantization="none", rng_seed=12345):
    """"{t}_{c}_hnsw_meta" (single row), "_hnsw_nodes", "_hnsw_edges" as the reference creates them (src/shadow.rs:407-500,
    508-600): same columns, defaults and the WITHOUT ROWID primary key (from_rowid, level, to_rowid) of the edge table."""
    conn.execute(
        f'CREATE TABLE IF NOT EXISTS "{table}_{column}_hnsw_meta" ('
        "id INTEGER PRIMARY KEY CHECK (id = 1), m INTEGER NOT NULL DEFAULT 32, max_m0 INTEGER NOT NULL DEFAULT 64, "
        "ef_construction INTEGER NOT NULL DEFAULT 400, ef_search INTEGER NOT NULL DEFAULT 200, max_level INTEGER NOT NULL DEFAULT 16, "
        "level_factor REAL NOT NULL DEFAULT 0.28768207245178085, entry_point_rowid INTEGER NOT NULL DEFAULT -1, "
        "entry_point_level INTEGER NOT NULL DEFAULT -1, num_nodes INTEGER NOT NULL DEFAULT 0, dimensions INTEGER NOT NULL DEFAULT 0, "
        "element_type TEXT NOT NULL DEFAULT 'float32', distance_metric TEXT NOT NULL DEFAULT 'l2', rng_seed INTEGER NOT NULL DEFAULT 12345, "
        "hnsw_version INTEGER NOT NULL DEFAULT 1, index_quantization TEXT NOT NULL DEFAULT 'none', normalize_vectors INTEGER NOT NULL DEFAULT 1)")
    conn.execute(
        f'INSERT OR IGNORE INTO "{table}_{column}_hnsw_meta" '
        "(id, m, max_m0, ef_construction, dimensions, element_type, distance_metric, index_quantization, rng_seed, normalize_vectors) "
        "VALUES (1, ?, ?, ?, ?, ?, ?, ?, ?, ?)",
        (m, 2 * m, ef_construction, dims, element_type, distance_metric, index_quantization, rng_seed, 1 if distance_metric == "cosine" else 0))
    conn.execute(f'CREATE TABLE IF NOT EXISTS "{table}_{column}_hnsw_nodes" ('
                 "rowid INTEGER PRIMARY KEY, level INTEGER NOT NULL, vector BLOB, created_at INTEGER DEFAULT (unixepoch()))")
    conn.execute(f'CREATE TABLE IF NOT EXISTS "{table}_{column}_hnsw_edges" ('
                 "from_rowid INTEGER NOT NULL, to_rowid INTEGER NOT NULL, level INTEGER NOT NULL, distance REAL NOT NULL DEFAULT 0.0, "
                 "PRIMARY KEY (from_rowid, level, to_rowid)) WITHOUT ROWID")


def storage_schema(conn, table):
    row = conn.execute(f'SELECT value FROM "main"."{table}_info" WHERE key = \'STORAGE_SCHEMA\'').fetchone()
    return None if row is None else row[0]


def read_column(conn, table, column_idx, row_bytes, chunk=65536):
    """-> (rowids int64[n], vectors uint8[n, row_bytes], skip uint8[n]) in ascending rowid order
    (SELECT rowid, vecNN ... ORDER BY rowid: src/shadow.rs:853-868 / 721-740).  Rows are fetched in chunks and each chunk's
    well-formed blobs are joined with one b"".join (the per-row Python work is a length check)."""
    cur = conn.execute(f'SELECT rowid, vec{column_idx:02d} FROM "main"."{table}_data" ORDER BY rowid')
    ids, parts, skips = [], [], []
    zero = bytes(row_bytes)
    while True:
        rows = cur.fetchmany(chunk)
        if not rows:
            break
        rid = np.fromiter((r[0] for r in rows), dtype="<i8", count=len(rows))
        # empty / NULL / wrong length: the scan skips the row (src/vtab.rs:2596-2613)
        bad = np.fromiter((b is None or len(b) != row_bytes for _, b in rows), dtype=bool, count=len(rows))
        if bad.any():
            blob = b"".join(zero if x else bytes(b) for x, (_, b) in zip(bad.tolist(), rows))
        else:
            blob = b"".join(b for _, b in rows)
        ids.append(rid)
        parts.append(np.frombuffer(blob, dtype="u1").reshape(len(rows), row_bytes))
        skips.append(bad.astype("u1"))
    if not ids:
        return np.zeros(0, dtype="<i8"), np.zeros((0, row_bytes), dtype="u1"), np.zeros(0, dtype="u1")
    return np.concatenate(ids), np.concatenate(parts), np.concatenate(skips)


class Vec0Table:
    """One vector column of one vec0 table, served from an HBM slab."""

    def __init__(self, conn, table, vec_type, dims, distance_metric=vec0.DistanceMetric.Cosine, column_idx=0, device=0,
                 slab_factory=None):
        # the column's metric defaults to Cosine (src/vtab.rs:240-248)
        self.conn, self.table, self.column_idx = conn, table, column_idx
        self.vec_type = vec0.VectorType(vec_type)
        self.dims = int(dims)
        self.metric = vec0.DistanceMetric(distance_metric)
        self.row_bytes = self.vec_type.row_bytes(self.dims)
        schema = storage_schema(conn, table)
        if schema != "unified":
            raise vec0.InvalidState(f"{table}: unsupported STORAGE_SCHEMA {schema!r} (expected 'unified')")
        self._slab_factory = slab_factory or (lambda: vec0.Slab(self.vec_type, self.dims, device=device))
        self.slab = None
        self._fingerprint = None

    # ---- staging
    def _current_fingerprint(self):
        cnt, mx = self.conn.execute(f'SELECT COUNT(*), MAX(rowid) FROM "main"."{self.table}_data"').fetchone()
        return (cnt, mx, self.conn.total_changes)

    def is_stale(self):
        return self.slab is None or self._fingerprint != self._current_fingerprint()

    def stage(self):
        rowids, vec, skip = read_column(self.conn, self.table, self.column_idx, self.row_bytes)
        if self.slab is None:
            self.slab = self._slab_factory()
        if len(rowids):
            self.slab.load(vec, rowids)
            for i in np.flatnonzero(skip):
                self.slab.upsert(int(rowids[i]), b"")
        else:
            self.slab.load(np.zeros((0, self.row_bytes), dtype="u1"), np.zeros(0, dtype="<i8"))
        self._fingerprint = self._current_fingerprint()
        return len(rowids)

    # ---- the KNN arm of Vec0TabCursor::filter (src/vtab.rs:2286-2305)
    def knn(self, query, k, metric=None):
        """query: blob, or a JSON array text as vec_f32('[...]') accepts (src/vtab.rs:2119-2143)."""
        if isinstance(query, str):
            query = vec0.Vector.from_json(query, self.vec_type).as_bytes()
        if self.is_stale():
            self.stage()
        return vec0.brute_force_search(self.slab, query, k, self.metric if metric is None else metric)

    # ---- write hooks: SQLite first (shadow::insert_row / update_row / delete_row), then the slab
    def next_rowid(self):
        mx = self.conn.execute(f'SELECT MAX(rowid) FROM "main"."{self.table}_data"').fetchone()[0]
        return 1 if mx is None else mx + 1  # src/shadow.rs:888-900

    def insert(self, blob, rowid=None):
        rowid = self.next_rowid() if rowid is None else int(rowid)
        if blob is not None and len(blob) not in (0, self.row_bytes):
            # the vtab rejects wrong-length vectors on insert (src/vtab.rs:1474-1498)
            raise vec0.DimensionMismatch(
                f"Dimension mismatch: expected {self.dims}, got {len(blob)} bytes", self.dims, None)
        appended = self.slab is not None and rowid > self._max_rowid()
        self.conn.execute(f'INSERT INTO "main"."{self.table}_data" (rowid, vec{self.column_idx:02d}) VALUES (?, ?)', (rowid, blob))
        if self.slab is not None:
            self.slab.upsert(rowid, blob or b"")
            self._fingerprint = self._current_fingerprint()
        idx = getattr(self, "_hnsw", None)
        if idx is not None and not getattr(self, "_hnsw_stale", False) and appended:
            # Vec0Tab::insert -> insert_hnsw (src/vtab.rs:1409, src/hnsw/insert.rs:279-532): a row that arrives in rowid order
            # is inserted into the resident graph at once (vecgpu_hnsw_insert_appended); the shadow tables follow at the
            # next flush_hnsw_shadow() / rebuild_hnsw()
            if blob:
                if self._hnsw_slab is not self.slab:
                    idx.insert_appended(new_vectors=np.frombuffer(blob, dtype="<f4").reshape(1, -1), new_rowids=[rowid])
                else:
                    idx.insert_appended()
                self._hnsw_shadow_lag = True
        elif idx is not None and not getattr(self, "_hnsw_stale", False) and self.slab is not None:
            # an explicit rowid below the highest one: the row went between existing rows and every later row moved one
            # position up — the resident graph is renumbered on the device and the row inserted (vecgpu_hnsw_insert_at)
            try:
                if self._hnsw_slab is not self.slab:
                    idx.insert_at(rowid, None if not blob else np.frombuffer(blob, dtype="<f4"))
                else:
                    idx.insert_at(rowid)
                self._hnsw_shadow_lag = True
            except vec0.VecError:
                self._hnsw_stale = True  # (rows the index never saw, e.g. after writers that bypassed the hooks): rebuilt on demand
        else:
            self._hnsw_stale = idx is not None
        return rowid

    def _max_rowid(self):
        mx = self.conn.execute(f'SELECT MAX(rowid) FROM "main"."{self.table}_data"').fetchone()[0]
        return 0 if mx is None else mx

    def update(self, rowid, blob):
        self.conn.execute(f'UPDATE "main"."{self.table}_data" SET vec{self.column_idx:02d} = ? WHERE rowid = ?', (blob, int(rowid)))
        if self.slab is not None:
            self.slab.upsert(int(rowid), blob or b"")
            self._fingerprint = self._current_fingerprint()
        idx = getattr(self, "_hnsw", None)
        if idx is not None and not getattr(self, "_hnsw_stale", False):
            # Vec0Tab::update (src/vtab.rs:1860-1895): delete the node and its edges, insert the row again
            if self._hnsw_slab is not self.slab:
                idx.reinsert(int(rowid), None if not blob else np.frombuffer(blob, dtype="<f4"))
            else:
                idx.reinsert(int(rowid))
            self._hnsw_shadow_lag = True

    def delete(self, rowid):
        rowid = int(rowid)
        col = getattr(self, "_hnsw_column", None)
        if col is not None:
            # Vec0Tab::delete (src/vtab.rs:1340-1397): the node, its edges in both directions, num_nodes / hnsw_version, and a new
            # entry point (the highest remaining node) when the entry point itself goes
            nodes, edges, meta = (f'"{self.table}_{col}_hnsw_{x}"' for x in ("nodes", "edges", "meta"))
            self.conn.execute(f"DELETE FROM {nodes} WHERE rowid = ?", (rowid,))
            self.conn.execute(f"DELETE FROM {edges} WHERE from_rowid = ? OR to_rowid = ?", (rowid, rowid))
            row = self.conn.execute(f"SELECT entry_point_rowid, num_nodes FROM {meta} WHERE id = 1").fetchone()
            if row is not None:
                entry, num = row
                self.conn.execute(f"UPDATE {meta} SET num_nodes = ?, hnsw_version = hnsw_version + 1 WHERE id = 1", (max(0, num - 1),))
                if entry == rowid:
                    new = self.conn.execute(f"SELECT rowid, level FROM {nodes} ORDER BY level DESC LIMIT 1").fetchone()
                    self.conn.execute(f"UPDATE {meta} SET entry_point_rowid = ?, entry_point_level = ? WHERE id = 1", new or (-1, -1))
        self.conn.execute(f'DELETE FROM "main"."{self.table}_data" WHERE rowid = ?', (rowid,))
        if self.slab is not None:
            self.slab.delete(rowid)
            self._fingerprint = self._current_fingerprint()
        if getattr(self, "_hnsw", None) is not None and self._hnsw_slab is not self.slab:
            self._hnsw_slab.delete(rowid)  # the stored-representation slab carries its own tombstones
        if getattr(self, "_hnsw", None) is not None and not getattr(self, "_hnsw_stale", False):
            self._hnsw.reinsert(rowid)     # the row is deleted: its node and edges leave the resident graph too (vtab.rs:1340-1397)

    # ---- vec_rebuild_hnsw (src/sql_functions.rs:436-534) with the graph built on the GPU and written back in bulk
    def rebuild_hnsw(self, column, new_m=None, new_ef_construction=None):
        """Rebuild the HNSW index of `column` from the resident slab and replace the contents of the reference's shadow
        tables "{t}_{c}_hnsw_nodes / _hnsw_edges / _hnsw_meta" (src/shadow.rs:407-500; edges in insert_edges_batch shape,
        src/hnsw/storage.rs:346-383).  Stored vectors are normalised for cosine columns (src/hnsw/insert.rs:300-322).
        -> number of indexed vectors."""
        if self.is_stale():
            self.stage()
        meta = f'"{self.table}_{column}_hnsw_meta"'
        m, efc, seed, metric_s, quant = self.conn.execute(
            f"SELECT m, ef_construction, rng_seed, distance_metric, index_quantization FROM {meta} WHERE id = 1").fetchone()
        m = int(new_m) if new_m is not None else m
        efc = int(new_ef_construction) if new_ef_construction is not None else efc
        metric = vec0.DistanceMetric.from_str(metric_s)
        if getattr(self, "_hnsw", None) is not None:
            self._hnsw.close()
        # the graph lives over the STORED representation (src/hnsw/insert.rs:300-322): unit vectors for cosine columns, int8
        # with index_quantization=int8 — derived from the resident column slab on the device
        self._hnsw = vec0.HnswIndex.for_column(self.slab, metric, M=m, ef_construction=efc, seed=int(seed) & 0x7FFFFFFFFFFFFFFF,
                                               index_quantization=quant or "none")
        self._hnsw_slab = self._hnsw.slab
        self._hnsw_column = column
        self._hnsw_stale = False
        self._hnsw.rebuild()
        return self.flush_hnsw_shadow(m, efc)

    def flush_hnsw_shadow(self, m=None, efc=None):
        """Write the resident graph back into "{t}_{c}_hnsw_nodes / _hnsw_edges / _hnsw_meta" in bulk (after a rebuild, or after
        rows were inserted incrementally).  -> number of nodes."""
        column = self._hnsw_column
        meta = f'"{self.table}_{column}_hnsw_meta"'
        if m is None:
            m, efc = self.conn.execute(f"SELECT m, ef_construction FROM {meta} WHERE id = 1").fetchone()
        self._hnsw_shadow_lag = False
        rid, lv = self._hnsw.export_nodes()
        fr, to, elv, dist = self._hnsw.export_edges()
        entry, entry_level = self._hnsw.entry_point()
        if len(rid) and entry not in set(rid.tolist()):  # the entry point was deleted since: the highest remaining node (vtab.rs:1380-1394)
            top = int(np.argmax(lv))
            entry, entry_level = int(rid[top]), int(lv[top])
        nodes, edges = f'"{self.table}_{column}_hnsw_nodes"', f'"{self.table}_{column}_hnsw_edges"'
        self.conn.execute(f"DELETE FROM {edges}")
        self.conn.execute(f"DELETE FROM {nodes}")
        blobs = (self._hnsw_slab.get(int(r)) for r in rid)
        self.conn.executemany(f"INSERT INTO {nodes} (rowid, level, vector) VALUES (?, ?, ?)",
                              ((int(r), int(l), b) for r, l, b in zip(rid, lv, blobs)))
        self.conn.executemany(f"INSERT OR REPLACE INTO {edges} (from_rowid, to_rowid, level, distance) VALUES (?, ?, ?, ?)",
                              zip(fr.tolist(), to.tolist(), elv.tolist(), dist.tolist()))
        self.conn.execute(f"UPDATE {meta} SET m = ?, max_m0 = ?, ef_construction = ?, entry_point_rowid = ?, entry_point_level = ?, "
                          "num_nodes = ?, hnsw_version = hnsw_version + 1 WHERE id = 1", (m, 2 * m, efc, entry, entry_level, len(rid)))
        return len(rid)

    def hnsw_knn(self, query, k, ef_search=200, auto_rebuild=False):
        """search_hnsw over the resident index (src/hnsw/search.rs:267-335): -> [(rowid, distance in the column's metric)].
        Rows deleted since the rebuild are never returned; rows inserted in rowid order since are in the graph already
        (insert()) and updated rows have been re-inserted (update()).  A row inserted out of rowid order is inserted as well
        (the resident graph is renumbered on the device).  Only if that failed the index is stale: the call
        then refuses to answer unless auto_rebuild=True rebuilds it first."""
        if getattr(self, "_hnsw", None) is None:
            raise vec0.InvalidState("no HNSW index: call rebuild_hnsw() first")
        if getattr(self, "_hnsw_stale", False):
            if not auto_rebuild:
                raise vec0.InvalidState("rows were inserted or updated since the last rebuild_hnsw(): rebuild the index")
            self.rebuild_hnsw(self._hnsw_column)
        if isinstance(query, str):
            query = vec0.Vector.from_json(query, self.vec_type).as_bytes()
        q = np.frombuffer(query, dtype="u1").reshape(1, -1)
        r, d, c = self._hnsw.search(q, k, ef_search=ef_search)
        return [(int(r[0, j]), float(d[0, j])) for j in range(int(c[0]))]

    def compact(self, min_dead_fraction=0.0):
        """Drop the tombstones left by delete() (and rows with unreadable blobs) from the resident slab when they make up
        more than `min_dead_fraction` of it; SQLite itself is untouched.  -> rows removed."""
        if self.slab is None:
            return 0
        rows, live = self.slab.count()
        if rows == live or (rows - live) < min_dead_fraction * rows:
            return 0
        return self.slab.compact()

    def close(self):
        if getattr(self, "_hnsw", None) is not None:
            self._hnsw.close()  # also releases the stored-representation slab it owns
            self._hnsw = None
        if self.slab is not None:
            self.slab.close()
            self.slab = None
