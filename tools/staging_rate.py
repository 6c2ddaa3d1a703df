#!/usr/bin/env python
"""How fast does a vec0 table travel from a SQLite FILE into the HBM slab?  Builds "{t}_data" with N rows of
f32[768] in a file-backed database, then times sqlite_shim.Vec0Table.stage() = SELECT rowid, vec00 ... ORDER BY rowid
(src/shadow.rs:853-868) + assembly + vecgpu_slab_load (two pinned staging buffers, asynchronous copies), and the load alone
from an array already in RAM.   python tools/staging_rate.py [rows]"""
import json
import os
import sqlite3
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import importlib  # noqa: E402

import sqlite_vec_hnsw_b200 as vg  # noqa: E402

sh = importlib.import_module(vg.Slab.__module__.rsplit(".", 1)[0] + ".sqlite_shim")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
dims = 768
rng = np.random.default_rng(1)
v = rng.standard_normal((n, dims)).astype("<f4")
with tempfile.TemporaryDirectory() as d:
    path = os.path.join(d, "stage.db")
    conn = sqlite3.connect(path)
    sh.create_shadow_tables(conn, "docs", 1, [])
    t0 = time.perf_counter()
    conn.executemany('INSERT INTO "docs_data" (rowid, vec00) VALUES (?, ?)', ((i + 1, v[i].tobytes()) for i in range(n)))
    conn.commit()
    t_ins = time.perf_counter() - t0
    conn.close()
    conn = sqlite3.connect(path)
    t = sh.Vec0Table(conn, "docs", 0, dims)
    t0 = time.perf_counter()
    t.stage()
    t_stage = time.perf_counter() - t0
    t0 = time.perf_counter()
    rowids, vec, skip = sh.read_column(conn, "docs", 0, dims * 4)
    t_read = time.perf_counter() - t0
    s2 = vg.Slab(0, dims)
    s2.load(v)
    t0 = time.perf_counter()
    s2.load(v)
    t_load = time.perf_counter() - t0
    gb = n * dims * 4 / 1e9
    print(json.dumps({"rows": n, "dims": dims, "gigabytes": gb, "sqlite_insert_s": t_ins, "stage_s": t_stage, "stage_gb_per_s": gb / t_stage,
                      "of_which_sqlite_read_and_assembly_s": t_read, "slab_load_from_ram_s": t_load, "slab_load_gb_per_s": gb / t_load,
                      "file_bytes": os.path.getsize(path)}))
    t.close()
    s2.close()
