"""oracle/hnsw_seq.c — the strictly sequential restatement of insert_hnsw / search_hnsw (src/hnsw/insert.rs:279-532,
src/hnsw/search.rs:267-543) — pinned on the CPU against what the reference's own tests hold for HNSW:
recall@10 >= 95 % (L2, tests/test_recall_accuracy.rs:6-135), >= 90 % (cosine, tests/test_recall_cosine.rs:15-125),
>= 90 % with the int8 index quantisation (tests/test_quantization_perf.rs:194-289), the 3-row cosine order
(src/vtab.rs:3246-3286), and against the independent Python restatement of the walk (oracle/hnsw_ref.py)."""
import numpy as np

import oracle as orc
from helpers import COSINE, F32, I8, L2
from oracle import hnsw_ref


def _recall(found, truth):
    return sum(len(set(f.tolist()) & set(t.tolist())) for f, t in zip(found, truth)) / truth.size


def test_ref_recall_l2_1000x128():
    n, dims = 1000, 128
    i = np.arange(n, dtype=np.int64)[:, None]
    j = np.arange(dims, dtype=np.int64)[None, :]
    v = ((i * 100 + j).astype("<f4") / np.float32(1000.0)).astype("<f4")
    q = np.full((1, dims), 0.5, dtype="<f4")
    h = orc.HnswSeq(F32, dims, L2, v, M=32, ef_construction=400, quirk=True)  # HnswParams defaults, src/hnsw/mod.rs:35-47
    h.build(orc.HnswSeq.levels(42, n, 32))
    r, d = h.search(q, 10, 200)
    er, _, _ = orc.knn(F32, dims, v, q, 10, L2)
    assert _recall(r + 1, er) >= 0.95
    assert np.all(np.diff(d[0]) >= 0)


def test_ref_recall_cosine_100x128_and_three_row_order():
    n, dims = 100, 128
    i = np.arange(n, dtype=np.int64)[:, None]
    j = np.arange(dims, dtype=np.int64)[None, :]
    v = ((((7 * i + 13 * j) % 100).astype("<f4")) / np.float32(100.0)).astype("<f4")
    vn = orc.normalize(v)  # cosine columns store normalised vectors and search with L2 (src/hnsw/mod.rs:129-146)
    h = orc.HnswSeq(F32, dims, L2, vn, M=32, ef_construction=400, quirk=True)
    h.build(orc.HnswSeq.levels(42, n, 32))
    r, d = h.search(orc.normalize(v[:10]), 10, 200)
    er, _, _ = orc.knn(F32, dims, v, v[:10], 10, COSINE)
    assert _recall(r + 1, er) >= 0.90
    assert np.all(r[:, 0] == np.arange(10))
    # src/vtab.rs:3246-3286: rows [1,0,0],[0.9,0.1,0],[0,1,0], query [1,0,0] -> rowids [1, 2] in order
    v3 = orc.normalize(np.array([[1, 0, 0], [0.9, 0.1, 0], [0, 1, 0]], dtype="<f4"))
    h3 = orc.HnswSeq(F32, 3, L2, v3, M=32, ef_construction=400, quirk=True)
    h3.build(orc.HnswSeq.levels(42, 3, 32))
    r3, _ = h3.search(v3[:1], 2, 200)
    assert list(r3[0] + 1) == [1, 2]


def test_ref_recall_int8_index_quantisation_5000x128():
    # tests/test_quantization_perf.rs:194-289: 5000 x 128 vectors, index_quantization=int8: stored vectors and the query are
    # quantize_int8_for_index(normalize(v)) (src/hnsw/insert.rs:300-322, src/hnsw/search.rs:285-302), scored with int8 L2;
    # recall@10 against the f32 ground truth must be >= 90 %
    n, dims, nq = 5000, 128, 20
    rng = np.random.default_rng(5)
    v = rng.standard_normal((n, dims)).astype("<f4")
    q = v[rng.choice(n, nq, replace=False)] + 0.05 * rng.standard_normal((nq, dims)).astype("<f4")
    stored = orc.quantize_int8_for_index(orc.normalize(v))
    h = orc.HnswSeq(I8, dims, L2, stored, M=32, ef_construction=400, quirk=True)
    h.build(orc.HnswSeq.levels(42, n, 32))
    r, _ = h.search(orc.quantize_int8_for_index(orc.normalize(q)), 10, 200)
    er, _, _ = orc.knn(F32, dims, v, q, 10, COSINE)
    assert _recall(r + 1, er) >= 0.90


def test_sequential_walk_equals_the_python_restatement():
    n, dims = 1500, 24
    v = orc.synth_rows(F32, 6, 1, n, dims, 1)
    q = orc.synth_rows(F32, 7, 1, 8, dims, 1)
    for quirk in (False, True):
        h = orc.HnswSeq(F32, dims, L2, v, M=8, ef_construction=40, quirk=quirk)
        h.build(orc.HnswSeq.levels(3, n, 8))
        fr, to, lv, ds = h.export()
        info = h.info()
        assert info["nodes"] == n
        nbrs = hnsw_ref.adjacency_from_edges(fr, to, lv)
        r, d = h.search(q, 10, 32)
        for qi in range(len(q)):
            want = hnsw_ref.search_hnsw(lambda node: float(orc.distance(F32, q[qi], v[node], L2)), nbrs, info["entry"], info["entry_level"], 10, 32)
            assert [w[0] for w in want] == [int(x) for x in r[qi]]
            assert np.array_equal(np.array([w[1] for w in want], dtype="<f4").view("<u4"), d[qi].view("<u4"))
        # graph invariants of insert.rs: degree caps, no self loops, stored distance == distance(from, to)
        levels = orc.HnswSeq.levels(3, n, 8)
        deg = {}
        for f, t, l in zip(fr, to, lv):
            deg[(int(f), int(l))] = deg.get((int(f), int(l)), 0) + 1
            assert f != t
        assert max(c for (f, l), c in deg.items() if l == 0) <= 16 and max([c for (f, l), c in deg.items() if l > 0] or [0]) <= 8
        pick = np.random.default_rng(1).choice(len(fr), 200, replace=False)
        for e in pick:
            assert np.float32(orc.distance(F32, v[fr[e]], v[to[e]], L2)).view("<u4") == ds[e].view("<u4")
        if not quirk:  # without the upper-layer quirk no node has edges above its own level
            assert all(l <= levels[f] for f, l in zip(fr, lv))
        h.close()
