"""The product's HNSW build against the strictly sequential restatement of the reference (oracle/hnsw_seq.c):
with insert batches of ONE the GPU build IS the sequential algorithm, and must produce the same graph edge for edge
(same neighbour sets, same stored distance bits) and the same query results; with the default large batches the
members of a batch do not see each other — the deviation is measured as recall, not hidden."""
import numpy as np
import pytest

from helpers import F32, I8, L2

pytestmark = pytest.mark.gpu


def _edge_map(fr, to, lv, ds, node_levels=None):
    m = {}
    for f, t, l, d in zip(fr.tolist(), to.tolist(), lv.tolist(), ds.view("<u4").tolist()):
        if node_levels is not None and (l > node_levels[f] or l > node_levels[t]):
            continue
        m.setdefault((f, l), {})[t] = d
    return m


@pytest.mark.parametrize("elem,dims,n,M,efc", [(F32, 32, 3000, 16, 100), (F32, 96, 1500, 8, 40)])
def test_batch_of_one_build_equals_the_sequential_reference_graph(vg, orc, gpu, elem, dims, n, M, efc):
    v = orc.synth_rows(elem, 6, 1, n, dims, 1)
    q = orc.synth_rows(elem, 7, 1, 12, dims, 1)
    with vg.Slab(elem, dims) as s:
        s.load(v)
        idx = vg.HnswIndex(s, L2, M=M, ef_construction=efc, seed=1)
        idx.rebuild(batch=1)
        rid, lv_nodes = idx.export_nodes()
        assert np.array_equal(rid, np.arange(1, n + 1))
        want_levels = orc.HnswSeq.levels(1, n, M)
        assert np.array_equal(lv_nodes.astype("i1"), want_levels), "the oracle's level sequence is the product's"
        h = orc.HnswSeq(elem, dims, L2, v, M=M, ef_construction=efc, quirk=False)
        h.build(want_levels)
        fr, to, lv, ds = idx.export_edges()
        got = _edge_map(fr - 1, to - 1, lv, ds)
        ofr, oto, olv, ods = h.export()
        want = _edge_map(ofr, oto, olv, ods)
        assert got.keys() == want.keys()
        for key in want:
            assert got[key] == want[key], f"adjacency of node {key[0]} at level {key[1]} differs from the sequential build"
        info = h.info()
        assert idx.entry_point() == (info["entry"] + 1, info["entry_level"])
        # the same walk step for step: the expansion batch-size histogram (search.rs:443-455 buckets) and the number of
        # distances computed by the build are identical too
        assert idx.batch_histogram() == info["batch_hist"]
        assert idx.stats()["distances_scored"] == info["distances"]
        # same walk results (ids and distance bits) on the two graphs
        r, d, c = idx.search(q, 10, ef_search=64)
        orr, od = h.search(q, 10, 64)
        assert np.array_equal(r, orr + 1) and np.array_equal(d.view("<u4"), od.view("<u4"))
        # the reference's upper-layer quirk (old entry point linked on layers above its own level) changes a handful of edges
        hq = orc.HnswSeq(elem, dims, L2, v, M=M, ef_construction=efc, quirk=True)
        hq.build(want_levels)
        qfr, qto, qlv, qds = hq.export()
        extra = sum(1 for f, l in zip(qfr.tolist(), qlv.tolist()) if l > want_levels[f])
        rq, _ = hq.search(q, 10, 64)
        print(f"quirk edges above a node's own level: {extra}; queries with a different top-10: {int(np.sum(np.any(rq != orr, axis=1)))}/12")
        idx.close()
        h.close()
        hq.close()


def test_batched_build_recall_tracks_the_sequential_build(vg, orc, gpu):
    """20 k x 64: recall@10 of the default batched build is within 2 points of the sequential reference procedure."""
    n, dims, nq, M, efc = 20_000, 64, 200, 16, 200
    v = orc.synth_rows(F32, 6, 1, n, dims, 1)
    q = orc.synth_rows(F32, 67, 1, nq, dims, 1)
    er, _, _ = orc.knn_select(F32, dims, v, q, 10, L2)
    h = orc.HnswSeq(F32, dims, L2, v, M=M, ef_construction=efc)
    h.build(orc.HnswSeq.levels(1, n, M))
    rs, _ = h.search(q, 10, 100)
    rec_seq = np.mean([len(set(a.tolist()) & set((b - 1).tolist())) / 10 for a, b in zip(rs, er)])
    with vg.Slab(F32, dims) as s:
        s.load(v)
        idx = vg.HnswIndex(s, L2, M=M, ef_construction=efc, seed=1)
        idx.rebuild()
        r, _, _ = idx.search(q, 10, ef_search=100)
        rec_gpu = np.mean([len(set(a.tolist()) & set(b.tolist())) / 10 for a, b in zip(r, er)])
        idx.close()
    print(f"recall@10 ef=100: sequential {rec_seq:.3f}, batched GPU build {rec_gpu:.3f}")
    assert rec_gpu >= rec_seq - 0.02
    h.close()


def test_int8_index_quantised_column_equals_the_sequential_reference(vg, orc, gpu):
    """HnswIndex.for_column(cosine float32 column, index_quantization=int8): stored slab == quantize_int8_for_index(normalize(v))
    bit for bit, and — built with insert batches of one — the graph and the query results equal the sequential oracle over
    the same int8 vectors (massively tied integer distances included: edge SETS are compared, order within a list is not)."""
    n, dims, M, efc = 2000, 48, 8, 40
    v = orc.synth_rows(F32, 6, 1, n, dims, 1)
    q = orc.synth_rows(F32, 7, 1, 10, dims, 1)
    stored = orc.quantize_int8_for_index(orc.normalize(v))
    with vg.Slab(F32, dims) as col:
        col.load(v)
        idx = vg.HnswIndex.for_column(col, vg.DistanceMetric.Cosine, M=M, ef_construction=efc, seed=1, index_quantization="int8")
        assert idx.slab.vec_type == vg.VectorType.Int8
        for rid in (1, 2, 777, n):
            assert idx.slab.get(rid) == stored[rid - 1].tobytes()
        idx.rebuild(batch=1)
        r, d, c = idx.search(q, 10, ef_search=64)
        h = orc.HnswSeq(I8, dims, L2, stored, M=M, ef_construction=efc)
        h.build(orc.HnswSeq.levels(1, n, M))
        orr, od = h.search(orc.quantize_int8_for_index(orc.normalize(q)), 10, 64)
        # integer distances tie constantly and the reference leaves tie order to heap internals: compare distances exactly,
        # ids as sets per distance value
        want_d = ((od * od) / np.float32(2.0)).astype("<f4")  # convert_distance_for_output (src/hnsw/mod.rs:139-146)
        assert np.array_equal(d.view("<u4"), want_d.view("<u4"))
        agree = np.mean([len(set(a.tolist()) & set((b + 1).tolist())) / 10 for a, b in zip(r, orr)])
        assert agree >= 0.9
        idx.close()
        h.close()


@pytest.mark.parametrize("elem,dims", [(F32, 384), (I8, 128), (I8, 16)])
def test_cta_per_query_walk_equals_the_warp_walk(vg, orc, gpu, elem, dims, monkeypatch):
    """Few queries take hnsw_search_cta_kernel (one CTA per query: all fresh neighbours scored at once, visited set in
    shared memory, the neighbours of an expansion admitted in one merge); it must return exactly what the
    one-warp-per-query kernel returns, ids and distance bits — also where distances tie constantly (int8[16]: the merge
    hands a batch whose cut falls inside a tie back to the one-by-one admission, and the helper warps' look-ahead
    mispredicts there, so its visited entries are taken back) and with each of the three mechanisms switched off."""
    n, M, efc = 30_000, 16, 100
    v = orc.synth_rows(elem, 6, 1, n, dims, 1 if elem == F32 else 0)
    q = orc.synth_rows(elem, 7, 1, 40, dims, 1 if elem == F32 else 0)
    with vg.Slab(elem, dims) as s:
        s.load(v)
        idx = vg.HnswIndex(s, L2, M=M, ef_construction=efc, seed=3)
        idx.rebuild()
        for ef, k in ((10, 10), (64, 10), (200, 10), (400, 50)):
            monkeypatch.setenv("VECGPU_HNSW_CTA_MAX_NQ", "0")
            wr, wd, wc = idx.search(q, k, ef_search=ef)      # warp kernel
            monkeypatch.setenv("VECGPU_HNSW_CTA_MAX_NQ", "64")
            cr, cd, cc = idx.search(q, k, ef_search=ef)      # CTA kernel (40 queries)
            c1 = [idx.search(q[i], k, ef_search=ef) for i in range(3)]  # and one query at a time
            variants = []
            for knob in ("VECGPU_HNSW_BATCH_ADMIT", "VECGPU_HNSW_PREFETCH", "VECGPU_HNSW_SPEC_ROWS"):
                # one sorted insert per admitted neighbour / no look-ahead by the helper warps / look-ahead into L2 only
                monkeypatch.setenv(knob, "0")
                variants.append((knob, idx.search(q, k, ef_search=ef)))
                monkeypatch.delenv(knob)
            monkeypatch.delenv("VECGPU_HNSW_CTA_MAX_NQ")
            assert np.array_equal(wr, cr) and np.array_equal(wd.view("<u4"), cd.view("<u4")) and np.array_equal(wc, cc)
            for knob, (sr, sd, sc) in variants:
                assert np.array_equal(wr, sr) and np.array_equal(wd.view("<u4"), sd.view("<u4")) and np.array_equal(wc, sc), knob
            for i in range(3):
                assert np.array_equal(c1[i][0][0], wr[i]) and np.array_equal(c1[i][1][0].view("<u4"), wd[i].view("<u4"))
        assert idx.device_stats()["fallbacks"] == 0
        idx.close()


@pytest.mark.parametrize("elem,dims,metric,M", [(2, 256, 3, 16), (F32, 96, 2, 32), (I8, 64, 1, 40), (F32, 8, L2, 5)])
def test_cta_per_query_walk_other_metrics_and_degrees(vg, orc, gpu, elem, dims, metric, M, monkeypatch):
    """The one-CTA-per-query walk on bit / Hamming, cosine and L1 columns and with adjacency lists longer than one 32-entry
    batch (M = 32, 40: max_m0 = 64, 80 -> several merges per expansion, look-ahead without the second row buffer when more
    than 32 neighbours are new) and very short ones (M = 5): identical to the one-warp walk, single queries included."""
    n, efc = 12_000, 80
    kind = 1 if elem == F32 else 0
    v = orc.synth_rows(elem, 16, 1, n, dims, kind)
    q = orc.synth_rows(elem, 17, 1, 24, dims, kind)
    with vg.Slab(elem, dims) as s:
        s.load(v)
        idx = vg.HnswIndex(s, metric, M=M, ef_construction=efc, seed=5)
        idx.rebuild()
        for ef, k in ((1, 1), (40, 10), (300, 20), (512, 100)):
            monkeypatch.setenv("VECGPU_HNSW_CTA_MAX_NQ", "0")
            wr, wd, wc = idx.search(q, k, ef_search=ef)
            monkeypatch.setenv("VECGPU_HNSW_CTA_MAX_NQ", "64")
            cr, cd, cc = idx.search(q, k, ef_search=ef)
            one = [idx.search(q[i], k, ef_search=ef) for i in range(2)]
            monkeypatch.delenv("VECGPU_HNSW_CTA_MAX_NQ")
            assert np.array_equal(wr, cr) and np.array_equal(wd.view("<u4"), cd.view("<u4")) and np.array_equal(wc, cc), (ef, k)
            for i in range(2):
                assert np.array_equal(one[i][0][0], wr[i]) and np.array_equal(one[i][1][0].view("<u4"), wd[i].view("<u4"))
        assert idx.device_stats()["fallbacks"] == 0
        idx.close()


def test_insert_appended_with_batch_one_continues_the_sequential_build(vg, orc, gpu):
    """vecgpu_hnsw_insert_appended: rows appended to the slab after a build are indexed by continuing the rebuild's insertion
    loop.  With batches of one, building the first 1800 rows and then inserting 700 more (in two appends) gives the strictly
    sequential reference graph of all 2500 rows, edge for edge and distance bit for distance bit (insert.rs:279-532)."""
    elem, dims, n, M, efc = F32, 32, 2500, 12, 60
    v = orc.synth_rows(elem, 6, 1, n, dims, 1)
    q = orc.synth_rows(elem, 7, 1, 12, dims, 1)
    with vg.Slab(elem, dims) as s:
        s.load(v[:1800])
        idx = vg.HnswIndex(s, L2, M=M, ef_construction=efc, seed=1)
        assert idx.insert_appended(batch=1) == 1800      # an empty index is simply built
        assert idx.insert_appended(batch=1) == 0         # nothing new
        s.append(v[1800:2200])
        assert idx.insert_appended(batch=1) == 400
        for i in range(2200, n):                         # the reference's shape: one INSERT at a time
            s.upsert(i + 1, v[i].tobytes())
        assert idx.insert_appended(batch=1) == n - 2200
        want_levels = orc.HnswSeq.levels(1, n, M)
        h = orc.HnswSeq(elem, dims, L2, v, M=M, ef_construction=efc, quirk=False)
        h.build(want_levels)
        fr, to, lv, ds = idx.export_edges()
        got = _edge_map(fr - 1, to - 1, lv, ds)
        ofr, oto, olv, ods = h.export()
        want = _edge_map(ofr, oto, olv, ods)
        assert got.keys() == want.keys()
        for key in want:
            assert got[key] == want[key], f"adjacency of node {key[0]} at level {key[1]} differs from the sequential build"
        info = h.info()
        assert idx.entry_point() == (info["entry"] + 1, info["entry_level"]) and idx.stats()["nodes"] == n
        r, d, c = idx.search(q, 10, ef_search=64)
        orr, od = h.search(q, 10, 64)
        assert np.array_equal(r, orr + 1) and np.array_equal(d.view("<u4"), od.view("<u4"))
        idx.close()
        h.close()


def test_insert_appended_batched_and_its_failure_modes(vg, orc, gpu):
    """Default batching: 20 000 rows built, 10 000 appended and inserted — every row is a node, recall against the exact scan is
    that of a full rebuild; deleted rows are never returned; an out-of-order insert moves row positions, after which the
    index refuses to be extended or searched until it is rebuilt."""
    dims, n0, n = 48, 20_000, 30_000
    v = orc.synth_rows(F32, 6, 1, n, dims, 1)
    q = orc.synth_rows(F32, 7, 1, 200, dims, 1)
    rowids = np.arange(1, n + 1, dtype="<i8") * 2           # even rowids: room for an out-of-order insert
    with vg.Slab(F32, dims) as s:
        s.load(v[:n0], rowids[:n0])
        idx = vg.HnswIndex(s, L2, M=16, ef_construction=100, seed=3)
        assert idx.rebuild() == n0
        s.append(v[n0:], rowids[n0:])
        s.delete(int(rowids[n0 + 5]))                        # a new row deleted before it is indexed: never a node
        assert idx.insert_appended() == n - n0 - 1
        assert idx.stats()["nodes"] == n - 1
        er, _, _ = s.knn(q, 10, L2)
        r, d, c = idx.search(q, 10, ef_search=100)
        rec_inc = np.mean([len(set(a.tolist()) & set(b.tolist())) / 10 for a, b in zip(r, er)])
        assert not np.isin(r, [rowids[n0 + 5]]).any()
        assert (r > rowids[n0 - 1]).any(), "appended rows are found"
        idx2 = vg.HnswIndex(s, L2, M=16, ef_construction=100, seed=3)
        idx2.rebuild()
        r2, _, _ = idx2.search(q, 10, ef_search=100)
        rec_full = np.mean([len(set(a.tolist()) & set(b.tolist())) / 10 for a, b in zip(r2, er)])
        assert rec_inc >= rec_full - 0.03 and rec_inc >= 0.8, (rec_inc, rec_full)
        idx2.close()
        s.upsert(7, v[3].tobytes())                          # rowid 7 lies between 6 and 8: row positions move
        with pytest.raises(vg.InvalidState):
            idx.insert_appended()
        with pytest.raises(vg.InvalidState):
            idx.search(q[:1], 10)
        assert idx.rebuild() == n                            # n - 1 live rows + the new one
        idx.close()


def test_insert_appended_on_a_cosine_int8_column(vg, orc, gpu):
    """An index made by for_column() owns the slab of stored vectors: new rows are passed raw, normalised and quantised like the
    rest (insert.rs:300-322); a zero vector cannot be normalised and stays out of the index."""
    dims, n0, n = 64, 3000, 4000
    v = orc.synth_rows(F32, 6, 1, n, dims, 1)
    v[3500] = 0
    q = orc.synth_rows(F32, 7, 1, 50, dims, 1)
    with vg.Slab(F32, dims) as col:
        col.load(v[:n0])
        idx = vg.HnswIndex.for_column(col, vg.DistanceMetric.Cosine, M=16, ef_construction=100, index_quantization="int8")
        assert idx.rebuild() == n0
        col.append(v[n0:])
        assert idx.insert_appended(new_vectors=v[n0:]) == n - n0 - 1
        er, ed, _ = col.knn(q, 10, vg.DistanceMetric.Cosine)
        r, d, c = idx.search(q, 10, ef_search=200)
        rec = np.mean([len(set(a.tolist()) & set(b.tolist())) / 10 for a, b in zip(r, er)])
        assert rec >= 0.9 and (r > n0).any() and not (r == 3501).any()   # the reference's bar for int8 indexes (>= 90 %)
        idx.close()


def test_reinsert_equals_the_sequential_reference_update(vg, orc, gpu):
    """vecgpu_hnsw_reinsert = Vec0Tab::update of an indexed column (src/vtab.rs:1860-1895): delete the node and its edges in
    both directions, insert the row again with its new vector.  After a batch-of-one build and a series of updates — ordinary
    nodes, the entry point itself, a row that becomes empty — the graph equals the sequential restatement's after the same
    operations, edge for edge and distance bit for distance bit, and searches agree."""
    elem, dims, n, M, efc = F32, 24, 1500, 8, 50
    v = orc.synth_rows(elem, 6, 1, n, dims, 1).copy()
    new = orc.synth_rows(elem, 9, 1, 8, dims, 1)
    q = np.concatenate([orc.synth_rows(elem, 7, 1, 8, dims, 1), new])
    levels = orc.HnswSeq.levels(1, n, M)
    with vg.Slab(elem, dims) as s:
        s.load(v)
        idx = vg.HnswIndex(s, L2, M=M, ef_construction=efc, seed=1)
        idx.rebuild(batch=1)
        h = orc.HnswSeq(elem, dims, L2, v, M=M, ef_construction=efc, quirk=False)   # borrows v: edits below are seen
        h.build(levels)
        ep = idx.entry_point()[0]
        targets = [17, 900, ep, 3, 1499, ep]   # the entry point twice (it changes hands the first time)
        for j, rid in enumerate(targets):
            if j == 5:
                rid = idx.entry_point()[0]
            v[rid - 1] = new[j]
            h._v[rid - 1] = new[j]
            s.upsert(rid, new[j].tobytes())
            idx.reinsert(rid)
            h.reinsert(rid - 1, levels[rid - 1])
        s.upsert(40, b"")                          # the row becomes empty: it only leaves the graph
        idx.reinsert(40)
        h.reinsert(39, levels[39], insert_again=False)
        fr, to, lv, ds = idx.export_edges()
        got = _edge_map(fr - 1, to - 1, lv, ds)
        ofr, oto, olv, ods = h.export()
        want = _edge_map(ofr, oto, olv, ods)
        assert got.keys() == want.keys()
        for key in want:
            assert got[key] == want[key], f"adjacency of node {key[0]} at level {key[1]} differs from the sequential update"
        info = h.info()
        assert idx.entry_point() == (info["entry"] + 1, info["entry_level"]) and idx.stats()["nodes"] == n - 1
        r, d, c = idx.search(q, 5, ef_search=64)
        orr, od = h.search(q, 5, 64)
        assert np.array_equal(r, orr + 1) and np.array_equal(d.view("<u4"), od.view("<u4"))
        assert 40 not in r
        for j in (0, 1, 3, 4):                     # the updated rows are found by their new vectors
            assert r[8 + j, 0] == targets[j] and d[8 + j, 0] == 0
        with pytest.raises(vg.InvalidParameter):
            idx.reinsert(n + 5)                    # not in the slab
        idx.close()
        h.close()


@pytest.mark.parametrize("resident", [True, False])
def test_insert_at_equals_the_sequential_build_in_insertion_order(vg, orc, gpu, monkeypatch, resident):
    """vecgpu_hnsw_insert_at = Vec0Tab::insert with an explicit rowid below the highest one (src/vtab.rs:1409-1682 ->
    insert_hnsw, src/hnsw/insert.rs:279-532).  The slab keeps rows in rowid order, so the new row lands between existing rows
    and every later row (= node id) moves up by one; the resident graph is renumbered on the device and the row inserted.
    After a batch-of-one build over even rowids and a series of odd-rowid inserts — first row, last gap, middle, next to the
    entry point, an empty blob — interleaved with appended rows, the graph equals the sequential restatement's built over
    the same vectors in INSERTION order, edge for edge and distance bit for distance bit (node ids compared as rowids).
    resident=False: lockstep mode (VECGPU_HNSW_DEVICE=0), where the same renumbering happens on the host lists."""
    if not resident:
        monkeypatch.setenv("VECGPU_HNSW_DEVICE", "0")
    elem, dims, n0, M, efc = F32, 24, 1500, 8, 50
    extra = 40
    v = orc.synth_rows(elem, 6, 1, n0 + extra, dims, 1)
    q = np.concatenate([orc.synth_rows(elem, 7, 1, 12, dims, 1), v[n0:n0 + 8]])
    rowids0 = np.arange(1, n0 + 1, dtype="<i8") * 2
    with vg.Slab(elem, dims) as s:
        s.load(v[:n0], rowids0)
        idx = vg.HnswIndex(s, L2, M=M, ef_construction=efc, seed=1)
        idx.rebuild(batch=1)
        ep = idx.entry_point()[0]
        order = list(rowids0)                     # rowid of the k-th inserted row
        rng = np.random.default_rng(3)
        odd = [1, 2 * n0 - 1, 1501, ep - 1, ep + 1] + [int(x) for x in rng.choice(np.arange(3, 2 * n0 - 2, 2), 27, replace=False)]
        odd = [x for x in dict.fromkeys(odd) if 0 < x < 2 * n0][:extra - 8]
        spare = [x for x in range(3, 2 * n0, 2) if x not in odd][:3]   # odd rowids left free
        k = n0
        for j, rid in enumerate(odd):
            s.upsert(rid, v[k].tobytes())
            idx.insert_at(rid)
            order.append(rid)
            k += 1
            if j % 8 == 7:                         # an ordinary appended row in between: the level sequence goes on
                top = 2 * n0 + 2 * (j // 8 + 1)
                s.upsert(top, v[k].tobytes())
                assert idx.insert_appended(batch=1) == 1
                order.append(top)
                k += 1
        n = k
        order = np.array(order, dtype="<i8")
        assert idx.stats()["nodes"] == n
        s.upsert(spare[0], b"")                  # an empty blob out of order: a row, not a node
        idx.insert_at(spare[0])
        assert idx.stats()["nodes"] == n
        h = orc.HnswSeq(elem, dims, L2, v[:n], M=M, ef_construction=efc, quirk=False)
        h.build(orc.HnswSeq.levels(1, n, M))
        fr, to, lv, ds = idx.export_edges()
        pos_of = {int(r): i for i, r in enumerate(order)}
        got = _edge_map(np.array([pos_of[int(x)] for x in fr]), np.array([pos_of[int(x)] for x in to]), lv, ds)
        ofr, oto, olv, ods = h.export()
        want = _edge_map(ofr, oto, olv, ods)
        assert got.keys() == want.keys()
        for key in want:
            assert got[key] == want[key], f"adjacency of insertion {key[0]} at level {key[1]} differs from the sequential build"
        info = h.info()
        assert idx.entry_point() == (int(order[info["entry"]]), info["entry_level"])
        r, d, c = idx.search(q, 5, ef_search=64)
        orr, od = h.search(q, 5, 64)
        assert np.array_equal(r, order[orr]) and np.array_equal(d.view("<u4"), od.view("<u4"))
        # the rows inserted out of order sit where the exact scan finds them, and the walk (M = 8: not every one) finds them too
        er, ed, _ = s.knn(q[12:], 1, L2)
        assert np.array_equal(er[:, 0], order[n0:n0 + 8]) and not ed.any()
        assert (r[12:, 0] == er[:, 0]).sum() >= 5
        # two out-of-order rows without a call in between: the index cannot follow, and says so
        with pytest.raises(vg.InvalidParameter):
            idx.insert_at(spare[1])                # not in the slab
        s.upsert(spare[1], v[0].tobytes())
        s.upsert(spare[2], v[1].tobytes())
        with pytest.raises(vg.InvalidState):
            idx.insert_at(spare[2])
        assert idx.rebuild() == n + 2
        idx.close()
        h.close()


@pytest.mark.parametrize("seed", [1, 2, 3, 4, 5])
def test_random_write_traffic_equals_the_sequential_reference(vg, orc, gpu, seed):
    """Differential test of every write hook on the resident graph, interleaved at random: rows appended (insert_appended), rows
    inserted between existing rowids (insert_at), rows updated (reinsert), rows deleted (delete + reinsert) and deleted rowids
    used again (insert_at on the tombstoned row).  The sequential restatement performs the same operations in the same order
    (insert_hnsw / the delete-and-insert of Vec0Tab::update / Vec0Tab::delete); its node ids are insertion numbers, the
    product's are row positions that keep moving — the graphs must agree edge for edge, distance bit for distance bit, after
    every 50 operations, and so must the walks."""
    rng = np.random.default_rng(seed)
    elem, dims, n0, M, efc, n_ops = F32, 16, 600, 6, 40, 200
    total = n0 + n_ops
    vecs = orc.synth_rows(elem, 20 + seed, 1, total, dims, 1).copy()
    fresh = orc.synth_rows(elem, 40 + seed, 1, n_ops, dims, 1)
    q = orc.synth_rows(elem, 7, 1, 16, dims, 1)
    levels = orc.HnswSeq.levels(seed, total, M)
    h = orc.HnswSeq(elem, dims, L2, vecs, M=M, ef_construction=efc, quirk=False)
    rowid_of = [10 * (i + 1) for i in range(n0)]          # insertion number -> rowid
    node_of = {r: i for i, r in enumerate(rowid_of)}       # rowid -> insertion number
    live, dead = set(rowid_of), []
    with vg.Slab(elem, dims) as s:
        s.load(vecs[:n0], np.array(rowid_of, dtype="<i8"))
        idx = vg.HnswIndex(s, L2, M=M, ef_construction=efc, seed=seed)
        idx.rebuild(batch=1)
        for i in range(n0):
            h.insert(i, levels[i])

        def check():
            fr, to, lv, ds = idx.export_edges()
            got = _edge_map(np.array([node_of[int(x)] for x in fr]), np.array([node_of[int(x)] for x in to]), lv, ds)
            ofr, oto, olv, ods = h.export()
            want = _edge_map(ofr, oto, olv, ods)
            assert got.keys() == want.keys()
            for key in want:
                assert got[key] == want[key], f"adjacency of insertion {key[0]} (rowid {rowid_of[key[0]]}) at level {key[1]} differs"
            info = h.info()
            assert idx.entry_point() == (rowid_of[info["entry"]], info["entry_level"]) and idx.stats()["nodes"] == info["nodes"] == len(live)
            r, d, c = idx.search(q, 5, ef_search=48)
            orr, od = h.search(q, 5, 48)
            assert np.array_equal(r, np.array(rowid_of, dtype="<i8")[orr]) and np.array_equal(d.view("<u4"), od.view("<u4"))

        check()
        counts = dict(append=0, between=0, update=0, delete=0, reuse=0)
        for op in range(n_ops):
            entry = idx.entry_point()[0]
            kind = rng.choice(["append", "between", "update", "delete", "reuse"], p=[0.25, 0.3, 0.2, 0.15, 0.1])
            if kind == "reuse" and not dead:
                kind = "between"
            if kind in ("append", "between"):
                k = len(rowid_of)
                top = max(rowid_of)
                if kind == "append":
                    rid = top + 10
                else:
                    rid = int(rng.integers(1, top))
                    while rid in node_of:
                        rid = int(rng.integers(1, top))
                s.upsert(rid, vecs[k].tobytes())
                if kind == "append":
                    assert idx.insert_appended(batch=1) == 1
                else:
                    idx.insert_at(rid)
                rowid_of.append(rid)
                node_of[rid] = k
                live.add(rid)
                h.insert(k, levels[k])
            elif kind == "update":
                rid = int(rng.choice(sorted(live - {entry})))
                h._v[node_of[rid]] = fresh[op]
                s.upsert(rid, fresh[op].tobytes())
                idx.reinsert(rid)
                h.reinsert(node_of[rid], levels[node_of[rid]])
            elif kind == "delete":
                rid = int(rng.choice(sorted(live - {entry})))
                s.delete(rid)
                idx.reinsert(rid)
                h.reinsert(node_of[rid], levels[node_of[rid]], insert_again=False)
                live.discard(rid)
                dead.append(rid)
            else:
                rid = dead.pop(int(rng.integers(len(dead))))
                h._v[node_of[rid]] = fresh[op]
                s.upsert(rid, fresh[op].tobytes())
                idx.insert_at(rid)                 # the tombstoned row is still in the slab: nothing moves
                h.insert(node_of[rid], levels[node_of[rid]])
                live.add(rid)
            counts[kind] += 1
            if op % 50 == 49:
                check()
        assert all(counts.values()), counts
        idx.close()
    h.close()
