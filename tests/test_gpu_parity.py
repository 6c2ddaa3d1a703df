"""GPU parity: libvecgpu.so (through the C ABI / host mirror) against the CPU
oracle and the committed golden fixtures.

Bar (BASELINE.json north_star): rowids bit-exact; int8-L2/L1 + Hamming distances
bit-exact; float distances within 1e-5 relative (in fact bit-exact against the
oracle because both follow the canonical accumulation order, asserted below).
"""
import json
import os

import numpy as np
import pytest

from helpers import BIT, COSINE, F32, HAMMING, I8, L1, L2, NP, PAIR_IDS, PAIRS, random_rows, rel_close, same_bits

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
REL_TOL = 1e-5  # north_star tolerance for floating-point distances


def bits(a):
    return np.ascontiguousarray(a, dtype="<f4").view("<u4")


def check_knn(vg, orc, elem, dims, vectors, queries, k, metric, rowids=None, skip=None, slab=None):
    own = slab is None
    if own:
        slab = vg.Slab(elem, dims)
        slab.load(vectors, rowids)
        if skip is not None:
            ids = np.arange(1, len(vectors) + 1) if rowids is None else rowids
            for i in np.flatnonzero(skip):
                slab.delete(int(ids[i]))
    r, d, c = slab.knn(queries, k, metric)
    er, ed, ec = orc.knn(elem, dims, vectors, queries, k, metric, rowids=rowids, skip=skip)
    assert np.array_equal(c, ec)
    assert np.array_equal(r, er), "rowids must be bit-exact"
    assert rel_close(d, ed, REL_TOL)
    assert same_bits(d, ed), "canonical order makes distances bit-exact vs the oracle"
    if own:
        slab.close()
    return r, d, c


# ------------------------------------------------------------------ reference tests, replayed on the GPU
def test_ref_distance_known_answers(vg, gpu):
    V, M = vg.Vector, vg.DistanceMetric
    assert abs(vg.distance(V.from_f32([1, 2, 3]), V.from_f32([4, 5, 6]), M.L2) - 5.196) < 0.01  # scalar.rs:120-130
    assert abs(vg.distance(V.from_f32([1, 2, 3]), V.from_f32([4, 5, 6]), M.L1) - 9.0) < 0.01  # scalar.rs:133-143
    assert abs(vg.distance(V.from_f32([1, 0, 0]), V.from_f32([0, 1, 0]), M.Cosine) - 1.0) < 0.01  # scalar.rs:146-157
    assert abs(vg.distance(V.from_f32([1, 2, 3]), V.from_f32([2, 4, 6]), M.Cosine)) < 0.01  # scalar.rs:160-171
    assert abs(vg.distance(V.from_i8([1, 2, 3]), V.from_i8([4, 5, 6]), M.L2) - 5.196) < 0.01  # scalar.rs:174-184
    assert abs(vg.distance(V.from_i8([1, 2, 3]), V.from_i8([4, 5, 6]), M.L1) - 9.0) < 0.01  # scalar.rs:187-197
    a = V(vg.VectorType.Bit, 32, bytes([1, 0, 1, 0]))
    b = V(vg.VectorType.Bit, 32, bytes([0, 1, 1, 0]))
    assert vg.distance(a, b, M.Hamming) == 2.0  # scalar.rs:200-212
    assert abs(vg.distance(V.from_f32([1, 0, 0]), V.from_f32([0, 1, 0]), M.L2) - 1.414) < 0.01  # integration_test.rs:437-456


def test_ref_distance_errors(vg, gpu):
    V, M = vg.Vector, vg.DistanceMetric
    with pytest.raises(vg.DimensionMismatch):  # mod.rs:155-162
        vg.distance(V.from_f32([1, 2, 3]), V.from_f32([1, 2]), M.L2)
    with pytest.raises(vg.InvalidDistanceMetric):  # mod.rs:78-82
        vg.distance(V.from_f32([1, 2]), V.from_f32([1, 2]), M.Hamming)
    with pytest.raises(vg.InvalidDistanceMetric):
        vg.distance(V(vg.VectorType.Bit, 8, b"\x01"), V(vg.VectorType.Bit, 8, b"\x03"), M.L2)
    with vg.Slab(vg.VectorType.Bit, 64) as s:
        s.load(np.zeros((2, 8), dtype="u1"))
        with pytest.raises(vg.InvalidDistanceMetric):
            s.knn(np.zeros(8, dtype="u1"), 1, M.Cosine)


def test_ref_knn_simple(vg, gpu):
    # tests/test_knn_simple.rs:34-53 (default metric cosine): 2 rows, first rowid 1, tie 2 vs 3 -> 2
    with vg.Slab(vg.VectorType.Float32, 3) as s:
        s.load(np.eye(3, dtype="<f4"))
        res = vg.brute_force_search(s, np.array([1, 0, 0], dtype="<f4").tobytes(), 2, vg.DistanceMetric.Cosine)
    assert [r for r, _ in res] == [1, 2] and res[0][1] == 0.0 and res[1][1] == 1.0


def test_ref_knn_integration_rows(vg, gpu):
    # tests/integration_test.rs:635-678
    v = np.array([[i, i + 1, i + 2] for i in range(1, 6)], dtype="<f4")
    with vg.Slab(vg.VectorType.Float32, 3) as s:
        for i in range(5):  # row-at-a-time inserts like the SQL INSERTs of the test
            s.upsert(i + 1, v[i].tobytes())
        res = vg.brute_force_search(s, v[0].tobytes(), 3, vg.DistanceMetric.Cosine)
        assert len(res) == 3 and res[0][0] == 1 and res[0][1] < 0.01
        res = vg.brute_force_search(s, v[0].tobytes(), 3, vg.DistanceMetric.L2)
        assert [r for r, _ in res] == [1, 2, 3]


def test_brute_force_k_semantics(vg, gpu):
    # `k as usize` (src/vtab.rs:2292): 0 -> empty; k > N -> N rows; negative -> all rows
    v = random_rows(F32, 9, 4, seed=3)
    with vg.Slab(F32, 4) as s:
        s.load(v)
        q = v[2].tobytes()
        assert vg.brute_force_search(s, q, 0, L2) == []
        assert len(vg.brute_force_search(s, q, 100, L2)) == 9
        assert len(vg.brute_force_search(s, q, -1, L2)) == 9
        assert vg.brute_force_search(s, q[:8], 3, L2) == []  # wrong-length query: every row errors and is skipped


@pytest.mark.parametrize("case", json.load(open(os.path.join(GOLDEN, "cases.json"))), ids=lambda c: c["name"])
def test_golden_fixtures(vg, gpu, case):
    z = np.load(os.path.join(GOLDEN, case["file"]))
    with vg.Slab(case["elem"], case["dims"]) as s:
        s.load(z["vectors"], z["rowids"])
        if "skip" in z:
            for i in np.flatnonzero(z["skip"]):
                s.delete(int(z["rowids"][i]))
        r, d, c = s.knn(z["queries"], case["k"], case["metric"])
    assert np.array_equal(r, z["out_rowids"])
    assert np.array_equal(c, z["out_counts"])
    assert rel_close(d, z["out_dists"], REL_TOL)
    if case["elem"] != F32 and case["metric"] != COSINE:
        assert np.array_equal(bits(d), bits(z["out_dists"]))  # integer classes: bit-exact


# ------------------------------------------------------------------ seeded sweeps vs the oracle
@pytest.mark.parametrize("elem,metric", PAIRS, ids=PAIR_IDS)
@pytest.mark.parametrize("dims", [1, 3, 8, 16, 17, 64, 100, 384, 768, 1024, 2000])
def test_knn_dims_sweep(vg, orc, gpu, elem, metric, dims):
    n = 700
    v = random_rows(elem, n, dims, seed=dims + 10 * elem + metric)
    q = random_rows(elem, 3, dims, seed=999 + dims)
    check_knn(vg, orc, elem, dims, v, q, 10, metric)


@pytest.mark.parametrize("elem,metric", PAIRS, ids=PAIR_IDS)
@pytest.mark.parametrize("nq", [1, 2, 3, 5, 8, 13])
def test_knn_query_batches(vg, orc, gpu, elem, metric, nq):
    dims = 96
    v = random_rows(elem, 3000, dims, seed=42 + elem)
    q = random_rows(elem, nq, dims, seed=43 + nq)
    check_knn(vg, orc, elem, dims, v, q, 7, metric)


@pytest.mark.parametrize("k", [1, 2, 10, 32, 33, 100, 128, 1000, 1024, 1025, 5000])
def test_knn_k_sweep(vg, orc, gpu, k):
    dims = 32
    v = random_rows(F32, 6000, dims, seed=77)
    q = random_rows(F32, 2, dims, seed=78)
    check_knn(vg, orc, F32, dims, v, q, k, L2)
    vi = random_rows(I8, 6000, dims, seed=79)
    check_knn(vg, orc, I8, dims, vi, vi[:2], k, L2)


@pytest.mark.parametrize("elem,metric", PAIRS, ids=PAIR_IDS)
def test_knn_heavy_ties(vg, orc, gpu, elem, metric):
    # tiny alphabet -> most distances collide; the rowid tie-break decides
    dims = 24 if elem == BIT else 6
    v = random_rows(elem, 5000, dims, seed=5, ties=True)
    q = random_rows(elem, 4, dims, seed=6, ties=True)
    check_knn(vg, orc, elem, dims, v, q, 50, metric)


def test_knn_adversarial_descending(vg, orc, gpu):
    # distances strictly decreasing with rowid: every row beats the running threshold
    n, dims = 20000, 8
    v = np.zeros((n, dims), dtype="<f4")
    v[:, 0] = np.arange(n, 0, -1)
    check_knn(vg, orc, F32, dims, v, np.zeros((1, dims), dtype="<f4"), 10, L2)
    check_knn(vg, orc, F32, dims, v, np.zeros((1, dims), dtype="<f4"), 200, L1)


def test_knn_all_equal_rows(vg, orc, gpu):
    v = np.ones((4097, 16), dtype="<f4")
    r, d, c = check_knn(vg, orc, F32, 16, v, v[:1], 10, L2)
    assert list(r[0]) == list(range(1, 11)) and np.all(d == 0)


@pytest.mark.parametrize("n", [0, 1, 2, 31, 32, 33, 1000])
def test_knn_small_tables(vg, orc, gpu, n):
    dims = 40
    v = random_rows(F32, max(n, 1), dims, seed=n)[:n]
    q = random_rows(F32, 2, dims, seed=500)
    with vg.Slab(F32, dims) as s:
        if n:
            s.load(v)
        r, d, c = s.knn(q, 5, COSINE)
        er, ed, ec = orc.knn(F32, dims, v, q, 5, COSINE)
        assert np.array_equal(r, er) and np.array_equal(c, ec) and np.array_equal(bits(d), bits(ed))


def test_knn_sparse_rowids_and_negative(vg, orc, gpu):
    dims = 20
    v = random_rows(I8, 900, dims, seed=8)
    rowids = np.sort(np.random.default_rng(9).choice(np.arange(-5000, 10**12, 7919), size=900, replace=False)).astype("<i8")
    check_knn(vg, orc, I8, dims, v, v[:3], 15, L1, rowids=rowids)


def test_knn_special_values(vg, orc, gpu):
    # zero vectors (cosine zero rules), huge magnitudes (inf distance), NaN ranks last
    dims = 12
    v = random_rows(F32, 300, dims, seed=10)
    v[5] = 0
    v[6] = 3e38
    v[7, 3] = np.nan
    q = random_rows(F32, 2, dims, seed=11)
    q[1] = 0
    for metric in (L2, L1, COSINE):
        check_knn(vg, orc, F32, dims, v, q, 300, metric)


# ------------------------------------------------------------------ slab maintenance (vtab.rs:1409/1684/1326 hooks)
def test_slab_upsert_delete_skip(vg, orc, gpu):
    dims = 10
    rng = np.random.default_rng(12)
    v = random_rows(F32, 50, dims, seed=12)
    rowids = np.arange(1, 51, dtype="<i8")
    skip = np.zeros(50, dtype="u1")
    with vg.Slab(F32, dims) as s:
        s.load(v)
        # update in place
        v[10] = rng.standard_normal(dims).astype("<f4")
        s.upsert(11, v[10].tobytes())
        # delete, delete absent, re-insert
        s.delete(20)
        s.delete(999)
        skip[19] = 1
        s.delete(30)
        v[29] = rng.standard_normal(dims).astype("<f4")
        s.upsert(30, v[29].tobytes())
        # wrong-length / empty blob => row skipped by scans (src/vtab.rs:2596-2613)
        s.upsert(40, b"\x00" * 7)
        skip[39] = 1
        s.upsert(41, b"")
        skip[40] = 1
        assert s.count() == (50, 47)
        assert s.get(20) is None and s.get(11) == v[10].tobytes()
        check_knn(vg, orc, F32, dims, v, v[:4], 50, L2, rowids=rowids, skip=skip, slab=s)
        # append (rowid = MAX+1, src/shadow.rs:888-900) and an out-of-order explicit rowid
        nv = rng.standard_normal((3, dims)).astype("<f4")
        s.upsert(51, nv[0].tobytes())
        s.upsert(1000, nv[1].tobytes())
        s.upsert(500, nv[2].tobytes())
        v2 = np.concatenate([v, nv[[0, 2, 1]]])
        rowids2 = np.concatenate([rowids, [51, 500, 1000]]).astype("<i8")
        skip2 = np.concatenate([skip, [0, 0, 0]]).astype("u1")
        assert s.count() == (53, 50)
        check_knn(vg, orc, F32, dims, v2, v2[-3:], 53, COSINE, rowids=rowids2, skip=skip2, slab=s)
        # scoring sees the same table
        out = s.score(v2[:1], np.array([51, 500, 1000, 20, 777], dtype="<i8"), np.array([0, 5], dtype="<u4"), L2)
        want = orc.distances(F32, dims, v2[[50, 51, 52]], v2[0], L2)
        assert np.array_equal(bits(out[:3]), bits(want)) and np.isnan(out[3]) and np.isnan(out[4])


def test_slab_append_in_pieces(vg, orc, gpu):
    dims = 16
    v = random_rows(BIT, 1000, dims * 8, seed=13)
    with vg.Slab(BIT, dims * 8) as s:
        s.load(v[:400])
        s.append(v[400:700])
        s.append(v[700:], rowids=np.arange(701, 1001))
        check_knn(vg, orc, BIT, dims * 8, v, v[:2], 10, HAMMING, slab=s)
        with pytest.raises(vg.InvalidParameter):
            s.append(v[:1], rowids=[5])  # not greater than the last rowid


# ------------------------------------------------------------------ K5: candidate scoring (search.rs:501-513)
@pytest.mark.parametrize("elem,metric", PAIRS, ids=PAIR_IDS)
def test_score_matches_oracle(vg, orc, gpu, elem, metric):
    dims = 384 if elem != BIT else 1024
    n = 2000
    v = random_rows(elem, n, dims, seed=21 + elem)
    q = random_rows(elem, 5, dims, seed=22)
    rng = np.random.default_rng(23)
    sizes = [1, 4, 32, 0, 17]  # per-expansion batch sizes incl. the 1-4 common case and max_m0=32
    offsets = np.concatenate([[0], np.cumsum(sizes)]).astype("<u4")
    cands = rng.integers(1, n + 1, size=offsets[-1]).astype("<i8")
    with vg.Slab(elem, dims) as s:
        s.load(v)
        out = s.score(q, cands, offsets, metric)
    for qi in range(5):
        for j in range(offsets[qi], offsets[qi + 1]):
            want = orc.distance(elem, q[qi], v[cands[j] - 1], metric)
            assert np.float32(out[j]).view("<u4") == np.float32(want).view("<u4")


@pytest.mark.parametrize("elem,metric", PAIRS, ids=PAIR_IDS)
def test_score_small_call_equals_the_general_path(vg, orc, gpu, elem, metric, monkeypatch):
    """A call of <= 256 pairs whose input fits 16 KB (one expansion of search_layer) is ONE launch reading pinned host memory
    (score_small_kernel); VECGPU_SCORE_SMALL=0 forces the upload / resolve / score / download sequence.  Same bits, on a
    sparse slab with deleted, empty-blob and absent rowids; and against the oracle."""
    dims = 96 if elem != BIT else 200
    n = 3000
    v = random_rows(elem, n, dims, seed=41 + elem)
    q = random_rows(elem, 3, dims, seed=42)
    rng = np.random.default_rng(43)
    rowids = np.sort(rng.choice(np.arange(1, 5 * n), size=n, replace=False)).astype("<i8")
    for sparse in (False, True):
        with vg.Slab(elem, dims) as s:
            s.load(v, rowids if sparse else None)
            ids = rowids if sparse else np.arange(1, n + 1, dtype="<i8")
            s.delete(int(ids[5]))
            s.upsert(int(ids[9]), b"")
            sizes = [32, 0, 63] if sparse else [32, 0, 64]  # an odd number of rowids: the queries behind them stay 16-byte aligned
            offsets = np.concatenate([[0], np.cumsum(sizes)]).astype("<u4")
            cands = ids[rng.integers(0, n, size=offsets[-1])].copy()
            cands[0], cands[1], cands[2], cands[40] = ids[5], ids[9], -7, int(ids[-1]) + 11  # deleted, skipped, absent, absent
            small = s.score(q, cands, offsets, metric)
            monkeypatch.setenv("VECGPU_SCORE_SMALL", "0")
            general = s.score(q, cands, offsets, metric)
            monkeypatch.delenv("VECGPU_SCORE_SMALL")
            assert np.array_equal(small.view("<u4"), general.view("<u4"))
            assert np.isnan(small[[0, 1, 2, 40]]).all()
            pos = {int(r): i for i, r in enumerate(ids)}
            for qi in range(3):
                for j in range(offsets[qi], offsets[qi + 1]):
                    if j in (0, 1, 2, 40):
                        continue
                    want = orc.distance(elem, q[qi], v[pos[int(cands[j])]], metric)
                    assert np.float32(small[j]).view("<u4") == np.float32(want).view("<u4")


def test_score_hnsw_cosine_contract(vg, orc, gpu):
    # HNSW cosine = L2 on normalised vectors, output d^2/2 (src/hnsw/mod.rs:129-146, insert.rs:300-322)
    dims = 64
    v = random_rows(F32, 200, dims, seed=31)
    vn = vg.normalize(v)
    assert np.array_equal(bits(vn), bits(orc.normalize(v)))
    qn = vg.Vector.from_f32(v[3]).normalize().as_f32()
    with vg.Slab(F32, dims) as s:
        s.load(vn)
        metric = vg.internal_distance_metric(vg.DistanceMetric.Cosine, True)
        d = s.score(qn, np.arange(1, 201, dtype="<i8"), np.array([0, 200], dtype="<u4"), metric)
    out = np.array([vg.convert_distance_for_output(vg.DistanceMetric.Cosine, True, x) for x in d], dtype="<f4")
    exact = np.array([orc.distance(F32, v[3], v[i], COSINE) for i in range(200)])
    assert np.all(np.abs(out - exact) < 1e-5)
    assert out[3] < 1e-6


def test_distance_pairs_bulk(vg, orc, gpu):
    for elem, metric in PAIRS:
        dims = 130 if elem != BIT else 136
        a = random_rows(elem, 300, dims, seed=41)
        b = random_rows(elem, 300, dims, seed=42)
        out = vg.distance_pairs(elem, a, b, metric)
        want = np.array([orc.distance(elem, a[i], b[i], metric) for i in range(300)], dtype="<f4")
        assert np.array_equal(bits(out), bits(want))


def test_i8_l1_exhaustive_bytes(vg, orc, gpu):
    # every (a, b) byte pair once: the SIMD abs-diff must be exact over the whole int8 range
    a = np.repeat(np.arange(-128, 128, dtype="i1"), 256).reshape(256, 256)
    b = np.tile(np.arange(-128, 128, dtype="i1"), 256).reshape(256, 256)
    for metric in (L1, L2, COSINE):
        out = vg.distance_pairs(I8, a, b, metric)
        want = np.array([orc.distance(I8, a[i], b[i], metric) for i in range(256)], dtype="<f4")
        assert np.array_equal(bits(out), bits(want))


# ------------------------------------------------------------------ K7 producers (vector.rs:444-608)
def test_producers_bit_exact(vg, orc, gpu):
    rng = np.random.default_rng(51)
    for dims in (1, 2, 7, 8, 9, 128, 384, 1000):
        x = (rng.standard_normal((64, dims)) * rng.choice([0.01, 1.0, 50.0], size=(64, 1))).astype("<f4")
        x[3] = 0.25  # all-equal row -> quantize_int8 gives zeros
        assert np.array_equal(vg.quantize_int8(x), orc.quantize_int8(x))
        assert np.array_equal(vg.quantize_int8_for_index(x), orc.quantize_int8_for_index(x))
        assert np.array_equal(vg.quantize_binary(x), orc.quantize_binary(x))
        assert np.array_equal(bits(vg.normalize(x)), bits(orc.normalize(x)))
    # half-way cases of round(): *.5 must round away from zero
    h = np.array([[0.5 / 127, 1.5 / 127, -0.5 / 127, 2.5 / 127, 1.0, -1.0, 3.0, -3.0]], dtype="<f4")
    assert np.array_equal(vg.quantize_int8_for_index(h), orc.quantize_int8_for_index(h))
    q = vg.Vector.from_f32([0.0, 0.5, 1.0]).quantize_int8().as_i8()  # src/vector.rs:777-789
    assert q[0] == -128 and q[2] == 127 and q[0] < q[1] < q[2]
    n = vg.Vector.from_f32([3.0, 4.0]).normalize().as_f32()  # src/vector.rs:746-759
    assert abs(n[0] - 0.6) < 1e-4 and abs(n[1] - 0.8) < 1e-4
    with pytest.raises(vg.InvalidParameter):
        vg.Vector.from_f32([0.0, 0.0]).normalize()  # src/vector.rs:451-455
    with pytest.raises(vg.InvalidVectorType):
        vg.Vector.from_i8([1, 2]).normalize()


# ------------------------------------------------------------------ synthetic generator: device == CPU restatement
@pytest.mark.parametrize("elem,dims,kind", [(F32, 384, 0), (F32, 768, 1), (F32, 5, 1), (I8, 1024, 0), (I8, 33, 0), (BIT, 1024, 0), (BIT, 77, 0)])
def test_synthetic_generator_matches_cpu(vg, orc, gpu, elem, dims, kind):
    n, first = 3000, 17
    with vg.Slab(elem, dims) as s:
        s.fill_synthetic(seed=1234, n=n, first_rowid=first, kind=kind)
        cpu = orc.synth_rows(elem, 1234, first, n, dims, kind)
        for rid in (first, first + 1, first + 1234, first + n - 1):
            assert s.get(rid) == cpu[rid - first].tobytes()
        assert s.get(first + n) is None
        metric = {F32: COSINE, I8: L2, BIT: HAMMING}[elem]
        q = orc.synth_rows(elem, 99, 1, 2, dims, kind)
        r, d, c = s.knn(q, 10, metric)
        er, ed, ec = orc.knn(elem, dims, cpu, q, 10, metric, rowids=np.arange(first, first + n))
        assert np.array_equal(r, er) and np.array_equal(bits(d), bits(ed))


# ------------------------------------------------------------------ larger shapes: config 1 in full, configs 2-4 scaled down
def test_config1_full(vg, orc, gpu):
    # vec0 float[384], 10K random vectors, brute-force L2 k=10 (+ cosine, L1), 100 queries
    n, dims = 10000, 384
    v = orc.synth_rows(F32, 1, 1, n, dims, 0)
    q = orc.synth_rows(F32, 2, 1, 100, dims, 0)
    with vg.Slab(F32, dims) as s:
        s.load(v)
        for metric in (L2, COSINE, L1):
            check_knn(vg, orc, F32, dims, v, q, 10, metric, slab=s)


@pytest.mark.parametrize("elem,dims,metric,k,n", [(F32, 768, COSINE, 10, 200_000), (I8, 1024, L2, 100, 200_000), (BIT, 1024, HAMMING, 10, 1_000_000)])
def test_configs_2_3_4_prefix(vg, orc, gpu, elem, dims, metric, k, n):
    kind = 1 if elem == F32 else 0
    with vg.Slab(elem, dims) as s:
        s.fill_synthetic(seed=3 + elem, n=n, kind=kind)
        cpu = orc.synth_rows(elem, 3 + elem, 1, n, dims, kind)
        q = orc.synth_rows(elem, 77, 1, 9, dims, kind)
        r, d, c = s.knn(q, k, metric)
        er, ed, ec = orc.knn(elem, dims, cpu, q, k, metric)
        assert np.array_equal(r, er) and np.array_equal(bits(d), bits(ed)) and np.array_equal(c, ec)
        # single-query launches agree with the batched pass
        r1, d1, _ = s.knn(q[:1], k, metric)
        assert np.array_equal(r1[0], r[0]) and np.array_equal(bits(d1[0]), bits(d[0]))


def test_full_size_properties(vg, orc, gpu):
    """Size-independent properties at a size the CPU cannot check row by row:
    sortedness, self-match at distance 0, idempotence, and k-prefix consistency."""
    n, dims = 2_000_000, 768
    with vg.Slab(F32, dims) as s:
        s.fill_synthetic(seed=3, n=n, kind=1)
        probe = [1, 777_777, n]
        q = orc.synth_rows(F32, 3, 1, 1, dims, 1)
        q = np.concatenate([orc.synth_rows(F32, 3, rid, 1, dims, 1) for rid in probe])
        r, d, c = s.knn(q, 10, COSINE)
        assert list(r[:, 0]) == probe and np.all(d[:, 0] <= 1e-6)  # each probe row is its own nearest neighbour
        assert np.all(np.diff(d, axis=1) >= 0)  # sorted
        r2, d2, _ = s.knn(q, 10, COSINE)
        assert np.array_equal(r, r2) and np.array_equal(bits(d), bits(d2))  # deterministic
        r5, d5, _ = s.knn(q, 5, COSINE)
        assert np.array_equal(r5, r[:, :5])  # top-5 is a prefix of top-10
        # the returned distances are the oracle's distances for those rows
        for qi in range(3):
            rows = np.concatenate([orc.synth_rows(F32, 3, int(x), 1, dims, 1) for x in r[qi]])
            want = orc.distances(F32, dims, rows, q[qi], COSINE)
            assert np.array_equal(bits(d[qi]), bits(want))


# ------------------------------------------------------------------ multi-GPU pieces on one GPU
@pytest.mark.parametrize("world", [2, 3, 8])
def test_cross_shard_merge_kernel(vg, orc, gpu, world):
    """Shards emulated on one device: per-shard vecgpu_knn_device results -> vecgpu_merge_device == one big slab."""
    import torch

    from sqlite_vec_hnsw_b200.dist import pack_local, shard_range, unpack_gathered

    dims, n, k = 16, 3001, 20
    v = random_rows(I8, n, dims, seed=61, ties=True)
    q = random_rows(I8, 5, dims, seed=62, ties=True)
    dq = torch.from_numpy(q).cuda()
    parts = []
    for rank in range(world):
        lo, hi = shard_range(n, rank, world)
        s = vg.Slab(I8, dims)
        s.load(v[lo:hi], np.arange(1 + lo, 1 + hi, dtype="<i8"))
        r, d = s.knn_device(dq, k, L2)
        torch.cuda.synchronize()
        parts.append(pack_local(r, d))
        s.close()
    gr, gd = unpack_gathered(torch.stack(parts))
    mr, md = vg.merge_device(gr, gd)
    torch.cuda.synchronize()
    er, ed, _ = orc.knn(I8, dims, v, q, k, L2)
    assert np.array_equal(mr.cpu().numpy(), er) and np.array_equal(bits(md.cpu().numpy()), bits(ed))


def test_knn_device_matches_host_api(vg, orc, gpu):
    import torch

    for elem, dims, metric in ((F32, 100, COSINE), (BIT, 1000, HAMMING), (F32, 6, L1)):
        v = random_rows(elem, 5000, dims, seed=71)
        q = random_rows(elem, 9, dims, seed=72)
        with vg.Slab(elem, dims) as s:
            s.load(v)
            r, d, c = s.knn(q, 10, metric)
            dr, dd = s.knn_device(torch.from_numpy(q).cuda(), 10, metric)
            torch.cuda.synchronize()
        assert np.array_equal(dr.cpu().numpy(), r) and np.array_equal(bits(dd.cpu().numpy()), bits(d))


def test_scan_repeat_launch_stress(vg, gpu):
    """Many back-to-back launches of every pipeline mode must neither hang nor change their answer
    (regression test for the barrier-phase aliasing bug)."""
    import torch

    for elem, dims, metric, n in ((F32, 768, COSINE, 300_000), (F32, 768, L1, 100_000), (I8, 1024, L2, 400_000), (BIT, 1024, HAMMING, 2_000_000)):
        with vg.Slab(elem, dims) as s:
            s.fill_synthetic(seed=9, n=n, kind=1 if elem == F32 else 0)
            q = torch.randn(dims, device="cuda") if elem == F32 else torch.randint(0, 255, (s.row_bytes,), dtype=torch.uint8, device="cuda")
            r0, d0 = s.knn_device(q, 10, metric)
            for _ in range(150):
                r, d = s.knn_device(q, 10, metric)
            torch.cuda.synchronize()
            assert torch.equal(r, r0) and torch.equal(d, d0)


# ------------------------------------------------------------------ K2: tensor-core batched path (tcgen05 TF32 candidate pass + exact re-rank)
@pytest.fixture(params=["1", "3"], ids=["tf32x1", "tf32x3"])
def tc_terms(request):
    """Both candidate passes: one TF32 MMA per product (default, wider certified bound) and 3xTF32."""
    old = os.environ.get("VECGPU_TC_TERMS")
    os.environ["VECGPU_TC_TERMS"] = request.param
    yield request.param
    if old is None:
        os.environ.pop("VECGPU_TC_TERMS", None)
    else:
        os.environ["VECGPU_TC_TERMS"] = old


@pytest.mark.parametrize("metric", [L2, COSINE], ids=["l2", "cos"])
@pytest.mark.parametrize("dims,nq,k", [(96, 16, 10), (768, 130, 10), (100, 33, 5), (384, 257, 32), (64, 40, 96), (17, 20, 3)])
def test_tc_batched_matches_oracle(vg, orc, gpu, tc_terms, metric, dims, nq, k):
    n = 30000
    kind = 1 if metric == COSINE else 0
    with vg.Slab(F32, dims) as s:
        s.fill_synthetic(seed=21, n=n, kind=kind)
        cpu = orc.synth_rows(F32, 21, 1, n, dims, kind)
        q = orc.synth_rows(F32, 22, 1, nq, dims, kind)
        before = vg.tc_stats()
        r, d, c = s.knn(q, k, metric)
        after = vg.tc_stats()
    assert after[0] - before[0] == nq, "the batch must have gone through the tensor-core path"
    er, ed, ec = orc.knn(F32, dims, cpu, q, k, metric)
    assert np.array_equal(r, er) and same_bits(d, ed) and np.array_equal(c, ec)


@pytest.mark.parametrize("cluster,nq", [("8", 1024), ("4", 512), ("2", 260), ("8", 1500)])
def test_tc_batched_cluster_multicast_matches_oracle(vg, orc, gpu, tc_terms, cluster, nq):
    # thread-block clusters: the CTAs of consecutive query tiles receive every row tile through TMA multicast
    # (VECGPU_TC_CLUSTER; 1500 queries = 12 query tiles -> clusters of 4)
    n, dims, k = 30000, 72, 10
    os.environ["VECGPU_TC_CLUSTER"] = cluster
    try:
        with vg.Slab(F32, dims) as s:
            s.fill_synthetic(seed=23, n=n, kind=1)
            cpu = orc.synth_rows(F32, 23, 1, n, dims, 1)
            q = orc.synth_rows(F32, 24, 1, nq, dims, 1)
            before = vg.tc_stats()
            for metric in (L2, COSINE):
                r, d, c = s.knn(q, k, metric)
                er, ed, ec = orc.knn(F32, dims, cpu, q, k, metric)
                assert np.array_equal(r, er) and same_bits(d, ed) and np.array_equal(c, ec)
            assert vg.tc_stats()[0] - before[0] == 2 * nq
    finally:
        del os.environ["VECGPU_TC_CLUSTER"]


@pytest.mark.parametrize("metric", [L2, COSINE], ids=["l2", "cos"])
@pytest.mark.parametrize("dims,nq,k", [(768, 130, 10), (100, 256, 5), (384, 1000, 32), (64, 300, 96)])
def test_tc_batched_cta_pairs_match_oracle(vg, orc, gpu, metric, dims, nq, k, monkeypatch):
    """VECGPU_TC_PAIR=1: tcgen05 cta_group::2 — two CTAs share one M = 256 MMA, each staging its own 128 queries and half of
    the row tile (pair TMA loads completing on the leader's mbarrier, multicast commits, remote accumulator release).  Same
    exact results; an odd number of query tiles (nq = 300 -> 3 tiles) falls back to single CTAs."""
    monkeypatch.setenv("VECGPU_TC_PAIR", "1")
    n = 20_000
    v = random_rows(F32, n, dims, seed=101)
    q = random_rows(F32, nq, dims, seed=102)
    with vg.Slab(F32, dims) as s:
        s.load(v)
        tc0 = vg.tc_stats()
        r, d, c = s.knn(q, k, metric)
        tc1 = vg.tc_stats()
        assert tc1[0] - tc0[0] == nq and tc1[1] == tc0[1]
    er, ed, ec = orc.knn_select(F32, dims, v, q, k, metric)
    assert np.array_equal(r, er) and same_bits(d, ed) and np.array_equal(c, ec)


def test_small_batches_take_the_cuda_core_route(vg, orc, gpu):
    # default routing: 32 queries x 20 k rows is below the tensor-core work threshold -> exact multi-query scan, same results
    n, dims, nq, k = 20000, 64, 32, 10
    v = random_rows(F32, n, dims, seed=91)
    q = random_rows(F32, nq, dims, seed=92)
    old = os.environ.pop("VECGPU_TC_MIN_WORK", None)
    try:
        with vg.Slab(F32, dims) as s:
            s.load(v)
            before = vg.tc_stats()[0]
            r, d, c = s.knn(q, k, L2)
            assert vg.tc_stats()[0] == before, "a small batch must not pay the tensor-core launch"
            os.environ["VECGPU_TC_MIN_WORK"] = "0"
            r1, d1, c1 = s.knn(q, k, L2)
            assert vg.tc_stats()[0] - before == nq
    finally:
        if old is not None:
            os.environ["VECGPU_TC_MIN_WORK"] = old
        else:
            os.environ.pop("VECGPU_TC_MIN_WORK", None)
    er, ed, ec = orc.knn(F32, dims, v, q, k, L2)
    assert np.array_equal(r, er) and same_bits(d, ed) and np.array_equal(r1, er) and same_bits(d1, ed)


def test_tc_batched_ties_fall_back_to_exact(vg, orc, gpu, tc_terms):
    # a tiny alphabet makes thousands of rows tie at the k-th distance: the candidate bound cannot be
    # certified, the affected queries must be re-run by the exact scan and still match bit for bit
    n, dims, nq, k = 20000, 32, 24, 10
    v = random_rows(F32, n, dims, seed=31, ties=True)
    q = random_rows(F32, nq, dims, seed=32, ties=True)
    with vg.Slab(F32, dims) as s:
        s.load(v)
        before = vg.tc_stats()
        for metric in (L2, COSINE):
            r, d, c = s.knn(q, k, metric)
            er, ed, ec = orc.knn(F32, dims, v, q, k, metric)
            assert np.array_equal(r, er) and same_bits(d, ed)
        after = vg.tc_stats()
    assert after[0] - before[0] == 2 * nq


def test_tc_batched_skips_and_sparse_rowids(vg, orc, gpu, tc_terms):
    n, dims, nq, k = 25000, 48, 64, 12
    v = random_rows(F32, n, dims, seed=41)
    q = random_rows(F32, nq, dims, seed=42)
    q[:8] = v[100:108]  # exact self matches: distance 0 must survive the ||q||^2+||x||^2-2qx cancellation
    rowids = (np.arange(n, dtype="<i8") * 3 + 7)
    skip = np.zeros(n, dtype="u1")
    skip[[100, 5000, n - 1]] = 1
    with vg.Slab(F32, dims) as s:
        s.load(v, rowids)
        for i in np.flatnonzero(skip):
            s.delete(int(rowids[i]))
        for metric in (L2, COSINE):
            r, d, c = s.knn(q, k, metric)
            er, ed, ec = orc.knn(F32, dims, v, q, k, metric, rowids=rowids, skip=skip)
            assert np.array_equal(r, er) and same_bits(d, ed) and np.array_equal(c, ec)
        # a write invalidates the cached row norms
        v[200] = q[20]
        s.upsert(int(rowids[200]), v[200].tobytes())
        r, d, c = s.knn(q, k, L2)
        er, ed, ec = orc.knn(F32, dims, v, q, k, L2, rowids=rowids, skip=skip)
        assert np.array_equal(r, er) and same_bits(d, ed) and r[20, 0] == rowids[200] and d[20, 0] == 0.0


def test_tc_batched_special_values(vg, orc, gpu, tc_terms):
    n, dims, nq, k = 16384, 40, 20, 10
    v = random_rows(F32, n, dims, seed=51)
    v[7] = 0
    v[8] = 3e38
    v[9, 2] = np.nan
    v[10] = 1e-30
    q = random_rows(F32, nq, dims, seed=52)
    q[3] = 0
    with vg.Slab(F32, dims) as s:
        s.load(v)
        for metric in (L2, COSINE):
            r, d, c = s.knn(q, k, metric)
            er, ed, ec = orc.knn(F32, dims, v, q, k, metric)
            assert np.array_equal(r, er) and same_bits(d, ed)


# ------------------------------------------------------------------ HNSW with GPU-batched candidate scoring (config 5)
def _recall(found, truth):
    hit = sum(len(set(f.tolist()) & set(t.tolist())) for f, t in zip(found, truth))
    return hit / truth.size


def test_hnsw_ref_recall_l2_1000x128(vg, orc, gpu):
    # tests/test_recall_accuracy.rs:6-135: vectors (i*100+j)/1000, query all 0.5, L2, k=10, recall >= 95 %
    n, dims = 1000, 128
    i = np.arange(n, dtype=np.int64)[:, None]
    j = np.arange(dims, dtype=np.int64)[None, :]
    v = ((i * 100 + j).astype("<f4") / np.float32(1000.0)).astype("<f4")
    q = np.full((1, dims), 0.5, dtype="<f4")
    with vg.Slab(F32, dims) as s:
        s.load(v)
        idx = vg.HnswIndex(s, vg.DistanceMetric.L2)  # default params M=32, efc=400 (src/hnsw/mod.rs:35-47)
        assert idx.rebuild() == n
        r, d, c = idx.search(q, 10, ef_search=200)
        er, ed, _ = orc.knn(F32, dims, v, q, 10, L2)
        assert c[0] == 10 and _recall(r, er) >= 0.95
        # every returned distance is the exact distance of that row (per-call scoring parity)
        want = np.array([orc.distance(F32, q[0], v[rid - 1], L2) for rid in r[0]], dtype="<f4")
        assert np.array_equal(bits(d[0]), bits(want)) and np.all(np.diff(d[0]) >= 0)
        idx.close()


def test_ref_scale_10k_x_128_query_latency(vg, orc, gpu):
    """tests/test_scale.rs:3-114: 10 000 vectors (i*128+j)/10000, default metric cosine, HNSW enabled; query j/128, k = 10;
    10 results, ascending distances, query latency < 100 ms.  Here: the exact scan equals the oracle, the HNSW answer is drawn
    from the same rows with their exact distances, and both calls meet the reference's latency bar with room to spare."""
    import time
    n, dims = 10_000, 128
    i = np.arange(n, dtype=np.int64)[:, None]
    j = np.arange(dims, dtype=np.int64)[None, :]
    v = ((i * 128 + j).astype("<f4") / np.float32(10000.0)).astype("<f4")
    q = (np.arange(dims, dtype="<f4") / np.float32(128.0)).reshape(1, dims)
    with vg.Slab(F32, dims) as s:
        for lo in range(0, n, 1000):  # the reference inserts in transactions of 1000 rows
            s.append(v[lo:lo + 1000])
        assert s.count() == (n, n)
        s.knn(q, 10, COSINE)
        t0 = time.perf_counter()
        r, d, c = s.knn(q, 10, COSINE)
        dt = time.perf_counter() - t0
        er, ed, ec = orc.knn(F32, dims, v, q, 10, COSINE)
        assert c[0] == 10 and np.array_equal(r, er) and np.array_equal(bits(d), bits(ed)) and np.all(np.diff(d[0]) >= 0)
        assert dt < 0.1
        idx = vg.HnswIndex.for_column(s, vg.DistanceMetric.Cosine)
        assert idx.rebuild() == n  # node count == row count (test_scale.rs:60-68)
        idx.search(q, 10)
        t0 = time.perf_counter()
        hr, hd, hc = idx.search(q, 10)
        dth = time.perf_counter() - t0
        assert hc[0] == 10 and np.all(np.diff(hd[0]) >= 0) and dth < 0.1
        want = np.array([orc.distance(F32, q[0], v[rid - 1], COSINE) for rid in hr[0]], dtype="<f4")
        assert np.all(np.abs(hd[0] - want) <= 1e-5 * np.maximum(1.0, np.abs(want)))  # d^2/2 on normalised vectors vs 1 - cos
        idx.close()


def test_hnsw_ref_recall_cosine_100x128(vg, orc, gpu):
    # tests/test_recall_cosine.rs:15-125: ((7i+13j)%100)/100, cosine, recall >= 90 %; HNSW cosine = L2 on
    # normalised vectors, output d^2/2 (src/hnsw/mod.rs:129-146)
    n, dims = 100, 128
    i = np.arange(n, dtype=np.int64)[:, None]
    j = np.arange(dims, dtype=np.int64)[None, :]
    v = ((((7 * i + 13 * j) % 100).astype("<f4")) / np.float32(100.0)).astype("<f4")
    v[v.sum(axis=1) == 0] = 1.0
    with vg.Slab(F32, dims) as s:
        s.load(vg.normalize(v))
        idx = vg.HnswIndex(s, vg.DistanceMetric.Cosine)
        idx.rebuild(batch=16)
        r, d, c = idx.search(v[:10], 10)
        er, ed, _ = orc.knn(F32, dims, v, v[:10], 10, COSINE)
        assert _recall(r, er) >= 0.90
        assert np.all(r[:, 0] == np.arange(1, 11)) and np.all(d[:, 0] < 1e-5)  # each query finds itself at distance ~0
        assert np.all(np.abs(d - np.take_along_axis(np.array([[orc.distance(F32, v[qi], v[rid - 1], COSINE) for rid in r[qi]] for qi in range(10)]), np.arange(10)[None, :].repeat(10, 0), 1)) < 1e-5)
        idx.close()


@pytest.mark.parametrize("elem,metric,dims", [(F32, L2, 64), (I8, L2, 128), (F32, L1, 32)])
def test_hnsw_random_recall_and_graph_invariants(vg, orc, gpu, elem, metric, dims):
    n, nq, k, M, efc = 20000, 200, 10, 16, 200
    v = random_rows(elem, n, dims, seed=91)
    q = random_rows(elem, nq, dims, seed=92)
    rowids = np.arange(n, dtype="<i8") * 2 + 5
    with vg.Slab(elem, dims) as s:
        s.load(v, rowids)
        idx = vg.HnswIndex(s, metric, M=M, ef_construction=efc, seed=7)
        idx.rebuild(batch=1024)
        st = idx.stats()
        assert st["nodes"] == n and st["distances_scored"] > 0
        er, ed, _ = orc.knn(elem, dims, v, q, k, metric, rowids=rowids)
        r, d, c = idx.search(q, k, ef_search=800)
        # i.i.d. random data is the hardest case for M=16 + keep-closest pruning; the reference's 95 % bar
        # (tests/test_recall_accuracy.rs:128-132) is met with a wider beam
        assert _recall(r, er) >= 0.95
        r200, _, _ = idx.search(q, k, ef_search=200)
        rec200 = _recall(r200, er)
        assert rec200 >= 0.85
        # lockstep batching must not cost recall: a 16x smaller insert batch gives the same quality
        idx_small = vg.HnswIndex(s, metric, M=M, ef_construction=efc, seed=7)
        idx_small.rebuild(batch=64)
        rs, _, _ = idx_small.search(q, k, ef_search=200)
        assert abs(_recall(rs, er) - rec200) < 0.03
        idx_small.close()
        fr, to, lv, ds = idx.export_edges()
        assert fr.size == st["edges"]
        # degree caps: max_m0 = 2M at level 0, M above (src/hnsw/insert.rs:421-426)
        for level, cap in ((0, 2 * M), (1, M)):
            m = lv == level
            if m.any():
                assert np.bincount(((fr[m] - 5) // 2).astype(np.int64)).max() <= cap
        assert np.all(fr != to) and set(np.unique(lv)) <= set(range(16))
        # stored edge distances are the exact distances of the pair
        for e in range(0, fr.size, max(1, fr.size // 50)):
            want = orc.distance(elem, v[(fr[e] - 5) // 2], v[(to[e] - 5) // 2], metric)
            assert np.float32(ds[e]).view("<u4") == np.float32(want).view("<u4")
        # same seed -> same graph (the reference's builds are not reproducible; ours are)
        idx2 = vg.HnswIndex(s, metric, M=M, ef_construction=efc, seed=7)
        idx2.rebuild(batch=1024)
        assert idx2.stats()["edges"] == st["edges"]
        idx.close()
        idx2.close()


def test_hnsw_empty_and_tiny(vg, gpu):
    with vg.Slab(F32, 8) as s:
        idx = vg.HnswIndex(s, L2, M=4, ef_construction=10)
        assert idx.rebuild() == 0
        r, d, c = idx.search(np.zeros((1, 8), dtype="<f4"), 3)
        assert c[0] == 0 and list(r[0]) == [-1, -1, -1]  # empty index -> no rows (src/hnsw/search.rs:279-281)
        s.load(np.eye(8, dtype="<f4")[:3])
        assert idx.rebuild() == 3
        r, d, c = idx.search(np.eye(8, dtype="<f4")[:1], 5)
        assert c[0] == 3 and r[0, 0] == 1 and d[0, 0] == 0.0
        with pytest.raises(vg.InvalidParameter):
            vg.HnswIndex(s, L2, M=1)  # M must be in [2,100] (src/sql_functions.rs:442-469)
        idx.close()


# ------------------------------------------------------------------ K4 batched: Hamming, lane = query (ham_batch_kernel)
@pytest.mark.parametrize("dims,nq,k,ties", [(1024, 70, 10, False), (1024, 33, 32, True), (96, 64, 17, False), (1000, 16, 1, True),
                                            (520, 100, 12, False)])
def test_hamming_batched_lane_per_query_matches_oracle(vg, orc, gpu, dims, nq, k, ties):
    n = 60000
    v = random_rows(BIT, n, dims, seed=601, ties=ties)
    q = random_rows(BIT, nq, dims, seed=602, ties=ties)
    rowids = np.arange(n, dtype="<i8") * 4 + 9
    skip = np.zeros(n, dtype="u1")
    skip[[0, 1, 31337, n - 1]] = 1
    with vg.Slab(BIT, dims) as s:
        s.load(v, rowids)
        for i in np.flatnonzero(skip):
            s.delete(int(rowids[i]))
        r, d, c = s.knn(q, k, HAMMING)
        os.environ["VECGPU_HAM_BATCH"] = "0"          # the multi-query scan it replaces gives the same answer
        try:
            r0, d0, c0 = s.knn(q, k, HAMMING)
        finally:
            del os.environ["VECGPU_HAM_BATCH"]
    er, ed, ec = orc.knn(BIT, dims, v, q, k, HAMMING, rowids=rowids, skip=skip)
    assert np.array_equal(r, er) and same_bits(d, ed) and np.array_equal(c, ec)
    assert np.array_equal(r0, er) and same_bits(d0, ed)


# ------------------------------------------------------------------ slab maintenance: tombstones + compaction
@pytest.mark.parametrize("elem,metric,dims", [(F32, COSINE, 48), (I8, L2, 64), (BIT, HAMMING, 96)])
def test_slab_compact_drops_tombstones_and_keeps_results(vg, orc, gpu, elem, metric, dims):
    n, k = 20000, 12
    v = random_rows(elem, n, dims, seed=501)
    q = random_rows(elem, 20, dims, seed=502)
    rng = np.random.default_rng(503)
    dead = np.sort(rng.choice(n, size=n // 3, replace=False))
    skip = np.zeros(n, dtype="u1")
    skip[dead] = 1
    with vg.Slab(elem, dims) as s:
        s.load(v)                                   # dense rowids 1..n
        for p in dead:
            s.delete(int(p) + 1)
        er, ed, ec = orc.knn(elem, dims, v, q, k, metric, skip=skip)
        r0, d0, c0 = s.knn(q, k, metric)            # tombstones are skipped by the scan ...
        assert np.array_equal(r0, er) and same_bits(d0, ed)
        assert s.count() == (n, n - len(dead))
        assert s.compact() == len(dead)             # ... and physically gone after the compaction
        assert s.count() == (n - len(dead), n - len(dead)) and s.compact() == 0
        r1, d1, c1 = s.knn(q, k, metric)
        assert np.array_equal(r1, er) and same_bits(d1, ed) and np.array_equal(c1, ec)
        r2, d2, c2 = s.knn(q[:1], k, metric)        # single-query path too
        assert np.array_equal(r2, er[:1]) and same_bits(d2, ed[:1])
        # a dropped rowid can come back (out-of-order insert), and candidate scoring resolves the sparse rowids
        back = int(dead[5])
        s.upsert(back + 1, v[back].tobytes())
        skip[back] = 0
        er3, ed3, _ = orc.knn(elem, dims, v, q, k, metric, skip=skip)
        r3, d3, _ = s.knn(q, k, metric)
        assert np.array_equal(r3, er3) and same_bits(d3, ed3)
        sc = s.score(q[:1], er3[0], np.array([0, k], dtype="<u4"), metric)
        assert same_bits(sc, ed3[0])


def test_hnsw_index_goes_stale_when_rows_move(vg, gpu):
    v = random_rows(F32, 4000, 16, seed=511)
    with vg.Slab(F32, 16) as s:
        s.load(v)
        idx = vg.HnswIndex(s, L2, M=8, ef_construction=40)
        idx.rebuild()
        idx.search(v[:2], 3)
        s.delete(7)
        idx.search(v[:2], 3)                      # a tombstone does not move rows: still fine
        assert s.compact() == 1
        with pytest.raises(vg.VecError):          # positions changed: the index must be rebuilt
            idx.search(v[:2], 3)
        idx.rebuild()
        r, d, c = idx.search(v[:2], 3)
        assert r[0, 0] == 1 and r[1, 0] == 2
        idx.close()


# ------------------------------------------------------------------ K6: whole search_layer on the device (one warp per query)
class _hnsw_mode:
    """VECGPU_HNSW_DEVICE=0 selects the lockstep driver (one scoring launch per expansion round)."""

    def __init__(self, device):
        self.v = "1" if device else "0"

    def __enter__(self):
        self.old = os.environ.get("VECGPU_HNSW_DEVICE")
        os.environ["VECGPU_HNSW_DEVICE"] = self.v

    def __exit__(self, *a):
        if self.old is None:
            os.environ.pop("VECGPU_HNSW_DEVICE", None)
        else:
            os.environ["VECGPU_HNSW_DEVICE"] = self.old


@pytest.mark.parametrize(
    "elem,metric,dims,n,M,efc",
    [
        (F32, L2, 48, 6000, 16, 200),
        (F32, COSINE, 40, 3000, 8, 64),
        (F32, L1, 24, 2500, 6, 40),
        (I8, L2, 8, 4000, 16, 100),      # tiny int8 rows: distances tie all the time -> the tie rules are exercised
        (I8, COSINE, 64, 2000, 12, 80),
        (I8, L1, 16, 2000, 5, 30),
        (BIT, HAMMING, 64, 5000, 16, 120),  # integer distances 0..64: massive ties
        (F32, L2, 16, 3000, 50, 30),     # max_m0 = 100 > ef: every list longer than a warp, take > ef
    ],
)
def test_hnsw_device_search_equals_lockstep_and_restatement(vg, orc, gpu, elem, metric, dims, n, M, efc):
    from oracle import hnsw_ref

    v = random_rows(elem, n, dims, seed=401)
    q = random_rows(elem, 96, dims, seed=402)
    rowids = np.arange(n, dtype="<i8") * 3 + 11
    with vg.Slab(elem, dims) as s:
        s.load(v, rowids)
        with _hnsw_mode(True):
            idx = vg.HnswIndex(s, metric, M=M, ef_construction=efc, seed=5, normalize_vectors=False)
            idx.rebuild(batch=256)
            ds = idx.device_stats()
            assert ds["queries"] == n - 1 and ds["launches"] > 0  # every insert but the first searched on the device
        with _hnsw_mode(False):
            idx_h = vg.HnswIndex(s, metric, M=M, ef_construction=efc, seed=5, normalize_vectors=False)
            idx_h.rebuild(batch=256)
            assert idx_h.device_stats()["queries"] == 0
        # identical graphs: same lists in the same order with the same stored distances
        e_d, e_h = idx.export_edges(), idx_h.export_edges()
        for a, b in zip(e_d, e_h):
            assert np.array_equal(a.view("u1"), b.view("u1"))
        # device search with the ordered edge replay on the host threads instead of the device link kernels: same graph
        os.environ["VECGPU_HNSW_LINK"] = "host"
        try:
            with _hnsw_mode(True):
                idx_hl = vg.HnswIndex(s, metric, M=M, ef_construction=efc, seed=5, normalize_vectors=False)
                idx_hl.rebuild(batch=256)
        finally:
            del os.environ["VECGPU_HNSW_LINK"]
        for a, b in zip(e_d, idx_hl.export_edges()):
            assert np.array_equal(a.view("u1"), b.view("u1"))
        idx_hl.close()
        assert idx.entry_point() == idx_h.entry_point()
        for k, ef in ((10, 64), (1, 1), (7, 300), (40, 10)):
            with _hnsw_mode(True):
                r1, d1, c1 = idx.search(q, k, ef_search=ef)
            with _hnsw_mode(False):
                r2, d2, c2 = idx.search(q, k, ef_search=ef)       # device-built graph, lockstep walk
                r3, d3, c3 = idx_h.search(q, k, ef_search=ef)
            assert np.array_equal(r1, r2) and same_bits(d1, d2) and np.array_equal(c1, c2)
            assert np.array_equal(r1, r3) and same_bits(d1, d3)
        # literal restatement of search_hnsw (oracle/hnsw_ref.py) over the exported graph
        fr, to, lv, _ = e_d
        nbrs = hnsw_ref.adjacency_from_edges(fr, to, lv)
        entry, entry_level = idx.entry_point()
        k, ef = 10, 50
        with _hnsw_mode(True):
            r1, d1, c1 = idx.search(q[:12], k, ef_search=ef)
        for qi in range(12):
            dist_of = lambda rid: float(orc.distance(elem, q[qi], v[(rid - 11) // 3], metric))
            want = hnsw_ref.search_hnsw(dist_of, nbrs, entry, entry_level, k, ef)
            assert c1[qi] == len(want)
            assert [int(x) for x in r1[qi, : len(want)]] == [w[0] for w in want]
            assert same_bits(d1[qi, : len(want)], np.array([w[1] for w in want], dtype="<f4"))
        assert idx.device_stats()["fallbacks"] <= ds["fallbacks"] + 96 * 8  # counters move, nothing else to assert
        idx.close()
        idx_h.close()


def test_hnsw_device_capacity_fallback_is_exact(vg, orc, gpu):
    # a visited table that is too small for the beam: the kernel flags the query and the lockstep driver answers it
    n, dims = 8000, 16
    v = random_rows(F32, n, dims, seed=421)
    q = random_rows(F32, 40, dims, seed=422)
    with vg.Slab(F32, dims) as s:
        s.load(v)
        idx = vg.HnswIndex(s, L2, M=16, ef_construction=100, seed=1)
        idx.rebuild()
        before = idx.device_stats()
        os.environ["VECGPU_HNSW_VIS_LOG2"] = "12"  # 4096 slots, 3072 usable: ef=1500 visits more than that
        try:
            r1, d1, c1 = idx.search(q, 10, ef_search=1500)
        finally:
            del os.environ["VECGPU_HNSW_VIS_LOG2"]
        after = idx.device_stats()
        assert after["fallbacks"] - before["fallbacks"] > 0
        with _hnsw_mode(False):
            r2, d2, c2 = idx.search(q, 10, ef_search=1500)
        assert np.array_equal(r1, r2) and same_bits(d1, d2) and np.array_equal(c1, c2)
        r3, d3, c3 = idx.search(q, 10, ef_search=1500)  # full-size table: no fallback, same answer
        assert idx.device_stats()["fallbacks"] == after["fallbacks"]
        assert np.array_equal(r1, r3) and same_bits(d1, d3)
        # massive ties (two distinct points, thousands of copies): the sorted array never overflows (<= 2 ef - 1 entries)
        vt = np.zeros((3000, 8), dtype="i1")
        vt[:, 0] = (np.arange(3000) % 2).astype("i1")
        with vg.Slab(I8, 8) as st:
            st.load(vt)
            with _hnsw_mode(True):
                it = vg.HnswIndex(st, L2, M=8, ef_construction=40, seed=1)
                it.rebuild(batch=128)
                ra, da, ca = it.search(np.zeros((4, 8), dtype="i1"), 10, ef_search=32)
                assert it.device_stats()["fallbacks"] == 0
            with _hnsw_mode(False):
                ih = vg.HnswIndex(st, L2, M=8, ef_construction=40, seed=1)
                ih.rebuild(batch=128)
                rb, db, cb = ih.search(np.zeros((4, 8), dtype="i1"), 10, ef_search=32)
            for a_, b_ in zip(it.export_edges(), ih.export_edges()):
                assert np.array_equal(a_.view("u1"), b_.view("u1"))
            assert np.array_equal(ra, rb) and same_bits(da, db) and np.array_equal(ca, cb)
            it.close()
            ih.close()
        idx.close()


def test_hnsw_device_build_overflow_batches_are_exact(vg, orc, gpu):
    # a rebuild whose inserts overflow the (shrunk) visited table: those batches are linked by the host loop after the
    # host lists have been refreshed from the device, the device copy is rebuilt, and the result is still the same graph
    n, dims = 6000, 12
    v = random_rows(F32, n, dims, seed=431)
    with vg.Slab(F32, dims) as s:
        s.load(v)
        os.environ["VECGPU_HNSW_VIS_LOG2"] = "12"
        try:
            with _hnsw_mode(True):
                idx = vg.HnswIndex(s, L2, M=16, ef_construction=1200, seed=2)
                idx.rebuild(batch=512)
                st = idx.device_stats()
        finally:
            del os.environ["VECGPU_HNSW_VIS_LOG2"]
        assert st["fallbacks"] > 0 and st["queries"] == n - 1
        with _hnsw_mode(False):
            idx_h = vg.HnswIndex(s, L2, M=16, ef_construction=1200, seed=2)
            idx_h.rebuild(batch=512)
        for a, b in zip(idx.export_edges(), idx_h.export_edges()):
            assert np.array_equal(a.view("u1"), b.view("u1"))
        q = random_rows(F32, 20, dims, seed=432)
        r1, d1, _ = idx.search(q, 10, ef_search=100)
        r2, d2, _ = idx_h.search(q, 10, ef_search=100)
        assert np.array_equal(r1, r2) and same_bits(d1, d2)
        idx.close()
        idx_h.close()


def test_hnsw_huge_ef_uses_the_lockstep_driver(vg, orc, gpu):
    # ef_search beyond what a warp's shared-memory array can hold: answered by the lockstep driver, same semantics
    n, dims = 30000, 8
    v = random_rows(F32, n, dims, seed=441)
    q = random_rows(F32, 3, dims, seed=442)
    with vg.Slab(F32, dims) as s:
        s.load(v)
        idx = vg.HnswIndex(s, L2, M=8, ef_construction=40, seed=3)
        idx.rebuild()
        before = idx.device_stats()["queries"]
        r, d, c = idx.search(q, 10, ef_search=20000)
        assert idx.device_stats()["queries"] == before      # not the device walk
        er, ed, _ = orc.knn(F32, dims, v, q, 10, L2)
        assert _recall(r, er) >= 0.9 and np.all(np.diff(d, axis=1) >= 0)
        idx.close()


def test_hnsw_device_large_batch_throughput_path(vg, orc, gpu):
    # many queries in one launch (more than one wave of warps), results independent of the launch shape
    n, dims, nq = 20000, 32, 20000
    v = random_rows(F32, n, dims, seed=411)
    q = random_rows(F32, nq, dims, seed=412)
    with vg.Slab(F32, dims) as s:
        s.load(v)
        idx = vg.HnswIndex(s, L2, M=16, ef_construction=100, seed=3)
        idx.rebuild()
        r_all, d_all, c_all = idx.search(q, 10, ef_search=64)
        for lo in (0, 7777, nq - 5):
            r1, d1, c1 = idx.search(q[lo : lo + 5], 10, ef_search=64)
            assert np.array_equal(r1, r_all[lo : lo + 5]) and same_bits(d1, d_all[lo : lo + 5])
        er, _, _ = orc.knn(F32, dims, v, q[:300], 10, L2)
        assert _recall(r_all[:300], er) >= 0.80
        assert idx.device_stats()["fallbacks"] == 0
        idx.close()


# ------------------------------------------------------------------ K3 batched: int8 L2 on the tensor cores (tcgen05 kind::i8), exact
@pytest.mark.parametrize("dims,nq,k", [(1024, 64, 100), (128, 16, 10), (100, 130, 7), (33, 40, 1), (2000, 20, 33)])
def test_tc_int8_batched_matches_oracle(vg, orc, gpu, dims, nq, k):
    n = 30000
    v = random_rows(I8, n, dims, seed=61)
    q = random_rows(I8, nq, dims, seed=62)
    rowids = np.arange(n, dtype="<i8") * 5 - 1000
    skip = np.zeros(n, dtype="u1")
    skip[[0, 777, n - 1]] = 1
    with vg.Slab(I8, dims) as s:
        s.load(v, rowids)
        for i in np.flatnonzero(skip):
            s.delete(int(rowids[i]))
        before = vg.tc_stats()
        r, d, c = s.knn(q, k, L2)
        assert vg.tc_stats()[0] - before[0] == nq, "the batch must have gone through the int8 tensor-core path"
    er, ed, ec = orc.knn(I8, dims, v, q, k, L2, rowids=rowids, skip=skip)
    assert np.array_equal(r, er) and np.array_equal(bits(d), bits(ed)) and np.array_equal(c, ec)


@pytest.mark.parametrize("sample_tiles", ["0", "40"])
@pytest.mark.parametrize("ties", [False, True])
def test_tc_int8_sampled_thresholds_are_exact(vg, orc, gpu, sample_tiles, ties):
    # nq = 1024 -> 8 query tiles, 18 row groups: the launch that ranks a prefix of the slab first and hands its k-th
    # key to the main launch as the admission bound (VECGPU_TCI_SAMPLE forces / disables it) must not change any result;
    # with ties the bound is hit by equal distances all the time
    n, dims, nq, k = 30000, 64, 1024, 100
    v = random_rows(I8, n, dims, seed=81, ties=ties)
    q = random_rows(I8, nq, dims, seed=82, ties=ties)
    rowids = np.arange(n, dtype="<i8") * 2 + 7
    skip = np.zeros(n, dtype="u1")
    skip[[3, 5000, 9999, n - 2]] = 1   # skipped rows inside the sampled prefix and outside it
    os.environ["VECGPU_TCI_SAMPLE"] = sample_tiles
    try:
        with vg.Slab(I8, dims) as s:
            s.load(v, rowids)
            for i in np.flatnonzero(skip):
                s.delete(int(rowids[i]))
            before = vg.tc_stats()
            r, d, c = s.knn(q, k, L2)
            assert vg.tc_stats()[0] - before[0] == nq
    finally:
        del os.environ["VECGPU_TCI_SAMPLE"]
    er, ed, ec = orc.knn(I8, dims, v, q, k, L2, rowids=rowids, skip=skip)
    assert np.array_equal(r, er) and np.array_equal(bits(d), bits(ed)) and np.array_equal(c, ec)


def test_tc_int8_batched_heavy_ties_and_extremes(vg, orc, gpu):
    # tiny alphabet: thousands of equal distances, the rowid tie-break decides every rank; plus +-128/127 rows
    n, dims, nq, k = 20000, 64, 32, 50
    v = random_rows(I8, n, dims, seed=71, ties=True)
    v[5] = -128
    v[6] = 127
    q = random_rows(I8, nq, dims, seed=72, ties=True)
    q[1] = 127
    q[2] = -128
    with vg.Slab(I8, dims) as s:
        s.load(v)
        r, d, c = s.knn(q, k, L2)
    er, ed, ec = orc.knn(I8, dims, v, q, k, L2)
    assert np.array_equal(r, er) and np.array_equal(bits(d), bits(ed))


def test_sharded_batched_queries_through_tensor_cores(vg, orc, gpu):
    """Rowid-range shards (emulated on one device) + tensor-core batches + cross-shard merge == one big slab."""
    import torch

    from sqlite_vec_hnsw_b200.dist import pack_local, shard_range, unpack_gathered

    dims, n, k, nq, world = 64, 40000, 10, 48, 3
    v = orc.synth_rows(F32, 31, 1, n, dims, 1)
    q = orc.synth_rows(F32, 32, 1, nq, dims, 1)
    dq = torch.from_numpy(q.copy()).cuda()
    parts = []
    before = vg.tc_stats()[0]
    for rank in range(world):
        lo, hi = shard_range(n, rank, world)
        s = vg.Slab(F32, dims)
        s.load(v[lo:hi], np.arange(1 + lo, 1 + hi, dtype="<i8"))
        r, d = s.knn_device(dq, k, COSINE)
        torch.cuda.synchronize()
        parts.append(pack_local(r, d))
        s.close()
    assert vg.tc_stats()[0] - before == world * nq
    mr, md = vg.merge_device(*unpack_gathered(torch.stack(parts)))
    torch.cuda.synchronize()
    er, ed, _ = orc.knn(F32, dims, v, q, k, COSINE)
    assert np.array_equal(mr.cpu().numpy(), er) and np.array_equal(bits(md.cpu().numpy()), bits(ed))


# ------------------------------------------------------------------ randomized sweep: every dispatch path vs the oracle
@pytest.mark.parametrize("seed", range(40))
def test_knn_fuzz(vg, orc, gpu, seed):
    rng = np.random.default_rng(1000 + seed)
    elem, metric = PAIRS[rng.integers(len(PAIRS))]
    dims = int(rng.choice([1, 2, 5, 8, 15, 16, 31, 32, 33, 63, 64, 96, 127, 128, 200, 256, 384, 513, 768, 1000, 1536]))
    if elem == BIT:
        dims = int(rng.choice([1, 7, 8, 9, 64, 100, 128, 1000, 1024, 2048, 4100]))
    n = int(rng.choice([1, 2, 17, 100, 1000, 5000, 9000, 20000]))
    nq = int(rng.choice([1, 1, 2, 3, 7, 8, 9, 16, 17, 40]))
    k = int(rng.choice([1, 2, 5, 10, 31, 32, 33, 100, 129]))
    ties = bool(rng.integers(4) == 0)
    v = random_rows(elem, n, dims, seed=seed * 3 + 1, ties=ties)
    q = random_rows(elem, nq, dims, seed=seed * 3 + 2, ties=ties)
    rowids = None
    if rng.integers(2):
        rowids = np.cumsum(rng.integers(1, 1000, size=n)).astype("<i8") - int(rng.integers(0, 10**6))
    skip = None
    if rng.integers(2) and n > 3:
        skip = np.zeros(n, dtype="u1")
        skip[rng.choice(n, size=max(1, n // 10), replace=False)] = 1
    check_knn(vg, orc, elem, dims, v, q, k, metric, rowids=rowids, skip=skip)


# ------------------------------------------------------------------ size-independent properties at BASELINE's full single-GPU sizes
@pytest.mark.parametrize("elem,dims,metric,k,n", [(I8, 1024, L2, 100, 50_000_000), (BIT, 1024, HAMMING, 10, 62_500_000)])
def test_full_size_properties_int8_and_bit(vg, orc, gpu, elem, dims, metric, k, n):
    with vg.Slab(elem, dims) as s:
        s.fill_synthetic(seed=4 + elem, n=n, kind=0)
        probe = [1, 31_415_926, n]
        q = np.concatenate([orc.synth_rows(elem, 4 + elem, rid, 1, dims, 0) for rid in probe])
        r, d, c = s.knn(q, k, metric)
        assert list(r[:, 0]) == probe and np.all(d[:, 0] == 0)          # every probe row is its own nearest neighbour
        assert np.all(np.diff(d, axis=1) >= 0) and np.all(c == k)          # sorted, full
        for qi in range(3):                                               # ties broken by ascending rowid
            same = d[qi, 1:] == d[qi, :-1]
            assert np.all(r[qi, 1:][same] > r[qi, :-1][same])
        r2, d2, _ = s.knn(q, k, metric)
        assert np.array_equal(r, r2) and np.array_equal(bits(d), bits(d2))  # deterministic
        rk, _, _ = s.knn(q, max(1, k // 2), metric)
        assert np.array_equal(rk, r[:, : max(1, k // 2)])                  # top-k/2 is a prefix of top-k
        for qi in range(3):                                               # returned distances are the oracle's for those rows
            rows = np.concatenate([orc.synth_rows(elem, 4 + elem, int(x), 1, dims, 0) for x in r[qi]])
            want = orc.distances(elem, dims, rows, q[qi], metric)
            assert np.array_equal(bits(d[qi]), bits(want))


# ------------------------------------------------------------------ round-2 regressions: slab maintenance
def test_reload_does_not_resurrect_tombstones(vg, orc, gpu):
    """Delete row A, load() new data, delete row B: the device copy of the skip flags must not keep A's old tombstone
    (it used to: the reload cleared the host mirror only and the next delete uploaded a single byte)."""
    dims, n = 16, 500
    v1 = random_rows(F32, n, dims, seed=1)
    v2 = random_rows(F32, n, dims, seed=2)
    with vg.Slab(F32, dims) as s:
        s.load(v1)
        s.delete(7)
        s.delete(300)
        s.load(v2)                      # every position now holds a live row again
        s.delete(450)                   # rowid B
        skip = np.zeros(n, dtype="u1")
        skip[449] = 1
        for q in (v2[6], v2[299], v2[449]):
            r, d, c = s.knn(q, 3, L2)
            er, ed, ec = orc.knn(F32, dims, v2, q, 3, L2, skip=skip)
            assert np.array_equal(r, er) and same_bits(d, ed)
        assert s.knn(v2[6], 1, L2)[0][0, 0] == 7 and s.knn(v2[299], 1, L2)[0][0, 0] == 300
        # same through fill_synthetic and through a compaction down to zero rows
        s.fill_synthetic(seed=5, n=n, kind=1)
        s.delete(2)
        cpu = orc.synth_rows(F32, 5, 1, n, dims, 1)
        assert s.knn(cpu[6], 1, L2)[0][0, 0] == 7 and s.knn(cpu[299], 1, L2)[0][0, 0] == 300 and s.knn(cpu[449], 1, L2)[0][0, 0] == 450


@pytest.mark.parametrize("elem,metric,dims", [(F32, COSINE, 96), (F32, L2, 64), (I8, L2, 128)])
def test_norm_cache_is_updated_by_upserts_not_discarded(vg, orc, gpu, elem, metric, dims):
    """Batched (tensor-core) queries cache |row|^2; upserts, appends and out-of-order inserts must leave it correct."""
    n, nq, k = 20_000, 32, 10
    v = random_rows(elem, n, dims, seed=41)
    q = random_rows(elem, nq, dims, seed=42)
    rowids = np.arange(10, 10 + 2 * n, 2, dtype="<i8")  # gaps, so out-of-order inserts are possible
    with vg.Slab(elem, dims) as s:
        s.load(v, rowids)
        tc0 = vg.tc_stats()[0]
        s.knn(q, k, metric)
        assert vg.tc_stats()[0] - tc0 == nq  # the cache exists now
        vv, rr = v.copy(), rowids.copy()
        # in-place update: make row 1234 the exact match of query 0
        vv[1234] = q[0]
        s.upsert(int(rr[1234]), vv[1234].tobytes())
        # append: exact match of query 1 at the end
        s.upsert(int(rr[-1] + 7), q[1].tobytes())
        vv = np.concatenate([vv, q[1:2]])
        rr = np.concatenate([rr, [rr[-1] + 7]])
        # out-of-order insert: exact match of query 2 in the middle (positions shift)
        new_id = int(rr[5000] + 1)
        s.upsert(new_id, q[2].tobytes())
        vv = np.concatenate([vv[:5001], q[2:3], vv[5001:]])
        rr = np.concatenate([rr[:5001], [new_id], rr[5001:]])
        l0 = vg.launch_count()
        r, d, c = s.knn(q, k, metric)
        l1 = vg.launch_count()
        er, ed, ec = orc.knn(elem, dims, vv, q, k, metric, rowids=rr)
        assert np.array_equal(r, er) and same_bits(d, ed)
        assert r[0, 0] == rr[1234] and r[1, 0] == rr[-1] and r[2, 0] == new_id
        # none of the three writes discarded the cache: the same call again launches exactly as many kernels (no norm pass)
        r2, d2, _ = s.knn(q, k, metric)
        assert vg.launch_count() - l1 == l1 - l0 and np.array_equal(r2, r) and same_bits(d2, d)
        # a zero row (cosine: on the always-re-ranked list) keeps its place on that list when rows before it move
        if metric == COSINE:
            s.upsert(int(rr[9000]), np.zeros(dims, dtype="<f4").tobytes())
            vv[9000] = 0
            base = rr.copy()
            for j in range(3):                       # every insert moves the positions behind it up by one
                nid = int(base[100 + 10 * j] + 1)
                s.upsert(nid, q[3 + j].tobytes())
                vv = np.concatenate([vv, q[3 + j:4 + j]])
                rr = np.concatenate([rr, [nid]])
            order = np.argsort(rr, kind="stable")
            vv, rr = vv[order], rr[order]
            r, d, c = s.knn(q, k, metric)
            er, ed, ec = orc.knn(elem, dims, vv, q, k, metric, rowids=rr)
            assert np.array_equal(r, er) and same_bits(d, ed) and np.array_equal(c, ec)


def test_out_of_order_insert_moves_the_tail_in_chunks(vg, orc, gpu, monkeypatch):
    """The tail of the slab moves up by one row through a bounded scratch buffer, from the end backwards (256 MB chunks; 1 MB
    here so that an 11 MB slab needs a dozen): vectors, rowids, tombstones and cached norms all arrive where they belong."""
    monkeypatch.setenv("VECGPU_SHIFT_CHUNK_MB", "1")
    n, dims, nq, k = 30_000, 96, 32, 10
    v = random_rows(F32, n, dims, seed=51)
    q = random_rows(F32, nq, dims, seed=52)
    rr = np.arange(10, 10 + 3 * n, 3, dtype="<i8")
    with vg.Slab(F32, dims) as s:
        s.load(v, rr)
        s.delete(int(rr[777]))
        s.delete(int(rr[n - 2]))
        s.knn(q, k, L2)                                   # the norm cache exists
        live = np.ones(n, dtype=bool)
        live[[777, n - 2]] = False
        vv, ids = v[live], rr[live]
        for j, at in enumerate([0, 1, 12_345, n // 2, n - 3, n - 1]):
            nid = int(rr[at] + 1)
            s.upsert(nid, q[j].tobytes())
            vv, ids = np.concatenate([vv, q[j:j + 1]]), np.concatenate([ids, [nid]])
        order = np.argsort(ids, kind="stable")
        vv, ids = vv[order], ids[order]
        r, d, c = s.knn(q, k, L2)                         # batched: tensor-core path with the moved norms
        er, ed, ec = orc.knn(F32, dims, vv, q, k, L2, rowids=ids)
        assert np.array_equal(r, er) and same_bits(d, ed) and np.array_equal(c, ec)
        r1, d1, _ = s.knn(q[:1], k, L2)                   # single query: the streaming scan
        assert np.array_equal(r1, er[:1]) and same_bits(d1, ed[:1])
        for i in (0, 1, 2, 5000, 12_345, 12_346, n // 2 + 3, len(ids) - 1):
            assert s.get(int(ids[i])) == vv[i].tobytes()
        assert s.count() == (n + 6, n + 4)


def test_out_of_order_inserts_with_tombstones(vg, orc, gpu):
    dims = 12
    rng = np.random.default_rng(9)
    v = random_rows(I8, 300, dims, seed=3, ties=True)
    ids = rng.permutation(np.arange(1, 1201, 4))[:300].astype("<i8")  # arbitrary insertion order
    q = random_rows(I8, 3, dims, seed=4, ties=True)
    with vg.Slab(I8, dims) as s:
        live = {}
        for i, (rid, row) in enumerate(zip(ids, v)):
            s.upsert(int(rid), row.tobytes())
            live[int(rid)] = row
            if i % 7 == 3:  # tombstone something already present, sometimes re-insert it
                victim = int(ids[rng.integers(0, i + 1)])
                s.delete(victim)
                live.pop(victim, None)
            if i % 50 == 49:
                rr = np.array(sorted(live), dtype="<i8")
                vv = np.stack([live[int(x)] for x in rr])
                r, d, c = s.knn(q, 8, L2)
                er, ed, ec = orc.knn(I8, dims, vv, q, 8, L2, rowids=rr)
                assert np.array_equal(r, er) and same_bits(d, ed) and np.array_equal(c, ec)


@pytest.mark.parametrize("dims,n", [(768, 30_000), (5, 1_200_000)])
def test_bulk_load_through_the_staging_buffers(vg, orc, gpu, dims, n):
    """Loads larger than one 16 MB staging buffer (packed rows and rows padded to the 16-byte stride)."""
    v = np.random.default_rng(dims).standard_normal((n, dims)).astype("<f4")
    with vg.Slab(F32, dims) as s:
        s.load(v)
        for rid in (1, 2, n // 3, n - 1, n):
            assert s.get(rid) == v[rid - 1].tobytes()
        q = v[[n // 2]]
        r, d, c = s.knn(q, 5, L2)
        er, ed, ec = orc.knn_select(F32, dims, v, q, 5, L2)
        assert np.array_equal(r, er) and same_bits(d, ed)


def test_hnsw_never_returns_deleted_rows_and_goes_stale_on_refill(vg, orc, gpu):
    dims, n = 32, 4000
    v = random_rows(F32, n, dims, seed=51)
    with vg.Slab(F32, dims) as s:
        s.load(v)
        idx = vg.HnswIndex(s, L2, M=16, ef_construction=100, seed=1)
        idx.rebuild()
        q = v[[100, 200, 300]]
        r, d, c = idx.search(q, 5, ef_search=64)
        assert list(r[:, 0]) == [101, 201, 301]
        for rid in (101, 201):
            s.delete(rid)
        r, d, c = idx.search(q, 5, ef_search=64)
        assert 101 not in r and 201 not in r and r[2, 0] == 301 and np.all(c == 5)
        # what comes back is the exact order of the live rows among the beam
        er, ed, _ = s.knn(q, 5, L2)
        assert np.array_equal(r[2], er[2])
        s.fill_synthetic(seed=2, n=n, kind=1)  # every row replaced: the index must fail loudly, not walk stale positions
        with pytest.raises(vg.InvalidState):
            idx.search(q, 5, ef_search=64)
        idx.close()


@pytest.mark.parametrize("elem,metric", PAIRS, ids=PAIR_IDS)
def test_fused_scan_tail_equals_separate_merge_launch(vg, orc, gpu, elem, metric, monkeypatch):
    """nq <= 8 runs scan + final merge in ONE launch (last CTA done); VECGPU_FUSE_MERGE=0 restores the two-launch form."""
    dims = 200 if elem != BIT else 520
    v = random_rows(elem, 60_000, dims, seed=81, ties=(elem != F32))
    q = random_rows(elem, 8, dims, seed=82, ties=(elem != F32))
    with vg.Slab(elem, dims) as s:
        s.load(v)
        for nq, k in ((1, 10), (3, 1), (8, 33), (1, 100), (2, 700)):
            er, ed, ec = orc.knn_select(elem, dims, v, q[:nq], k, metric)
            l0 = vg.launch_count()
            r, d, c = s.knn(q[:nq], k, metric)
            fused_launches = vg.launch_count() - l0
            monkeypatch.setenv("VECGPU_FUSE_MERGE", "0")
            l0 = vg.launch_count()
            r2, d2, c2 = s.knn(q[:nq], k, metric)
            split_launches = vg.launch_count() - l0
            monkeypatch.delenv("VECGPU_FUSE_MERGE")
            assert np.array_equal(r, er) and same_bits(d, ed) and np.array_equal(c, ec)
            assert np.array_equal(r2, er) and same_bits(d2, ed) and np.array_equal(c2, ec)
            if nq in (1, 8) and k <= 33 and not (elem == F32 and metric == L1):  # one query pass; f32 L1 has its own TMA kernel
                assert fused_launches < split_launches, (nq, k, fused_launches, split_launches)
        # many back-to-back fused launches: the ticket counter must re-arm itself every time
        for _ in range(200):
            r3, d3, _ = s.knn(q[:1], 10, metric)
        er, ed, _ = orc.knn_select(elem, dims, v, q[:1], 10, metric)
        assert np.array_equal(r3, er) and same_bits(d3, ed)


def test_two_stream_pipelining_of_single_query_scans(vg, orc, gpu):
    """Independent queries issued alternately on two streams (per-stream scratch sets) and a third stream that has to take
    over a set: same answers as one stream; a tensor-core batch in the middle runs exclusively."""
    import torch

    dims, n, k = 128, 400_000, 10
    with vg.Slab(F32, dims) as s:
        s.fill_synthetic(seed=12, n=n, kind=1)
        q = orc.synth_rows(F32, 13, 1, 40, dims, 1)
        er, ed, _ = orc.knn_synth(F32, dims, 12, 1, n, 1, q, k, COSINE)
        dq = torch.from_numpy(q).cuda()
        s0, s1, s2 = torch.cuda.current_stream(), torch.cuda.Stream(), torch.cuda.Stream()
        s1.wait_stream(s0)
        s2.wait_stream(s0)
        outs = []
        for j in range(32):
            st = (s0, s1)[j % 2]
            with torch.cuda.stream(st):
                outs.append(s.knn_device(dq[j], k, COSINE, stream=st.cuda_stream))
            if j == 15:  # a batched (tensor-core) call in between, on the second stream
                with torch.cuda.stream(s1):
                    big = s.knn_device(dq[:32], k, COSINE, stream=s1.cuda_stream)
        with torch.cuda.stream(s2):  # a third stream
            late = [s.knn_device(dq[32 + j], k, COSINE, stream=s2.cuda_stream) for j in range(8)]
        torch.cuda.synchronize()
        for j in range(32):
            assert np.array_equal(outs[j][0].cpu().numpy()[0], er[j]) and np.array_equal(bits(outs[j][1].cpu().numpy()[0]), bits(ed[j]))
        assert np.array_equal(big[0].cpu().numpy(), er[:32]) and np.array_equal(bits(big[1].cpu().numpy()), bits(ed[:32]))
        for j in range(8):
            assert np.array_equal(late[j][0].cpu().numpy()[0], er[32 + j])


@pytest.mark.parametrize("elem,metric,dims", [(F32, L2, 384), (F32, COSINE, 96), (I8, L2, 128), (BIT, HAMMING, 256)])
def test_small_table_with_a_long_result_list_scans_with_fewer_ctas(vg, orc, gpu, monkeypatch, elem, metric, dims):
    """A single query with 28 <= k <= 110 on a small table: the scan runs with fewer CTAs so that the fused final merge sorts
    fewer keys (vecgpu.cu, tools/small_table_gx.py).  Whatever the CTA count — the automatic choice, one CTA, one per SM —
    the answer is the oracle's, rowids and distance bits (ties included: duplicated rows)."""
    for n in (300, 10_000, 120_000):
        v = random_rows(elem, n, dims, seed=61)
        v[n // 2:n // 2 + 40] = v[:40]                    # exact duplicates: equal distances, rowid order decides
        q = random_rows(elem, 3, dims, seed=62)
        with vg.Slab(elem, dims) as s:
            s.load(v)
            for k in (27, 28, 50, 100, 110, 111):
                for nq in (1, 3):                          # a small batch shares one pass of the scan: same choice, per query
                    er, ed, ec = orc.knn(elem, dims, v, q[:nq], k, metric)
                    for cap in (None, "1", "148"):
                        if cap is None:
                            monkeypatch.delenv("VECGPU_SCAN_GX", raising=False)
                        else:
                            monkeypatch.setenv("VECGPU_SCAN_GX", cap)
                        r, d, c = s.knn(q[:nq], k, metric)
                        assert np.array_equal(r, er) and same_bits(d, ed) and np.array_equal(c, ec), (n, k, nq, cap)
