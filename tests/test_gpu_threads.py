"""Threading contract of the boundary: SQLite calls one connection from one thread at a time, but several
connections on several threads may work on their own tables — or the same one — concurrently
(tests/test_multithread_stress.rs:88-105, 329-342).  Handles carry an internal mutex and their own stream; every entry point
must be callable from different threads on different handles, and from different threads on ONE handle.  ctypes releases
the GIL around the C calls, so these threads really overlap inside the library."""
import threading

import numpy as np
import pytest

from helpers import BIT, COSINE, F32, HAMMING, I8, L2, random_rows, same_bits

pytestmark = pytest.mark.gpu


def _run(threads):
    errs = []

    def wrap(fn):
        def go():
            try:
                fn()
            except BaseException as e:  # noqa: BLE001 - reported by the main thread
                errs.append(e)
        return go

    ts = [threading.Thread(target=wrap(fn)) for fn in threads]
    for t in ts:
        t.start()
    for t in ts:
        t.join(timeout=120)
    assert not any(t.is_alive() for t in ts), "a worker is stuck"
    if errs:
        raise errs[0]


def test_threads_on_their_own_slabs(vg, orc, gpu):
    """Four 'connections', each with its own table of another type: load, delete, upsert, exact KNN (single queries, small and
    tensor-core-sized batches) and candidate scoring, all at once; every answer equals the oracle's."""
    cases = [(F32, 96, COSINE), (I8, 128, L2), (BIT, 256, HAMMING), (F32, 40, L2)]
    results = [None] * len(cases)

    def worker(i):
        elem, dims, metric = cases[i]
        n = 6000 + 500 * i
        v = random_rows(elem, n, dims, seed=100 + i)
        q = random_rows(elem, 40, dims, seed=200 + i)
        skip = np.zeros(n, dtype="u1")
        with vg.Slab(elem, dims) as s:
            s.load(v)
            for p in (3, 77, 4096):
                s.delete(p + 1)
                skip[p] = 1
            v[10] = random_rows(elem, 1, dims, seed=300 + i)[0]
            s.upsert(11, v[10].tobytes())
            out = []
            for rep in range(6):
                out.append(s.knn(q[rep], 10, metric))          # single queries
                out.append(s.knn(q[:5], 7, metric))            # multi-query scan
                out.append(s.knn(q, 10, metric))               # 40 queries: tensor-core / batched paths
                ids = np.arange(1, 65, dtype="<i8") + rep
                out.append(s.score(q[rep], ids, np.array([0, 64], dtype="<u4"), metric))
            results[i] = (v, q, skip, out)

    _run([lambda i=i: worker(i) for i in range(len(cases))])
    for i, (elem, dims, metric) in enumerate(cases):
        v, q, skip, out = results[i]
        for rep in range(6):
            for got, qq, k in ((out[4 * rep], q[rep:rep + 1], 10), (out[4 * rep + 1], q[:5], 7), (out[4 * rep + 2], q, 10)):
                er, ed, ec = orc.knn(elem, dims, v, qq, k, metric, skip=skip)
                assert np.array_equal(got[0].reshape(er.shape), er) and same_bits(got[1].reshape(ed.shape), ed)
            ids = np.arange(1, 65) + rep
            want = np.array([np.nan if skip[j - 1] else orc.distance(elem, q[rep], v[j - 1], metric) for j in ids], dtype="<f4")
            assert same_bits(out[4 * rep + 3], want)


def test_threads_on_one_slab(vg, orc, gpu):
    """Readers and a writer on ONE handle: four threads query while a fifth appends rows that cannot enter any top-k (far away) and
    deletes them again.  Calls serialise on the slab's mutex; every reader sees exactly the serial answers."""
    dims, n = 64, 8000
    v = random_rows(F32, n, dims, seed=7)
    q = random_rows(F32, 64, dims, seed=8)
    with vg.Slab(F32, dims) as s:
        s.load(v)
        want = s.knn(q, 10, L2)
        er, ed, _ = orc.knn(F32, dims, v, q, 10, L2)
        assert np.array_equal(want[0], er) and same_bits(want[1], ed)
        got = [None] * 4
        stop = threading.Event()

        def reader(t):
            acc = []
            for rep in range(12):
                lo = (t * 16 + rep) % 48
                acc.append((lo, s.knn(q[lo:lo + 16], 10, L2), s.knn(q[lo], 10, L2)))
            got[t] = acc

        def writer():
            far = np.full((1, dims), 1.0e6, dtype="<f4")
            rid = n + 1
            while not stop.is_set() and rid < n + 400:
                s.upsert(rid, far.tobytes())
                s.delete(rid)
                rid += 1

        def readers_then_stop():
            _run([lambda t=t: reader(t) for t in range(4)])
            stop.set()

        _run([readers_then_stop, writer])
        for t in range(4):
            for lo, batch, single in got[t]:
                assert np.array_equal(batch[0], want[0][lo:lo + 16]) and same_bits(batch[1], want[1][lo:lo + 16])
                assert np.array_equal(single[0].reshape(-1), want[0][lo]) and same_bits(single[1].reshape(-1), want[1][lo])
        rows, live = s.count()
        assert live == n


def test_threads_on_one_hnsw_index(vg, orc, gpu):
    """Concurrent searches of one index — single queries (one CTA per query) and batches (one warp per query past the
    crossover) — return what a serial caller gets."""
    dims, n = 48, 20_000
    v = orc.synth_rows(F32, 6, 1, n, dims, 1)
    q = orc.synth_rows(F32, 7, 1, 1200, dims, 1)
    with vg.Slab(F32, dims) as s:
        s.load(v)
        idx = vg.HnswIndex(s, L2, M=16, ef_construction=100, seed=3)
        idx.rebuild()
        want_all = idx.search(q, 10, ef_search=80)          # 1200 queries: one warp per query
        want_one = [idx.search(q[i], 10, ef_search=80) for i in range(8)]
        for i in range(8):
            assert np.array_equal(want_one[i][0][0], want_all[0][i]) and same_bits(want_one[i][1][0], want_all[1][i])
        got = [None] * 4

        def worker(t):
            acc = []
            for rep in range(6):
                i = (t + rep) % 8
                acc.append((i, idx.search(q[i], 10, ef_search=80), idx.search(q, 10, ef_search=80) if rep % 3 == 0 else None))
            got[t] = acc

        _run([lambda t=t: worker(t) for t in range(4)])
        for t in range(4):
            for i, one, full in got[t]:
                assert np.array_equal(one[0][0], want_all[0][i]) and same_bits(one[1][0], want_all[1][i])
                if full is not None:
                    assert np.array_equal(full[0], want_all[0]) and same_bits(full[1], want_all[1])
        idx.close()
