"""The extra forms of the CPU oracle added for full-size checks (no GPU needed):

  * orc_knn_select (per-thread selection) == orc_knn (the literal "score every row, stable sort, truncate" restatement
    of src/vtab.rs:2594-2620) on all seven pairs, ties / skips / sparse rowids included;
  * orc_knn_synth (rows regenerated on the fly, nothing materialised) == orc_knn over the materialised rows;
  * the AVX-512 build / the hand-vectorised canonical kernels == the portable plain-C loops, bit for bit;
  * round_haz == roundf (half away from zero, src/vector.rs:531 / :570 use f32::round).
"""
import ctypes as C
import math

import numpy as np
import pytest

import oracle as orc_mod
from helpers import BIT, COSINE, F32, HAMMING, I8, L1, L2, PAIR_IDS, PAIRS, random_rows, same_bits


@pytest.mark.parametrize("elem,metric", PAIRS, ids=PAIR_IDS)
@pytest.mark.parametrize("ties", [False, True], ids=["random", "ties"])
def test_select_equals_sort(orc, elem, metric, ties):
    dims = 40 if elem != BIT else 72
    n = 5000
    v = random_rows(elem, n, dims, 11, ties=ties)
    q = random_rows(elem, 5, dims, 12, ties=ties)
    rng = np.random.default_rng(13)
    rowids = np.cumsum(rng.integers(1, 5, size=n)).astype("<i8") - 1000
    skip = (rng.random(n) < 0.1).astype("u1")
    for k in (1, 10, 257):
        for kw in ({}, {"rowids": rowids, "skip": skip}):
            a = orc.knn(elem, dims, v, q, k, metric, **kw)
            b = orc.knn_select(elem, dims, v, q, k, metric, **kw)
            assert np.array_equal(a[0], b[0]) and same_bits(a[1], b[1]) and np.array_equal(a[2], b[2])


def test_select_k_larger_than_rows(orc):
    v = random_rows(F32, 7, 8, 1)
    q = random_rows(F32, 2, 8, 2)
    a = orc.knn(F32, 8, v, q, 20, L2)
    b = orc.knn_select(F32, 8, v, q, 20, L2)
    assert np.array_equal(a[0], b[0]) and same_bits(a[1], b[1]) and np.array_equal(a[2], b[2])


@pytest.mark.parametrize("elem,dims,kind,metric,k", [(F32, 768, 1, COSINE, 10), (F32, 384, 0, L2, 10), (F32, 33, 1, L1, 3),
                                                      (I8, 1024, 0, L2, 100), (I8, 70, 0, COSINE, 5), (BIT, 1024, 0, HAMMING, 10),
                                                      (BIT, 77, 0, HAMMING, 4)])
def test_synth_scan_equals_materialised_scan(orc, elem, dims, kind, metric, k):
    n, first = 20_000, 12345
    rows = orc.synth_rows(elem, 9, first, n, dims, kind)
    q = orc.synth_rows(elem, 10, 1, 3, dims, kind)
    a = orc.knn(elem, dims, rows, q, k, metric, rowids=np.arange(first, first + n, dtype="<i8"))
    b = orc.knn_synth(elem, dims, 9, first, n, kind, q, k, metric)
    assert np.array_equal(a[0], b[0]) and same_bits(a[1], b[1]) and np.array_equal(a[2], b[2])


def test_avx512_build_is_bit_identical_to_the_portable_build(orc):
    import os

    if orc.lib_path() == orc.LIB_PATH:
        pytest.skip("this host has no AVX-512: only the portable build can run")
    port = orc_mod._bind(orc.LIB_PATH)
    fast = orc_mod._bind(orc.LIB_PATH_AVX512)
    for elem, dims, kind in [(F32, 768, 1), (F32, 385, 0), (I8, 1024, 0), (I8, 33, 0), (BIT, 1024, 0), (BIT, 77, 0)]:
        rb = orc.row_bytes(elem, dims)
        a = np.empty(3000 * rb, dtype="u1")
        b = np.empty(3000 * rb, dtype="u1")
        port.orc_synth_rows(elem, 21, 5, 3000, dims, kind, a.ctypes.data_as(C.c_void_p))
        fast.orc_synth_rows(elem, 21, 5, 3000, dims, kind, b.ctypes.data_as(C.c_void_p))
        assert np.array_equal(a, b), (elem, dims, kind)
        for metric in (L2, L1, COSINE) if elem != BIT else (HAMMING,):
            da = np.empty(3000, dtype="<f4")
            db = np.empty(3000, dtype="<f4")
            port.orc_distances(elem, dims, a.ctypes.data_as(C.c_void_p), 3000, a.ctypes.data_as(C.c_void_p), metric, da.ctypes.data_as(C.c_void_p))
            fast.orc_distances(elem, dims, a.ctypes.data_as(C.c_void_p), 3000, a.ctypes.data_as(C.c_void_p), metric, db.ctypes.data_as(C.c_void_p))
            assert same_bits(da, db), (elem, dims, metric)
    x = np.random.default_rng(3).standard_normal((64, 100)).astype("<f4")
    for fn in ("orc_quantize_int8", "orc_quantize_int8_for_index"):
        oa = np.empty((64, 100), dtype="i1")
        ob = np.empty((64, 100), dtype="i1")
        for i in range(64):
            getattr(port, fn)(x[i].ctypes.data_as(C.c_void_p), 100, oa[i].ctypes.data_as(C.c_void_p))
            getattr(fast, fn)(x[i].ctypes.data_as(C.c_void_p), 100, ob[i].ctypes.data_as(C.c_void_p))
        assert np.array_equal(oa, ob)


def test_hand_vectorised_canonical_kernels_equal_the_plain_loops(orc):
    L = orc.lib()
    if not L.orc_has_avx512():
        pytest.skip("no AVX-512 on this host")
    for dims in (1, 15, 16, 17, 384, 768, 1000):
        v = random_rows(F32, 500, dims, dims)
        q = random_rows(F32, 1, dims, dims + 1)[0]
        try:
            L.orc_force_plain(1)
            p2, pc = orc.distances(F32, dims, v, q, L2), orc.distances(F32, dims, v, q, COSINE)
        finally:
            L.orc_force_plain(0)
        assert same_bits(p2, orc.distances(F32, dims, v, q, L2))
        assert same_bits(pc, orc.distances(F32, dims, v, q, COSINE))


def test_round_half_away_from_zero(orc):
    L = orc.lib()
    rng = np.random.default_rng(0)
    xs = np.concatenate([rng.uniform(-130, 130, 50_000).astype("<f4"), (np.arange(-600, 600) / 4).astype("<f4"),
                         np.float32([0.49999997, -0.49999997, 0.5, -0.5, 1.5, -1.5, 2.5, -2.5, 8388607.5, -8388607.5, 1e10, -1e10, 0.0])])
    for x in xs:
        x = float(x)
        want = math.copysign(math.floor(abs(x) + 0.5), x)  # exact in f64 for f32 inputs
        assert L.orc_round_haz(x) == np.float32(want), x
