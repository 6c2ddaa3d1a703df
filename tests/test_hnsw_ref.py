"""CPU checks of oracle/hnsw_ref.py (the literal restatement of src/hnsw/search.rs:267-543 used to pin the device walk)."""
import numpy as np

import oracle as orc
from helpers import F32, L2
from oracle import hnsw_ref


def _setup(n=60, dims=6, seed=3):
    rng = np.random.default_rng(seed)
    v = rng.standard_normal((n, dims)).astype("<f4")
    q = rng.standard_normal(dims).astype("<f4")
    return v, q, (lambda node: float(orc.distance(F32, q, v[node], L2)))


def test_complete_graph_gives_exact_topk():
    v, q, dist_of = _setup()
    n = len(v)
    nbrs = lambda node, level: [j for j in range(n) if j != node]
    got = hnsw_ref.search_hnsw(dist_of, nbrs, 0, 0, 5, 8)
    want = sorted(range(n), key=lambda j: (dist_of(j), j))[:5]
    assert [g[0] for g in got] == want
    assert [g[1] for g in got] == [dist_of(j) for j in want]


def test_stop_rule_and_ef_one_descent():
    # a path graph 0-1-2-...: greedy ef=1 walks towards the closest node and stops at the first local minimum
    pts = np.array([[float(i)] for i in range(10)], dtype="<f4")
    q = np.array([6.2], dtype="<f4")
    dist_of = lambda node: float(orc.distance(F32, q, pts[node], L2))
    nbrs = lambda node, level: [j for j in (node - 1, node + 1) if 0 <= j < 10]
    r = hnsw_ref.search_layer(dist_of, nbrs, 0, 1, 0)
    assert [x[0] for x in r] == [6]
    r = hnsw_ref.search_layer(dist_of, nbrs, 0, 3, 0)
    assert [x[0] for x in r] == [6, 7, 5]
    # upper levels only move the entry point (search.rs:300-323)
    upper = lambda node, level: nbrs(node, level) if level == 0 else ([9] if node == 0 else [0])
    got = hnsw_ref.search_hnsw(dist_of, upper, 0, 1, 2, 2)
    assert [x[0] for x in got] == [6, 7]


def test_nan_nodes_never_enter_the_heaps_and_empty_index():
    v, q, dist_of = _setup(n=20)
    nan_of = lambda node: float("nan") if node % 3 == 0 and node else dist_of(node)
    nbrs = lambda node, level: [j for j in range(20) if j != node]
    got = hnsw_ref.search_hnsw(nan_of, nbrs, 0, 0, 20, 20)
    assert all(g[0] % 3 != 0 or g[0] == 0 for g in got) and len(got) == 20 - 6
    assert hnsw_ref.search_hnsw(dist_of, nbrs, None, -1, 5, 5) == []


def test_adjacency_from_edges_keeps_stored_order():
    fr = np.array([5, 5, 7, 5], dtype="<i8")
    to = np.array([9, 2, 1, 4], dtype="<i8")
    lv = np.array([0, 0, 0, 1], dtype="<i4")
    f = hnsw_ref.adjacency_from_edges(fr, to, lv)
    assert f(5, 0) == [9, 2] and f(5, 1) == [4] and f(7, 0) == [1] and f(8, 0) == []
