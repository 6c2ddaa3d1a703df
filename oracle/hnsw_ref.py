"""Pure-Python restatement of the reference's HNSW query walk over an EXPORTED graph (small cases only).

TEST INFRASTRUCTURE, NOT PRODUCT (see oracle/__init__.py).  Follows, statement for statement:
  search_layer  src/hnsw/search.rs:340-543   (entry scored first :385-398; pop closest candidate, stop when it is
                farther than the worst result :403-410; neighbours filtered by the visited set before scoring :424-434;
                admission `results.len() < ef || d < worst` :516; trim to ef :528-531; final sort by distance :540)
  search_hnsw   src/hnsw/search.rs:267-335   (ef = max(ef_search, k) :282; greedy ef=1 descent from the entry point's
                level down to 1 :300-323; ef-wide search at level 0; first k).
Distances come from the C oracle (oracle.distance), so the walk is checked with the same arithmetic as the scan.

Tie order: the reference's heaps compare on distance only (MinCandidate / MaxCandidate, search.rs:212-250), which leaves
the order among equal distances to BinaryHeap internals.  The product defines it as (distance, rowid); so does this file.
Parity status: unpinned against the Rust build (no toolchain here); it pins the device kernel and the
lockstep driver against an independent, literal reading of the reference's algorithm.
"""
import heapq


def search_layer(dist_of, neighbors_of, entry, ef, level):
    """-> [(node, distance)] sorted closest first.  dist_of(node) -> float (NaN = node missing);
    neighbors_of(node, level) -> list in stored order."""
    visited = {entry}
    cand = []   # min-heap on (d, node)
    res = []    # max-heap on (d, node), stored negated
    d0 = dist_of(entry)
    if d0 == d0:
        heapq.heappush(cand, (d0, entry))
        heapq.heappush(res, (-d0, -entry))
    while cand:
        d, c = heapq.heappop(cand)
        if res and d > -res[0][0]:
            break
        unvisited = []
        for nb in neighbors_of(c, level):
            if nb not in visited:
                visited.add(nb)
                unvisited.append(nb)
        for nb in unvisited:
            dn = dist_of(nb)
            if dn != dn:
                continue
            if len(res) < ef or dn < -res[0][0]:
                heapq.heappush(cand, (dn, nb))
                heapq.heappush(res, (-dn, -nb))
                while len(res) > ef:
                    heapq.heappop(res)
    out = sorted((-nd, -nn) for nd, nn in res)
    return [(n, d) for d, n in out]


def search_hnsw(dist_of, neighbors_of, entry, entry_level, k, ef_search):
    """-> [(node, distance)] of at most k results, closest first (distances in the graph's internal metric)."""
    if entry is None or k == 0:
        return []
    ef = max(ef_search, k)
    cur = entry
    for level in range(entry_level, 0, -1):
        r = search_layer(dist_of, neighbors_of, cur, 1, level)
        if r:
            cur = r[0][0]
    return search_layer(dist_of, neighbors_of, cur, ef, 0)[:k]


def adjacency_from_edges(fr, to, lv):
    """Exported edge arrays (stored order per list) -> neighbors_of(node, level)."""
    adj = {}
    for f, t, l in zip(fr.tolist(), to.tolist(), lv.tolist()):
        adj.setdefault((f, l), []).append(t)
    return lambda node, level: adj.get((node, level), [])
