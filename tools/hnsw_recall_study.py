#!/usr/bin/env python
"""Recall of the reference's OWN insertion procedure (strictly sequential, oracle/hnsw_seq.c) versus the product's
batched GPU build, on BASELINE cfg5's data: f32[384] i.i.d. bell-shaped values (seed 6), L2, M=16, ef_construction=200.

    python tools/hnsw_recall_study.py cpu  100000 [nq]     # sequential CPU build (no GPU needed), recall@10 for several ef
    python tools/hnsw_recall_study.py gpu  100000 [nq]     # product build at batch 1 / 256 / 16384 (needs a GPU), same queries

Both print one JSON line per configuration; profiles/r2_hnsw_recall_*.txt hold the outputs."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle  # noqa: E402

DIMS, M, EFC, SEED, QSEED, K = 384, 16, 200, 6, 67, 10
EFS = [10, 50, 200, 500, 1000, 2000]


def ground_truth(n, nq):
    q = oracle.synth_rows(0, QSEED, 1, nq, DIMS, 1)
    er, _, _ = oracle.knn_synth(0, DIMS, SEED, 1, n, 1, q, K, 0)
    return q, er


def recall(found_rowids, truth):
    return float(np.mean([len(set(a.tolist()) & set(b.tolist())) / K for a, b in zip(found_rowids, truth)]))


def main():
    mode, n = sys.argv[1], int(sys.argv[2])
    nq = int(sys.argv[3]) if len(sys.argv) > 3 else 200
    oracle.build()
    q, truth = ground_truth(n, nq)
    if mode == "cpu":
        v = oracle.synth_rows(0, SEED, 1, n, DIMS, 1)
        for quirk in (False,) + ((True,) if n <= 200_000 else ()):
            h = oracle.HnswSeq(0, DIMS, 0, v, M=M, ef_construction=EFC, quirk=quirk)
            t0 = time.time()
            h.build(oracle.HnswSeq.levels(1, n, M))
            tb = time.time() - t0
            info = h.info()
            edges = int(h.export()[0].size)
            for ef in EFS:
                r, _ = h.search(q, K, ef)
                print(json.dumps({"build": "sequential (oracle/hnsw_seq.c, one insert at a time)", "upper_layer_entry_quirk": quirk, "rows": n,
                                  "M": M, "efc": EFC, "ef_search": ef, "recall_at_10": recall(r + 1, truth), "queries": nq,
                                  "build_s": tb, "edges": edges, "build_distances": info["distances"]}), flush=True)
            print(json.dumps({"rows": n, "expansion_batch_histogram_build_and_queries": info["batch_hist"], "fetches": info["fetches"]}), flush=True)
            h.close()
        return
    import sqlite_vec_hnsw_b200 as vg

    sl = vg.Slab(0, DIMS)
    sl.fill_synthetic(seed=SEED, n=n, kind=1)
    for batch in ([1] if n <= 100_000 else []) + [256, 4096, 16384]:
        idx = vg.HnswIndex(sl, 0, M=M, ef_construction=EFC, seed=1)
        t0 = time.time()
        idx.rebuild(batch=batch)
        tb = time.time() - t0
        st = idx.stats()
        for ef in EFS:
            r, _, _ = idx.search(q, K, ef_search=ef)
            print(json.dumps({"build": f"product, insert batches of <= {batch} (never more than a quarter of the graph)", "rows": n, "M": M, "efc": EFC,
                              "ef_search": ef, "recall_at_10": recall(r, truth), "queries": nq, "build_s": tb, "edges": st["edges"],
                              "build_distances": st["distances_scored"]}), flush=True)
        idx.close()
    sl.close()


if __name__ == "__main__":
    main()
