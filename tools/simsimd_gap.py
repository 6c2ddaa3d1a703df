#!/usr/bin/env python
"""How far can the choice of SimSIMD back end move a distance or flip a returned rowid?

The reference's f32 L2 / f32 cosine / i8 cosine arithmetic lives in simsimd 6.5.16 (src/distance/scalar.rs:17, :48, :94),
which cannot be built here.  oracle/simsimd_shapes.c restates the accumulation and finishing shapes SimSIMD publishes
for its back ends; this script scores BASELINE.json's cfg1 (10 k x 384, L2 and cosine) and a prefix of cfg2
(rows x 768 gauss, cosine) under every shape and reports, against the canonical order the kernels implement:
  * the largest relative deviation of any distance (north_star tolerance: 1e-5),
  * how many top-10 result lists change (as a set, and as an ordered list) — the "rowid flip rate".
CPU only (test infrastructure).   python tools/simsimd_gap.py [--rows2 200000] [--nq2 16] > profiles/r2_simsimd_gap.txt
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle  # noqa: E402

F32, I8, L2, COSINE = 0, 1, 0, 2


def topk(d, k):
    # (distance, position) ascending == stable sort over ascending rowids
    idx = np.argsort(d, kind="stable")[:k]
    return idx


def gap_f32(vectors, queries, metric, k=10):
    """-> list of dicts, one per supported (acc, fin) shape."""
    rows = []
    base = [oracle.shape_distances_f32(0, 0, vectors, q, metric) for q in queries]
    base_top = [topk(d, k) for d in base]
    fins = range(len(oracle.FIN_SHAPES)) if metric == COSINE else [0]
    for acc in range(len(oracle.ACC_SHAPES)):
        for fin in fins:
            if not oracle.shape_supported(acc, fin):
                continue
            max_rel, set_flips, order_flips, top_rel = 0.0, 0, 0, 0.0
            for qi, q in enumerate(queries):
                d = oracle.shape_distances_f32(acc, fin, vectors, q, metric)
                b = base[qi].astype(np.float64)
                with np.errstate(divide="ignore", invalid="ignore"):
                    rel = np.abs(d.astype(np.float64) - b) / np.maximum(np.abs(b), np.abs(d.astype(np.float64)))
                rel[~np.isfinite(rel)] = 0.0
                max_rel = max(max_rel, float(rel.max()))
                t = topk(d, k)
                top_rel = max(top_rel, float(rel[base_top[qi]].max()))
                if set(t.tolist()) != set(base_top[qi].tolist()):
                    set_flips += 1
                if not np.array_equal(t, base_top[qi]):
                    order_flips += 1
            rows.append({"acc": oracle.ACC_SHAPES[acc], "fin": oracle.FIN_SHAPES[fin] if metric == COSINE else "-",
                         "max_rel": max_rel, "max_rel_in_top_k": top_rel, "set_flips": set_flips, "order_flips": order_flips,
                         "queries": len(queries)})
    return rows


def gap_i8cos(vectors, queries, k=10):
    rows = []
    base = [oracle.shape_distances_i8cos(0, vectors, q) for q in queries]
    base_top = [topk(d, k) for d in base]
    for fin in range(len(oracle.FIN_SHAPES)):
        if not oracle.shape_supported(0, fin):
            continue
        max_rel, set_flips, order_flips = 0.0, 0, 0
        for qi, q in enumerate(queries):
            d = oracle.shape_distances_i8cos(fin, vectors, q)
            b = base[qi].astype(np.float64)
            with np.errstate(divide="ignore", invalid="ignore"):
                rel = np.abs(d.astype(np.float64) - b) / np.maximum(np.abs(b), np.abs(d.astype(np.float64)))
            rel[~np.isfinite(rel)] = 0.0
            max_rel = max(max_rel, float(rel.max()))
            t = topk(d, k)
            set_flips += set(t.tolist()) != set(base_top[qi].tolist())
            order_flips += not np.array_equal(t, base_top[qi])
        rows.append({"acc": "exact-int", "fin": oracle.FIN_SHAPES[fin], "max_rel": max_rel, "max_rel_in_top_k": None,
                     "set_flips": int(set_flips), "order_flips": int(order_flips), "queries": len(queries)})
    return rows


def show(title, rows):
    print(f"\n## {title}")
    print(f"{'accumulation':16s} {'finish':16s} {'max rel dev':>12s} {'in top-10':>12s} {'set flips':>10s} {'order flips':>12s}")
    for r in rows:
        tr = "-" if r["max_rel_in_top_k"] is None else f"{r['max_rel_in_top_k']:.3e}"
        print(f"{r['acc']:16s} {r['fin']:16s} {r['max_rel']:12.3e} {tr:>12s} {r['set_flips']:>6d}/{r['queries']:<3d} {r['order_flips']:>8d}/{r['queries']:<3d}")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows2", type=int, default=200_000)
    ap.add_argument("--nq1", type=int, default=100)
    ap.add_argument("--nq2", type=int, default=16)
    a = ap.parse_args()
    oracle.build()
    print("# SimSIMD back-end shapes vs the canonical order (oracle/simsimd_shapes.c); rows: (accumulation, finish) shape;")
    print("# 'set flips' = queries whose top-10 rowid SET differs from the canonical one, 'order flips' = whose ordered list differs")
    v1 = oracle.synth_rows(F32, 1, 1, 10_000, 384, 0)
    q1 = oracle.synth_rows(F32, 2, 1, a.nq1, 384, 0)
    show(f"cfg1 10000 x f32[384] uniform, L2, k=10, {a.nq1} queries", gap_f32(v1, q1, L2))
    show(f"cfg1 10000 x f32[384] uniform, cosine, k=10, {a.nq1} queries", gap_f32(v1, q1, COSINE))
    v2 = oracle.synth_rows(F32, 3, 1, a.rows2, 768, 1)
    q2 = oracle.synth_rows(F32, 33, 1, a.nq2, 768, 1)
    show(f"cfg2 prefix {a.rows2} x f32[768] gauss, cosine, k=10, {a.nq2} queries", gap_f32(v2, q2, COSINE))
    v3 = oracle.synth_rows(I8, 4, 1, 50_000, 1024, 0)
    q3 = oracle.synth_rows(I8, 78, 1, 16, 1024, 0)
    show("cfg3-shaped 50000 x i8[1024], cosine (i8 cosine has no reference test at all), k=10, 16 queries", gap_i8cos(v3, q3))


if __name__ == "__main__":
    main()
