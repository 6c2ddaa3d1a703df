"""Generates tests/golden/*.npz + cases.json from the CPU oracle.

Run from the repo root:  python tests/golden/make_golden.py
The reference itself cannot run here (no Rust toolchain, simsimd not vendored),
so the fixtures are outputs of oracle/vecgpu_oracle.c on seeded inputs plus the
closed-form inputs of the reference's own tests (cited per case).  They pin the
oracle against regressions and give the GPU tests an oracle-independent target.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle  # noqa: E402
from helpers import BIT, COSINE, F32, HAMMING, I8, L1, L2, random_rows  # noqa: E402

cases = []


def emit(name, elem, dims, metric, k, vectors, queries, rowids=None, skip=None, note=""):
    n = vectors.shape[0]
    rowids = np.arange(1, n + 1, dtype="<i8") if rowids is None else np.asarray(rowids, dtype="<i8")
    r, d, c = oracle.knn(elem, dims, vectors, queries, k, metric, rowids=rowids, skip=skip)
    arrays = dict(vectors=vectors, queries=queries, rowids=rowids, out_rowids=r, out_dists=d, out_counts=c)
    if skip is not None:
        arrays["skip"] = np.asarray(skip, dtype="u1")
    fn = f"{name}.npz"
    np.savez_compressed(os.path.join(HERE, fn), **arrays)
    cases.append(dict(name=name, file=fn, elem=elem, dims=dims, metric=metric, k=k, note=note))


# seeded cases, one per supported (type, metric) pair; sparse rowids + skipped rows
spec = [
    ("f32_l2_d33", F32, 33, L2), ("f32_l1_d33", F32, 33, L1), ("f32_cos_d128", F32, 128, COSINE),
    ("i8_l2_d64", I8, 64, L2), ("i8_l1_d64", I8, 64, L1), ("i8_cos_d70", I8, 70, COSINE),
    ("bit_hamming_d100", BIT, 100, HAMMING),
]
for i, (name, elem, dims, metric) in enumerate(spec):
    v = random_rows(elem, 300, dims, seed=100 + i)
    q = random_rows(elem, 3, dims, seed=200 + i)
    rowids = np.cumsum(np.random.default_rng(300 + i).integers(1, 5, size=300)).astype("<i8")
    skip = np.zeros(300, dtype="u1")
    skip[[0, 17, 299]] = 1
    emit(name, elem, dims, metric, 10, v, q, rowids=rowids, skip=skip)

# heavy ties: rowid tie-break decides almost every rank
emit("f32_l2_ties", F32, 8, L2, 25, random_rows(F32, 400, 8, 1, ties=True), random_rows(F32, 2, 8, 2, ties=True))
emit("i8_l1_ties", I8, 8, L1, 25, random_rows(I8, 400, 8, 3, ties=True), random_rows(I8, 2, 8, 4, ties=True))
emit("bit_hamming_ties", BIT, 24, HAMMING, 25, random_rows(BIT, 400, 24, 5, ties=True), random_rows(BIT, 2, 24, 6, ties=True))

# k larger than the table
emit("f32_cos_k_gt_n", F32, 16, COSINE, 12, random_rows(F32, 7, 16, 7), random_rows(F32, 1, 16, 8))

# reference tests' own closed-form inputs
n, dims = 1000, 128
i = np.arange(n, dtype=np.int64)[:, None]
j = np.arange(dims, dtype=np.int64)[None, :]
v = ((i * 100 + j).astype("<f4") / np.float32(1000.0)).astype("<f4")
emit("ref_recall_l2_1000x128", F32, dims, L2, 10, v, np.full((1, dims), 0.5, dtype="<f4"),
     note="tests/test_recall_accuracy.rs:28-44,79-94")
n, dims = 100, 128
i = np.arange(n, dtype=np.int64)[:, None]
v = ((((7 * i + 13 * j) % 100).astype("<f4")) / np.float32(100.0)).astype("<f4")
emit("ref_recall_cos_100x128", F32, dims, COSINE, 10, v, v[:2].copy(), note="tests/test_recall_cosine.rs:15-125")
emit("ref_knn_simple", F32, 3, COSINE, 2, np.eye(3, dtype="<f4"), np.eye(3, dtype="<f4")[:1],
     note="tests/test_knn_simple.rs:34-53")
v = np.array([[a, a + 1, a + 2] for a in range(1, 6)], dtype="<f4")
emit("ref_integration_rows", F32, 3, COSINE, 3, v, v[:1].copy(), note="tests/integration_test.rs:635-678")

with open(os.path.join(HERE, "cases.json"), "w") as f:
    json.dump(cases, f, indent=1)
print(f"wrote {len(cases)} cases")
