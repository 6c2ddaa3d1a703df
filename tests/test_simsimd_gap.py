"""Bounds the part of the parity claim that cannot be pinned: f32 L2, f32 cosine and i8 cosine are computed inside
simsimd 6.5.16 in the reference (src/distance/scalar.rs:17, :48, :94), which cannot be built here, so the kernels follow a
canonical order.  oracle/simsimd_shapes.c restates the accumulation / finish shapes SimSIMD publishes
for its back ends (serial, 8-lane with f64 reduce, 16-lane with hadd reduce, 4-lane, f64; IEEE / rsqrt + one Newton
step finishes).  This test asserts that NO such shape moves any distance of BASELINE cfg1 / a cfg2 prefix by more
than the 1e-5 relative tolerance north_star states, and reports how many top-10 lists would change (the full-size
report is tools/simsimd_gap.py -> profiles/r2_simsimd_gap.txt).  CPU only."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
import simsimd_gap  # noqa: E402

F32, I8, L2, COSINE = 0, 1, 0, 2
TOL = 1e-5  # BASELINE.json north_star: "float distances must be within 1e-5 relative error"


def test_shapes_reproduce_the_canonical_oracle(orc):
    # shape (canonical, ieee_f64) of simsimd_shapes.c is an independent restatement of the oracle's order: bit-identical
    v = orc.synth_rows(F32, 1, 1, 2000, 384, 0)
    q = orc.synth_rows(F32, 2, 1, 2, 384, 0)
    for metric in (L2, COSINE):
        for qi in range(2):
            a = orc.shape_distances_f32(0, 0, v, q[qi], metric)
            b = orc.distances(F32, 384, v, q[qi], metric)
            assert np.array_equal(a.view("<u4"), b.view("<u4"))
    vi = orc.synth_rows(I8, 4, 1, 2000, 1024, 0)
    qi8 = orc.synth_rows(I8, 78, 1, 1, 1024, 0)[0]
    assert np.array_equal(orc.shape_distances_i8cos(0, vi, qi8).view("<u4"), orc.distances(I8, 1024, vi, qi8, COSINE).view("<u4"))


def test_cfg1_distances_within_tolerance_under_every_shape(orc):
    v = orc.synth_rows(F32, 1, 1, 10_000, 384, 0)
    q = orc.synth_rows(F32, 2, 1, 6, 384, 0)
    for metric in (L2, COSINE):
        rows = simsimd_gap.gap_f32(v, q, metric)
        assert len(rows) >= 7
        worst = max(r["max_rel"] for r in rows)
        flips = sum(r["order_flips"] for r in rows)
        print(f"cfg1 metric={metric}: worst relative deviation {worst:.3e} over {len(rows)} shapes, {flips} changed top-10 lists")
        assert worst <= TOL


def test_cfg2_prefix_distances_within_tolerance_under_every_shape(orc):
    v = orc.synth_rows(F32, 3, 1, 30_000, 768, 1)
    q = orc.synth_rows(F32, 33, 1, 3, 768, 1)
    rows = simsimd_gap.gap_f32(v, q, COSINE)
    worst = max(r["max_rel"] for r in rows)
    flips = sum(r["order_flips"] for r in rows)
    print(f"cfg2 prefix: worst relative deviation {worst:.3e} over {len(rows)} shapes, {flips} changed top-10 lists")
    assert worst <= TOL


def test_i8_cosine_within_tolerance_under_every_finish(orc):
    v = orc.synth_rows(I8, 4, 1, 20_000, 1024, 0)
    q = orc.synth_rows(I8, 78, 1, 4, 1024, 0)
    rows = simsimd_gap.gap_i8cos(v, q)
    assert max(r["max_rel"] for r in rows) <= TOL
