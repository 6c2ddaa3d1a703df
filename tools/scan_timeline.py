#!/usr/bin/env python
"""Per-CTA timeline of one single-query scan (globaltimer stamps written by scan_kernel when the developer hook
vecgpu_debug_scan_timeline is armed): how long the ramp (launch -> pipelines primed), the streaming part, the spread of
the CTAs' finish times (tail) and the fused final merge take.   python tools/scan_timeline.py [rows] [dims]"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sqlite_vec_hnsw_b200 as vg  # noqa: E402

lib = vg.load_library()
cases = [(10_000, 384, 0), (1_250_000, 768, 2), (10_000_000, 768, 2)]
if len(sys.argv) > 2:
    cases = [(int(sys.argv[1]), int(sys.argv[2]), 2)]
for n, dims, metric in cases:
    sl = vg.Slab(0, dims)
    sl.fill_synthetic(seed=3, n=n, kind=1)
    q = torch.randn(dims, device="cuda")
    buf = torch.zeros(148 * 4 + 8, dtype=torch.int64, device="cuda")
    for fuse, dynmode in [(f, d) for d in os.environ.get("TIMELINE_DYN_MODES", "0").split(",") for f in os.environ.get("TIMELINE_FUSE", "1,0").split(",")]:
        os.environ["VECGPU_FUSE_MERGE"] = fuse
        os.environ["VECGPU_SCAN_DYNAMIC"] = dynmode
        for _ in range(3):
            sl.knn_device(q, 10, metric)
        torch.cuda.synchronize()
        rows = []
        for it in range(20):
            buf.zero_()
            lib.vecgpu_debug_scan_timeline(C.c_void_p(buf.data_ptr()))
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            sl.knn_device(q, 10, metric)
            b.record()
            torch.cuda.synchronize()
            lib.vecgpu_debug_scan_timeline(None)
            t = buf.cpu().numpy().astype(np.int64)
            st = t[: 148 * 4].reshape(148, 4)
            st = st[st[:, 0] > 0]
            t0 = st[:, 0].min()
            rows.append([a.elapsed_time(b) * 1e3, (st[:, 0].max() - t0) / 1e3, (st[:, 1].max() - t0) / 1e3, (st[:, 2].min() - t0) / 1e3,
                         (st[:, 2].max() - t0) / 1e3, (st[:, 3].max() - t0) / 1e3, (t[148 * 4] - t0) / 1e3 if t[148 * 4] else float("nan")]
                        + [(t[148 * 4 + i] - t0) / 1e3 if t[148 * 4 + i] else float("nan") for i in (1, 2, 3, 4)])
        r = np.median(np.array(rows), axis=0)
        print(f"{n} x f32[{dims}] fuse={fuse} dyn={dynmode}: event {r[0]:8.1f} us | last CTA start +{r[1]:5.1f} | all primed +{r[2]:5.1f} | first CTA rows done +{r[3]:7.1f} | "
              f"last CTA rows done +{r[4]:7.1f} | last CTA end +{r[5]:7.1f} | fused merge end +{r[6]:7.1f}", flush=True)
    sl.close()
