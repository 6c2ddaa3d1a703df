"""f32 tensor-core batches: CTA pairs (tcgen05 cta_group::2, M = 256) against single CTAs, 1024 queries x N x 768 cosine."""
import os, signal, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sqlite_vec_hnsw_b200 as vg
signal.alarm(170)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
dims = 768
s = vg.Slab(0, dims); s.fill_synthetic(seed=3, n=n, kind=1)
q = torch.randn(1024, dims, device="cuda")
ref = None
for pair in ("0", "1", "0", "1"):
    os.environ["VECGPU_TC_PAIR"] = pair
    r, d = s.knn_device(q, 10, 2); torch.cuda.synchronize()
    if ref is None: ref = (r.clone(), d.clone())
    same = bool(torch.equal(r, ref[0]) and torch.equal(d.view(torch.int32), ref[1].view(torch.int32)))
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5): r, d = s.knn_device(q, 10, 2)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 5
    print(f"pair={pair}: {ms:8.2f} ms per 1024-query batch  {1024 / ms * 1e3:8.0f} q/s  {2.0 * 1024 * n * dims / ms / 1e9:7.1f} TFLOP/s  identical to first={same}  tc_stats={vg.tc_stats()}", flush=True)
