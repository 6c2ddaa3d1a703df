"""Effect of the cluster size (TMA multicast of the row tiles) on the f32 batched path."""
import os, signal, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sqlite_vec_hnsw_b200 as vg
signal.alarm(200)
n = 10_000_000
s = vg.Slab(0, 768); s.fill_synthetic(seed=3, n=n, kind=1)
q = torch.randn(1024, 768, device="cuda")
ref = None
for nq in (1024, 512):
    for cs in ("8", "4", "2", "1"):
        os.environ["VECGPU_TC_CLUSTER"] = cs
        best = 1e9
        for rep in range(4):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); r, d = s.knn_device(q[:nq], 10, 2); b.record(); torch.cuda.synchronize()
            best = min(best, a.elapsed_time(b))
        if ref is None: ref = (r.clone(), d.clone())
        same = bool(torch.equal(r, ref[0][:nq]) and torch.equal(d, ref[1][:nq]))
        print(f"f32 cos {nq}x10M k=10 cluster={cs}: {best:8.2f} ms  {nq / best * 1e3:8.0f} q/s  {2.0 * nq * n * 768 / best / 1e9:8.1f} TFLOP/s  same={same} tc_stats={vg.tc_stats()}", flush=True)
