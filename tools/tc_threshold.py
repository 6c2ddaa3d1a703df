"""Where does the tensor-core batched path start to pay off? f32[384] L2, nq in {32, 100, 1024}, small tables."""
import os, signal, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sqlite_vec_hnsw_b200 as vg
signal.alarm(250)
for n in (10_000, 30_000, 100_000, 300_000):
    s = vg.Slab(0, 384); s.fill_synthetic(seed=3, n=n, kind=0)
    out = []
    for nq in (32, 100, 1024):
        q = torch.randn(nq, 384, device="cuda")
        row = []
        for tc in ("1", "0"):
            os.environ["VECGPU_TC"] = tc
            s.knn_device(q, 10, 0); torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(5): s.knn_device(q, 10, 0)
            b.record(); torch.cuda.synchronize()
            row.append(a.elapsed_time(b) / 5 * 1e3)
        out.append(f"nq={nq}: tc {row[0]:7.0f} us / cuda-core {row[1]:7.0f} us")
    print(f"n={n:7d}  " + "   ".join(out), flush=True)
    s.close()
