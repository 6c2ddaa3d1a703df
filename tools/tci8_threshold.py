"""int8 L2 batches on small tables: tensor-core path vs CUDA-core multi-query scan."""
import os, signal, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sqlite_vec_hnsw_b200 as vg
signal.alarm(250)
for dims in (128, 1024):
  for n in (10_000, 30_000, 100_000, 300_000):
    s = vg.Slab(1, dims); s.fill_synthetic(seed=4, n=n, kind=0)
    out = []
    for nq in (16, 32, 100, 1024):
        q = torch.randint(-128, 127, (nq, dims), dtype=torch.int8, device="cuda")
        row = []
        for tc in ("1", "0"):
            os.environ["VECGPU_TC"] = tc
            s.knn_device(q, 10, 0); torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(5): s.knn_device(q, 10, 0)
            b.record(); torch.cuda.synchronize()
            row.append(a.elapsed_time(b) / 5 * 1e3)
        out.append(f"nq={nq}: tc {row[0]:6.0f} / cc {row[1]:6.0f}")
    print(f"i8[{dims}] n={n:7d}  " + "   ".join(out), flush=True)
    s.close()
