"""One scan configuration in a fresh process: python tools/one_scan.py WARPS N_ROWS [ITERS] [ELEM DIMS METRIC K]"""
import os
import signal
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sqlite_vec_hnsw_b200 as vg  # noqa: E402

signal.alarm(120)
os.environ["VECGPU_SCAN_WARPS"] = sys.argv[1]
n = int(sys.argv[2])
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 5
elem, dims, metric, k = (int(x) for x in sys.argv[4:8]) if len(sys.argv) > 7 else (0, 768, 2, 10)
slab = vg.Slab(elem, dims)
slab.fill_synthetic(seed=7, n=n, kind=1 if elem == 0 else 0)
q = torch.randn(dims, device="cuda") if elem == 0 else torch.randint(0, 255, (slab.row_bytes,), dtype=torch.uint8, device="cuda")
for i in range(iters):
    r, d = slab.knn_device(q, k, metric)
    torch.cuda.synchronize()
print(f"warps={sys.argv[1]} n={n} elem={elem} dims={dims}: OK {r[0, :3].tolist()}", flush=True)
