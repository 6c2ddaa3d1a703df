import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sqlite_vec_hnsw_b200 as vg
for n in (10_000, 100_000, 1_000_000):
    s = vg.Slab(0, 384); s.fill_synthetic(seed=3, n=n, kind=1)
    q = torch.randn(384, device="cuda")
    for _ in range(3): s.knn_device(q, 10, 2)
    torch.cuda.synchronize(); s.close()
