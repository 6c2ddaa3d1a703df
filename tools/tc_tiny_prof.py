import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sqlite_vec_hnsw_b200 as vg
s = vg.Slab(0, 384); s.fill_synthetic(seed=3, n=10_000, kind=0)
q = torch.randn(32, 384, device="cuda")
for _ in range(3): s.knn_device(q, 10, 0)
torch.cuda.synchronize()
