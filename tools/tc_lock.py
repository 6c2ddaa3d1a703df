"""Effect of the soft lockstep (VECGPU_TC_LOCKSTEP) on the batched tensor-core paths."""
import os, signal, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sqlite_vec_hnsw_b200 as vg
signal.alarm(280)
def run(s, q, k, metric, label, flops):
    for lock, slack in (("1", "0"), ("0", "0")):
        os.environ["VECGPU_TC_LOCKSTEP"] = lock
        os.environ["VECGPU_TC_LOCKSLACK"] = slack
        best = 1e9
        for rep in range(4):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); s.knn_device(q, k, metric); b.record(); torch.cuda.synchronize()
            best = min(best, a.elapsed_time(b))
        print(f"{label} lockstep={lock} slack={slack}: {best:8.2f} ms  {q.shape[0] / best * 1e3:8.0f} q/s  {flops / best / 1e9:8.1f} T(FL)OP/s", flush=True)
n = 10_000_000
s = vg.Slab(0, 768); s.fill_synthetic(seed=3, n=n, kind=1)
q = torch.randn(1024, 768, device="cuda")
run(s, q, 10, 2, "f32 cos 1024x10M k=10", 2.0 * 1024 * n * 768)
run(s, q[:512], 10, 2, "f32 cos  512x10M k=10", 2.0 * 512 * n * 768)
s.close()
s = vg.Slab(1, 1024); s.fill_synthetic(seed=4, n=n, kind=0)
q8 = torch.randint(-128, 127, (1024, 1024), dtype=torch.int8, device="cuda")
run(s, q8, 100, 0, "i8 l2 1024x10M k=100", 2.0 * 1024 * n * 1024)
run(s, q8, 10, 0, "i8 l2 1024x10M k=10 ", 2.0 * 1024 * n * 1024)
