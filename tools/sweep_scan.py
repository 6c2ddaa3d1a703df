"""Sweep the scan kernel's launch-plan knobs on the GPU and print achieved GB/s.
Usage: python tools/sweep_scan.py [quick]"""
import itertools
import signal
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sqlite_vec_hnsw_b200 as vg  # noqa: E402

F32, I8, BIT = 0, 1, 2
L2, L1, COS, HAM = 0, 1, 2, 3


def time_knn(slab, q, k, metric, iters=6):
    """median per-launch time (ms) over `iters` launches, CUDA events on the launch stream"""
    st = torch.cuda.current_stream()
    for _ in range(2):
        slab.knn_device(q, k, metric)
    torch.cuda.synchronize()
    evs = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(st)
        slab.knn_device(q, k, metric)
        b.record(st)
        evs.append((a, b))
    torch.cuda.synchronize()
    return float(np.median([a.elapsed_time(b) for a, b in evs]))


def main():
    quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
    cfgs = [
        ("f32[768] cos k=10", F32, 768, COS, 10, 4_000_000, 1),
        ("f32[768] l2 k=10", F32, 768, L2, 10, 4_000_000, 1),
        ("f32[768] l1 k=10", F32, 768, L1, 10, 4_000_000, 1),
        ("f32[384] l2 k=10", F32, 384, L2, 10, 8_000_000, 0),
        ("i8[1024] l2 k=100", I8, 1024, L2, 100, 12_000_000, 0),
        ("bit[1024] ham k=10", BIT, 1024, HAM, 10, 64_000_000, 0),
    ]
    knobs = {
        "VECGPU_SCAN_WARPS": ["16"] if quick else ["8", "12", "16"],
        "VECGPU_SCAN_CB": ["2048"],
        "VECGPU_SCAN_STAGE_KB": ["8"] if quick else ["4", "8", "12"],
    }
    results = []
    only = os.environ.get("SWEEP_ONLY")
    for name, elem, dims, metric, k, n, kind in cfgs:
        if only and not any(o in name for o in only.split(",")):
            continue
        slab = vg.Slab(elem, dims)
        slab.fill_synthetic(seed=7, n=n, kind=kind)
        rb = slab.row_bytes
        q = torch.zeros(rb, dtype=torch.uint8, device="cuda")
        q.copy_(torch.from_numpy(np.random.default_rng(1).integers(0, 255, rb).astype("u1")))
        if elem == F32:
            q = torch.randn(dims, device="cuda")
        best = None
        for w, cb, skb in itertools.product(*knobs.values()):
            os.environ["VECGPU_SCAN_WARPS"], os.environ["VECGPU_SCAN_CB"], os.environ["VECGPU_SCAN_STAGE_KB"] = w, cb, skb
            signal.alarm(40)  # a hung kernel must not eat the GPU budget: default SIGALRM action kills us
            try:
                ms = time_knn(slab, q, k, metric, iters=30)
            except Exception as e:  # plan may not fit
                print(f"{name:22s} warps={w} cb={cb} stage_kb={skb}: {e}")
                continue
            gbs = n * rb / ms / 1e6
            rec = dict(cfg=name, warps=int(w), cb=int(cb), stage_kb=int(skb), ms=ms, gbs=gbs)
            results.append(rec)
            print(f"{name:22s} warps={w} cb={cb:>5s} stage_kb={skb:>3s}: {ms:8.3f} ms  {gbs:8.1f} GB/s", flush=True)
            if best is None or gbs > best["gbs"]:
                best = rec
        print(f"BEST {name}: {best}", flush=True)
        # batched: multi-query passes (QB up to 8)
        for nq in (8, 64):
            for key in knobs:
                os.environ.pop(key, None)
            qb = q.repeat(nq, 1) if q.dim() == 1 else q
            signal.alarm(60)
            ms = time_knn(slab, qb.contiguous(), k, metric, iters=2)
            signal.alarm(0)
            print(f"{name:22s} nq={nq}: {ms:8.3f} ms  {nq / ms * 1e3:9.1f} q/s  ({n * rb * (nq / 8) / ms / 1e6:8.1f} GB/s eff)", flush=True)
        slab.close()
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(results, open("gpurun_out/sweep_scan.json", "w"), indent=1)


if __name__ == "__main__":
    main()
