#!/usr/bin/env python
"""K3 micro-benchmark: full-sort evaluation (scoring + train-item masking + top-K) on synthetic tables.

    python tools/eval_bench.py [--users 16384] [--items 2000000] [--k 20] [--deg 100] [--paths tf32,fp32] [--reps 5]

Prints one JSON line per path: users/s, ms, TFLOP/s (2*U*I*64 flop), and whether the two paths returned identical
lists.  Used for the ncu capture of eval_tc_kernel (profiles/).  Inputs are larger than L2 (item table 512 MB).
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--users", type=int, default=16384)
    ap.add_argument("--items", type=int, default=2_000_000)
    ap.add_argument("--k", type=int, default=20)
    ap.add_argument("--deg", type=int, default=100)
    ap.add_argument("--paths", default="tf32,fp32")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--dim", type=int, default=64)
    ap.add_argument("--auc", action="store_true", help="also time the per-user AUC pass (K3b)")
    ap.add_argument("--auc-skew", type=int, default=0,
                    help="0: random test items (AUC 0.5, every score lands between positives: worst case); C > 0: each "
                         "user's 25 test items are its best of C random candidates (trained-model-like, AUC -> 1)")
    args = ap.parse_args()
    import __graft_entry__ as G
    G.build()
    import tagrec_b200 as T
    from tagrec_b200.eval_ops import topk_scores
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev)
    g.manual_seed(0)
    U, I = args.users, args.items
    # LightGCN-like tables: mean of unit rows + small ego term => norms ~0.5-1
    ut = torch.nn.functional.normalize(torch.randn(U, args.dim, device=dev, generator=g), dim=1) * 0.8
    it = torch.nn.functional.normalize(torch.randn(I, args.dim, device=dev, generator=g), dim=1) * 0.8
    deg = torch.full((U,), args.deg, dtype=torch.int64, device=dev)
    ptr = torch.zeros(U + 1, dtype=torch.int64, device=dev)
    ptr[1:] = torch.cumsum(deg, 0)
    items = torch.randint(0, I, (U, args.deg), device=dev, generator=g).sort(dim=1).values.to(torch.int32).flatten()
    users = torch.arange(U, device=dev)
    res = {}
    for path in args.paths.split(","):
        for _ in range(2):
            ids, sc = topk_scores(users, ut, it, ptr, items, args.k, path=path)
        torch.cuda.synchronize()
        times = []
        for _ in range(args.reps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            ids, sc = topk_scores(users, ut, it, ptr, items, args.k, path=path)
            b.record()
            torch.cuda.synchronize()
            times.append(a.elapsed_time(b))
        ms = float(np.median(times))
        res[path] = ids
        print(json.dumps({"path": path, "users": U, "items": I, "k": args.k, "ms": ms, "users_per_s": U / ms * 1e3,
                          "dim": args.dim, "tflops": 2.0 * U * I * args.dim / (ms * 1e-3) / 1e12, "launches": T.launch_count()}))
    if len(res) == 2:
        a, b = res.values()
        print(json.dumps({"paths_identical": bool(torch.equal(a, b))}))
    if args.auc:
        from tagrec_b200.eval_ops import auc_sums
        n_test = 25
        tptr = torch.arange(0, (U + 1) * n_test, n_test, device=dev, dtype=torch.int64)
        if args.auc_skew > 0:
            cand = torch.randperm(I, device=dev, generator=g)[:args.auc_skew]
            best = (ut[users] @ it[cand].T).topk(n_test, dim=1).indices
            titems = cand[best].sort(dim=1).values.to(torch.int32).flatten()
        else:
            titems = torch.randint(0, I, (U, n_test), device=dev, generator=g).sort(dim=1).values.to(torch.int32).flatten()
        outs = {}
        for apath in (["tf32", "fp32"] if args.dim == 64 else ["fp32"]):
            for _ in range(2):
                out = auc_sums(users, ut, it, ptr, items, tptr, titems, path=apath)
            torch.cuda.synchronize()
            times = []
            for _ in range(args.reps):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                out = auc_sums(users, ut, it, ptr, items, tptr, titems, path=apath)
                b.record()
                torch.cuda.synchronize()
                times.append(a.elapsed_time(b))
            ms = float(np.median(times))
            outs[apath] = out
            print(json.dumps({"path": "auc_" + apath, "users": U, "items": I, "dim": args.dim, "ms": ms,
                              "users_per_s": U / ms * 1e3, "tflops": 2.0 * U * I * args.dim / (ms * 1e-3) / 1e12,
                              "mean_auc": float(out[0] / out[1])}))
        if len(outs) == 2:
            d = (outs["tf32"] - outs["fp32"]).abs().cpu().numpy()
            print(json.dumps({"auc_paths_abs_diff_of_sum": float(d[0]), "users_counted_equal": bool(d[1] == 0)}))

if __name__ == "__main__":
    main()
