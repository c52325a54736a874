#!/usr/bin/env python
"""Per-model step/eval timings on the BASELINE.json parity configs (C1-C4) — secondary numbers, not the bench line.

    python tools/model_bench.py [--models lightgcn,lightgcn_tag,ngcf,dgcf,disengcn,tgcn] [--steps 10] [--profile]

For each model: builds the synthetic graph of its named shape (tagrec_b200.data.SHAPES), composes the drop-in objects
the way com.py does (model, sampler, Adam, Basic_test), times `steps` training steps (CUDA events, after 3 warm-up
steps, L2 flushed between steps because these tables fit in L2) and one full evaluation, and prints one JSON line.
--profile brackets the timed steps with cudaProfilerStart/Stop for `ncu --profile-from-start off`.
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CASES = {
    # name: (model, shape, use_tag, sampler class, config overrides)
    "lightgcn": ("lightgcn", "lastfm", False, "BPR", {}),
    "lightgcn_tag": ("lightgcn", "delicious_tags", True, "BPR", {}),
    "ngcf": ("ngcf", "amazon_book", False, "BPR", {}),
    "dgcf": ("dgcf", "gowalla", False, "DGCF", {}),
    "disengcn": ("disengcn", "delicious_tags", True, "DGCF", {}),
    "tgcn": ("tgcn", "delicious_tags", True, "BPR", {"dim_layer_list": [64, 64]}),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--models", default="lightgcn,lightgcn_tag,ngcf,dgcf,disengcn,tgcn")
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--profile", action="store_true")
    ap.add_argument("--graph", action="store_true", help="record the step into a CUDA graph (T.GraphedStep)")
    ap.add_argument("--kineto", action="store_true", help="print the in-pipeline kernel times of 3 extra steps (CUPTI)")
    args = ap.parse_args()
    import __graft_entry__ as G
    G.build()
    import tagrec_b200 as T
    dev = torch.device("cuda:0")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for name in args.models.split(","):
        model_name, shape, use_tag, sampler, over = CASES[name]
        t0 = time.time()
        ds = T.data.synth_named(shape)
        cfg = dict(use_tag=use_tag, reg=1e-4, dim_latent=64, dim_layer_list=[64, 64, 64], train_batch=2048, device=dev,
                   lr=0.001, sampler="device", topks=[20], test_batch=512)
        cfg.update(over)
        T.set_config(model_name, **cfg)
        torch.manual_seed(2020)
        if model_name == "tgcn":
            ds.get_all_neighbor = lambda ds=ds: T.data.get_all_neighbor(ds, width=25)      # data/tgcn_load.py:41-53, restated
            model = T.TGCN(ds).to(dev)
        else:
            model = {"lightgcn": T.LightGCN, "ngcf": T.NGCF, "dgcf": T.DGCF, "disengcn": T.DisenGCN}[model_name](ds).to(dev)
        data = (T.DGCF_training_data if sampler == "DGCF" else T.BPR_training_data)(ds, None)
        if args.graph:
            opt = T.FusedAdam(model.parameters(), lr=0.001, capturable=True)
            graphed = T.GraphedStep(model, opt, warmup=2)
        else:
            opt = torch.optim.Adam(model.parameters(), lr=0.001)
        test = T.Basic_test(ds, None)
        model.train()
        data.reset()
        batches = []
        for b in data.mini_batch():
            batches.append(b)
            if len(batches) >= args.steps + 3:
                break
        while len(batches) < args.steps + 3:
            batches.append(batches[len(batches) % max(1, len(batches))])
        setup_s = time.time() - t0

        def step(b):
            if args.graph:
                return graphed.loss(b)
            lossx = model.loss(b)
            loss = sum(lossx)
            opt.zero_grad()
            loss.backward()
            opt.step()
            return lossx

        for b in batches[:3]:
            step(b)
        torch.cuda.synchronize()
        l0 = T.launch_count()
        if args.profile:
            torch.cuda.profiler.start()
        evs = []
        for b in batches[3:3 + args.steps]:
            flush.zero_()
            a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            lossx = step(b)
            e.record()
            evs.append((a, e))
        torch.cuda.synchronize()
        if args.profile:
            torch.cuda.profiler.stop()
        ms = float(np.median([a.elapsed_time(e) for a, e in evs]))
        if args.kineto:
            from torch.profiler import ProfilerActivity, profile
            with profile(activities=[ProfilerActivity.CUDA]) as prof:
                for b in batches[3:6]:
                    step(b)
                torch.cuda.synchronize()
            rows = sorted(prof.key_averages(), key=lambda r: -r.device_time_total)
            tot = sum(r.device_time_total for r in rows)
            print(f"# {name}: {tot / 3e3:.2f} ms of kernel time per step (in-pipeline, warm)")
            for r in rows[:24]:
                print(f"#  {r.device_time_total / 3:9.1f} us {r.count / 3:6.1f}x  {100 * r.device_time_total / tot:5.1f} %  {r.key[:90]}")
        launches = (T.launch_count() - l0) / args.steps
        bsz = (batches[0][0] if isinstance(batches[0], tuple) else batches[0]).shape[0]
        # evaluation (K3 + K3b + metrics) through the drop-in Basic_test
        res = test.run(model)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        res = test.run(model)
        torch.cuda.synchronize()
        eval_s = time.perf_counter() - t1
        graph = getattr(model, "norm_adj", None)
        print(json.dumps({
            "case": name, "model": model_name, "shape": shape, "users": ds.num["user"], "items": ds.num["item"],
            "tags": ds.num.get("tag", 0), "nnz": graph._nnz() if graph is not None else None, "batch": int(bsz),
            "cuda_graph": bool(args.graph), "ms_per_step": ms, "triples_per_s": bsz / ms * 1e3, "tagrec_launches_per_step": launches,
            "eval_users": len(ds.user_items["test"]), "eval_s": eval_s,
            "eval_users_per_s": len(ds.user_items["test"]) / eval_s,
            "loss": [float(x) for x in lossx], "ndcg@20": res["ndcg"][0], "auc": res.get("auc", [None])[0],
            "setup_s": round(setup_s, 1)}), flush=True)
        del model, opt, data, test, ds
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
