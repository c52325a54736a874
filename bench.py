#!/usr/bin/env python
"""bench.py — BPR training throughput of the LightGCN hot path on B200 (BASELINE.json metric) + live roofline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

A "step" = one mini-batch through the reference's training step: full-graph propagation (3 fused SpMM layers),
fused BPR loss + gradient scatter, backward (1 elementwise + 3 SpMM on A^T) and torch.optim.Adam over the whole
tables — exactly what training/basic_train.py:14-27 does per batch.  Workload at N=1: the 1 B-edge LightGCN
configuration (BASELINE.json configs[4]: 10 M users x 2 M items, ~1 B interactions, dim 64, 3 layers, batch 2048),
which fits one 180 GB B200.

Prints ONE JSON line (see the task contract): value = triples/s with the batch stream resident in HBM;
e2e = the same through the public Python API with every batch coming from pinned host memory and the loss read back
each step; roofline = the dominant kernel (K1 forward SpMM layer) timed live with CUDA events inside the timed
region; cpu_baseline = the oracle's port of the reference step on the host cores, on a bounded sample graph.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: users, items, interactions, note
    "lightgcn_1b": dict(n_user=10_000_000, n_item=2_000_000, n_edge=1_000_000_000),
    "lightgcn_100m": dict(n_user=2_000_000, n_item=400_000, n_edge=100_000_000),
    "lightgcn_10m": dict(n_user=400_000, n_item=80_000, n_edge=10_000_000),
    "amazon_book": dict(n_user=52_643, n_item=91_599, n_edge=2_984_108),
    "lastfm": dict(n_user=1_892, n_item=17_632, n_edge=92_834),
}
DIM, LAYERS, BATCH = 64, 3, 2048
METRIC, UNIT = "bpr_train_edges_per_sec", "triples/s"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "200"], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.25)
            self.proc.terminate()
            self.t.join(timeout=2)

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for n, v in zip(names, r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------------------------
def cpu_reference(shape, steps, warmup, full_nnz, threads=None):
    """The oracle's port of the reference training step on the host cores, on a bounded sample graph of the same
    shape family (same mean user / item degree), extrapolated linearly in nnz to the full workload (the step is
    SpMM-bound: 74-88 % of the reference's CPU step is aten::addmm)."""
    import numpy as np
    import torch
    from oracle import adjacency as OA
    from oracle.train_step import LightGCNStep
    import tagrec_b200 as T

    threads = threads or os.cpu_count()
    torch.set_num_threads(threads)
    ds = T.data.synth_bipartite(shape["n_user"], shape["n_item"], shape["n_edge"], seed=2020)
    e = ds.edge_index["train"]
    U, I = ds.num["user"], ds.num["item"]
    n, rowptr, col, val = OA.creat_adj(U, I, (e[:, 0], e[:, 1]), "bi_norm")
    m = LightGCNStep(U, I, (n, rowptr, col, val), DIM, LAYERS, reg=1e-4)
    rng = np.random.RandomState(0)
    times = []
    for s in range(warmup + steps):
        sel = rng.randint(0, len(e), BATCH)
        batch = torch.from_numpy(np.stack([e[sel, 0], e[sel, 1], rng.randint(0, I, BATCH)], 1))
        t0 = time.perf_counter()
        m.step(batch)
        if s >= warmup:
            times.append(time.perf_counter() - t0)
    t_sample = sum(times) / len(times)
    nnz = int(rowptr[-1])
    scale = full_nnz / nnz
    return {"value": BATCH / (t_sample * scale), "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"oracle/train_step.py (torch-CPU port of the reference step) on a {U}x{I}, nnz={nnz} graph of "
                      f"the same degree profile: {t_sample*1e3:.1f} ms/step measured, scaled x{scale:.1f} by nnz to "
                      f"the full workload",
            "ms_per_step_sample": t_sample * 1e3}, t_sample * scale


def sample_shape(shape, target_edges=3_000_000):
    f = min(1.0, target_edges / shape["n_edge"])
    return dict(n_user=max(1000, int(shape["n_user"] * f)), n_item=max(1000, int(shape["n_item"] * f)),
                n_edge=int(shape["n_edge"] * f))


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path (oracle port; the reference is Python
    and /root/reference is not on the GPU box) — rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    shape = WORKLOADS[args.workload]
    full_nnz = 2 * int(shape["n_edge"] * 0.97)   # expected after per-user de-duplication
    cb, t_full = cpu_reference(sample_shape(shape), max(1, min(args.steps, 5)), max(1, min(args.warmup, 2)), full_nnz)
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": t_full * 1e3, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": dict(config_of(args, shape), optimizer="torch.optim.Adam (com.py:25, as the reference composes it)"),
            "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(_finite(line)))


def config_of(args, shape):
    return {"optimizer": "tagrec_b200.FusedAdam" if getattr(args, "optimizer", "fused") == "fused" else "torch.optim.Adam",
            "workload": f"LightGCN {LAYERS}-layer dim-{DIM} BPR batch {BATCH}, synthetic {shape['n_user']} users x "
                        f"{shape['n_item']} items, ~{shape['n_edge']} interactions ({args.workload}); full-graph "
                        f"propagation fwd+bwd + Adam every step (reference semantics)",
            "batch": BATCH, "dim": DIM, "layers": LAYERS, "l2": "inputs larger than L2 (tables >> 126 MB)"
            if shape["n_edge"] >= 10_000_000 else "L2 flushed between timed steps"}


# ----------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    import __graft_entry__ as G
    G.build()
    import tagrec_b200 as T
    from tagrec_b200 import functional as Fn

    shape = WORKLOADS[args.workload]
    small = shape["n_edge"] < 10_000_000
    t0 = time.time()
    T.set_config("lightgcn", use_tag=False, reg=1e-4, dim_latent=DIM, dim_layer_list=[DIM] * LAYERS, train_batch=BATCH,
                 device=dev, init_device=dev, lr=0.001, sampler="device")
    if world > 1:
        from tagrec_b200 import distributed as D
        model, triples, info = D.build_sharded_lightgcn(shape, dev, rank, world, BATCH * (args.steps + args.warmup),
                                                        eval_users_per_rank=args.eval_users)
    else:
        ui_row, ui_col = T.data.synth_bipartite_device(shape["n_user"], shape["n_item"], int(shape["n_edge"]),
                                                       dev, seed=2020)
        n_train = ui_row.numel()
        graph = T.build_csr(shape["n_user"], shape["n_item"], (ui_row, ui_col), "bi_norm", dev)

        class Data:
            num = {"user": shape["n_user"], "item": shape["n_item"]}
            prebuilt_adj = graph
        torch.manual_seed(2020)
        model = T.LightGCN(Data)
        # batch stream: device sampler over a random subset of the positive edges (all steps' triples)
        need = BATCH * (args.steps + args.warmup)
        g = torch.Generator(device=dev); g.manual_seed(1)
        idx = torch.randint(0, n_train, (need,), device=dev, generator=g)
        edges = torch.stack([ui_row[idx], ui_col[idx]], 1).contiguous()
        del ui_row, ui_col, idx
        train_ptr = graph.rowptr[:shape["n_user"] + 1].contiguous()
        train_items = (graph.col[:n_train] - shape["n_user"]).contiguous()
        triples = torch.empty((need, 3), dtype=torch.int64, device=dev)
        T._lib.check(T._lib.lib().tagrec_sample_bpr_device(T._lib.ptr(edges), need, T._lib.ptr(train_ptr),
                                                           T._lib.ptr(train_items), shape["n_item"], 2020, 0,
                                                           T._lib.ptr(triples), T._lib.stream_ptr(dev)), "sampler")
        del edges, train_items
        info = {"nnz": graph._nnz(), "n": graph.n, "n_long_rows": graph.n_long, "parallelism": "single"}
    # com.py:25 composes optim.Adam(model.parameters(), lr); FusedAdam is this package's drop-in for it (same update
    # rule, one kernel per tensor).  --optimizer torch runs the reference's own choice.
    opt = (T.FusedAdam if args.optimizer == "fused" else torch.optim.Adam)(model.parameters(), lr=0.001)
    model.train()
    torch.cuda.synchronize()
    setup_s = time.time() - t0
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev) if small else None

    def step(batch):
        lossx = model.loss(batch)
        loss = sum(lossx)
        opt.zero_grad()
        loss.backward()
        opt.step()
        return lossx

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    W, K = args.warmup, args.steps
    for s in range(W):
        step(triples[s * BATCH:(s + 1) * BATCH])
    # ---- timed region A: inputs resident in HBM -> "value" ----
    timer = Fn.KernelTimer()
    launches0 = T.launch_count()
    barrier()
    profiling = os.environ.get("TAGREC_PROFILE") == "1"      # ncu --profile-from-start off captures region A only
    if profiling:
        torch.cuda.profiler.start()
    with ClockSampler(local) as clocks:
        Fn.KERNEL_TIMER = timer
        if small:
            evs = []
            for s in range(W, W + K):
                flush.zero_()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); step(triples[s * BATCH:(s + 1) * BATCH]); b.record()
                evs.append((a, b))
            barrier()
            ms = sum(a.elapsed_time(b) for a, b in evs)
        else:
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for s in range(W, W + K):
                step(triples[s * BATCH:(s + 1) * BATCH])
            b.record()
            barrier()
            ms = a.elapsed_time(b)
        Fn.KERNEL_TIMER = None
    if profiling:
        torch.cuda.profiler.stop()
    launches = T.launch_count() - launches0
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = BATCH * K / (ms / 1e3)

    # ---- timed region B: end to end through the public API, batches from pinned host memory, loss read back ----
    host = triples[W * BATCH:(W + K) * BATCH].cpu().pin_memory()
    barrier()
    t1 = time.perf_counter()
    last = None
    for s in range(K):
        if small:
            flush.zero_()
        batch = host[s * BATCH:(s + 1) * BATCH].to(dev, non_blocking=True)
        lossx = step(batch)
        last = [x.cpu().item() for x in lossx]               # basic_train.py:16 — the step's result read back
    barrier()
    e2e_s = time.perf_counter() - t1
    if world > 1:
        t = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e = {"value": BATCH * K / e2e_s, "unit": UNIT, "h2d_bytes_per_step": BATCH * 3 * 8, "d2h_bytes_per_step": 8,
           "last_loss": last}

    eval_sharded = None
    if world > 1 and args.eval_users > 0 and info.get("eval_mask") is not None:
        eval_sharded = eval_leg_sharded(T, model, shape, dev, info["eval_mask"], world, dist)
    per_rank = None
    if world > 1:
        mine = {"rank": rank, "fwd_ms": timer.mean_ms("spmm_fwd"), "bwd_ms": timer.mean_ms("spmm_bwd"),
                "all_gather_ms": timer.mean_ms("all_gather"), "all_gathers_per_step": timer.count("all_gather") // K,
                "barrier_ms": timer.mean_ms("barrier"),
                "nnz": info["nnz"], "rows": info["n"]}
        per_rank = [None] * world
        dist.all_gather_object(per_rank, mine)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    # ---- roofline of the dominant kernel (K1 forward SpMM layer), live events from region A ----
    peak, peak_src = peaks()
    nnz_l, n_l = info["nnz"], info["n"]                                  # per-rank (local rows) figures
    bytes_per_launch = nnz_l * (8 + 4 * DIM) + n_l * (8 + 3 * 4 * DIM)    # 264 B/nnz + 776 B/row at dim 64
    fwd_ms, bwd_ms = timer.mean_ms("spmm_fwd"), timer.mean_ms("spmm_bwd")
    achieved = bytes_per_launch / (fwd_ms * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "spmm_traffic.json")
    if os.path.exists(tp) and world == 1:
        try:
            traffic = json.load(open(tp)).get(args.workload)
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "kernel": "spmm_kernel<16,EPI_FWD> (tagrec_lightgcn_fwd_layer)", "achieved": achieved,
                "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "bytes_per_launch": bytes_per_launch, "ms_per_launch": fwd_ms, "launches_timed": timer.count("spmm_fwd"),
                "bwd_layer_ms": bwd_ms, "bwd_layer_gbs": bytes_per_launch / (bwd_ms * 1e-3) / 1e9,
                "bpr_ms": timer.mean_ms("bpr"), "bwd_elementwise_ms": timer.mean_ms("bwd_elementwise"),
                # the backward launches of the last timed step, in order: G_{L-1} (source non-zero on the batch rows
                # only), G_{L-2} (batch rows + neighbours), ..., dE0 (dense source); zero source rows are skipped
                "bwd_launch_ms": [round(x.elapsed_time(y), 3) for x, y in timer.pairs.get("spmm_bwd", [])[-LAYERS:]],
                "row_mask_ms": timer.mean_ms("row_mask")}
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": dict(config_of(args, shape), parallelism=info["parallelism"],
                                                nnz=info["nnz"], nodes=info["n"], long_rows=info["n_long_rows"]),
            "e2e": e2e, "gpu_launches": launches, "clocks": clocks.summary(), "roofline": roofline,
            "setup_s": round(setup_s, 1)}
    if per_rank:
        line["per_rank"] = per_rank
        line["partition"] = {k: info.get(k) for k in ("bounds", "type_weight_s_per_nnz", "balance_feedback")}
    if world == 1 and not args.no_cpu_baseline:
        cb, _ = cpu_reference(sample_shape(shape), 3, 1, info["nnz"])
        line["cpu_baseline"] = cb
    if args.eval_users > 0 and world == 1:
        line["eval"] = eval_leg(T, model, shape, dev, args.eval_users)
    if eval_sharded is not None:
        line["eval"] = eval_sharded
    print(json.dumps(_finite(line)))
    if world > 1:
        dist.destroy_process_group()


def _finite(x):
    """NaN / inf are not JSON: replace them with null."""
    if isinstance(x, float):
        return x if x == x and abs(x) != float("inf") else None
    if isinstance(x, dict):
        return {k: _finite(v) for k, v in x.items()}
    if isinstance(x, (list, tuple)):
        return [_finite(v) for v in x]
    return x


def eval_leg_sharded(T, model, shape, dev, eval_mask, world, dist):
    """Full-rank evaluation sharded by user batch (BASELINE metric "eval users/sec at 1/2/4/8 GPU"): every rank scores
    its own share of the users (weak scaling: the same number of users per GPU as the 1-GPU leg) against the full item
    table with its users' train rows as masks; time = max over ranks."""
    import torch
    from tagrec_b200.eval_ops import topk_scores
    lo, ptr_l, items_l = eval_mask
    n = ptr_l.numel() - 1
    model.eval()
    with torch.no_grad():
        all_users, all_items = model.forward()[:2]            # collective: every rank calls it
    ut, it = all_users[lo:lo + n].contiguous(), all_items.contiguous()
    users = torch.arange(n, device=dev)
    for _ in range(2):
        topk_scores(users, ut, it, ptr_l, items_l, 20)
    torch.cuda.synchronize()
    dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    ids, _ = topk_scores(users, ut, it, ptr_l, items_l, 20)
    b.record()
    torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b)], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    total = n * world
    return {"users_per_s": total / (ms / 1e3), "users": total, "users_per_gpu": n, "items": shape["n_item"], "k": 20,
            "ms": ms, "tflops": 2.0 * total * shape["n_item"] * DIM / (ms * 1e-3) / 1e12, "scaling": "weak",
            "kernel": "eval_tc_kernel (tcgen05.mma kind::tf32 filter + exact fp32 re-score), users sharded over ranks"}


def eval_leg(T, model, shape, dev, n_users):
    """Secondary metric of BASELINE.json: full-rank eval users/s (K3: scoring + mask + top-20 + metric sums).
    Both scoring paths are timed: the tcgen05 TF32-filter path (default for dim 64) and the exact-fp32 CUDA-core
    path; they return identical lists (checked here on the benchmark inputs)."""
    import torch
    from tagrec_b200.eval_ops import metric_sums, topk_scores
    graph = model.norm_adj
    U = shape["n_user"]
    n_users = min(n_users, U)
    model.eval()
    users = torch.arange(0, n_users, device=dev)
    train_ptr = graph.rowptr[:U + 1].contiguous()
    train_items = (graph.col[:int(train_ptr[-1].item())] - U).contiguous()
    # synthetic ground truth: each user's "test item" is its first train neighbour (exercises the metric kernel)
    test_ptr = torch.arange(0, U + 1, device=dev)
    test_items = train_items[train_ptr[:-1].clamp(max=train_items.numel() - 1)].contiguous()
    with torch.no_grad():
        all_users, all_items = model.forward()[:2]
    all_users, all_items = all_users.contiguous(), all_items.contiguous()
    flops = 2.0 * n_users * shape["n_item"] * DIM
    out, ids_by_path = {}, {}
    for path in ("tf32", "fp32"):
        for _ in range(2):
            ids, _ = topk_scores(users, all_users, all_items, train_ptr, train_items, 20, path=path)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        ids, _ = topk_scores(users, all_users, all_items, train_ptr, train_items, 20, path=path)
        metric_sums(users, ids, test_ptr, test_items, [20])
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b)
        ids_by_path[path] = ids
        out[path] = {"users_per_s": n_users / (ms / 1e3), "ms": ms, "tflops": flops / (ms * 1e-3) / 1e12}
    same = bool(torch.equal(ids_by_path["tf32"], ids_by_path["fp32"]))
    # per-user AUC (training/utils.py:37-45) over the same users: 20 random test items each (a random-init model: AUC
    # 0.5, the worst case for the search among the positives); both dense-pass implementations, sums compared
    from tagrec_b200.eval_ops import auc_sums
    gen = torch.Generator(device=dev).manual_seed(5)
    auc_ptr = torch.clamp(torch.arange(U + 1, device=dev), max=n_users) * 20
    auc_items = torch.randint(0, shape["n_item"], (n_users, 20), device=dev, generator=gen).sort(dim=1).values
    auc_items = auc_items.to(torch.int32).flatten().contiguous()
    auc = {}
    for path in ("tf32", "fp32"):
        auc_sums(users, all_users, all_items, train_ptr, train_items, auc_ptr, auc_items, path=path)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        sums = auc_sums(users, all_users, all_items, train_ptr, train_items, auc_ptr, auc_items, path=path)
        b.record()
        torch.cuda.synchronize()
        auc[path] = (a.elapsed_time(b), sums.cpu().numpy())
    auc_same = bool(auc["tf32"][1][1] == auc["fp32"][1][1] and
                    abs(auc["tf32"][1][0] - auc["fp32"][1][0]) <= 1e-9 * max(1.0, auc["fp32"][1][1]))
    auc_out = {"ms": auc["tf32"][0], "users_per_s": n_users / (auc["tf32"][0] / 1e3),
               "mean_auc": _finite(float(auc["tf32"][1][0] / max(auc["tf32"][1][1], 1.0))),
               "kernel": "auc_tc_kernel (3xTF32 tcgen05.mma + exact band, canonical fp32 re-scores)",
               "fp32_cuda_core_path_ms": auc["fp32"][0], "sums_identical": auc_same}
    tf32_peak = 1100.0      # nominal dense TF32 TFLOP/s (B200_PROFILING.md); no measured TF32 figure in MEASURED_PEAKS
    return {"users_per_s": out["tf32"]["users_per_s"], "users": n_users, "items": shape["n_item"], "k": 20,
            "ms": out["tf32"]["ms"], "tflops": out["tf32"]["tflops"],
            "tensor_frac_of_nominal_tf32": out["tf32"]["tflops"] / tf32_peak,
            "kernel": "eval_tc_kernel (tcgen05.mma kind::tf32 filter + exact fp32 re-score)",
            "fp32_cuda_core_path": out["fp32"], "paths_identical": same, "auc": auc_out}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="lightgcn_1b", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--eval-users", type=int, default=16384)
    ap.add_argument("--optimizer", default="fused", choices=["fused", "torch"])
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
