#!/usr/bin/env python
"""bench.py — BPR training throughput of the graph-embedding hot path on B200 (BASELINE.json metric) + live roofline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

A "step" = one mini-batch through the reference's training step (training/basic_train.py:14-27): full-graph
propagation forward, fused BPR loss + gradient scatter, backward, Adam over all parameters.

Workloads (BASELINE.json configs):
  lightgcn_1b (default, configs[4])  LightGCN 3-layer dim-64 batch 2048 on a synthetic 10 M x 2 M graph, ~1 B
                                     interactions, built and sampled on the device; fits one 180 GB B200.  N > 1: node-range
                                     row blocks, strong scaling.  Also lightgcn_100m / lightgcn_10m (same family).
  lastfm (configs[0])                LightGCN on the LastFM-shaped graph — the reference's own CPU-runnable case.
  delicious_tags_tgcn (configs[1])   TGCN (k = 25, 2 layers) on the Delicious-tags-shaped tripartite graph.
  amazon_book_ngcf (configs[2])      NGCF [64,64,64] on the Amazon-book-shaped graph.
  gowalla_dgcf (configs[3])          DGCF 4 intents x 2 routing iterations on the Gowalla-shaped graph.
The four named shapes go through the drop-in classes end to end (sampler.reset(), model.loss, optimizer,
Basic_test.run) and BOTH arms run the FULL configuration.

ONE JSON line: value = triples/s with the batch stream resident in HBM; e2e = the same through the public Python API with
every batch coming from pinned host memory and the loss read back each step; check = last loss + parameter checksums
after all steps (comparable across N: every rank count trains the same batches); roofline = the dominant kernel (K1
forward SpMM layer) timed live with CUDA events inside the timed region; cpu_baseline = the reference's CPU path on the
box's host cores.  The default line also carries "c1": both arms on the full LastFM-shaped config in the same run.

--impl reference: the reference's CPU implementation of the path on the host cores — the UNMODIFIED reference classes
when a reference tree is present (baseline/_ref or /root/reference: build container only), else oracle/train_step.py's
op-for-op torch-CPU port (the GPU box has no reference tree).  Its line reports exactly what ran: `steps` steps of
`ms_per_step` each on the graph named in cpu_baseline.sample; for lightgcn_* that graph is a bounded sample of the same
degree profile and `value` is scaled by nnz ("extrapolated": true, "scale") — the reference cannot build the 1 B-edge
graph at all (scipy lil_matrix).
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

SCALE_WORKLOADS = {
    "lightgcn_1b": dict(n_user=10_000_000, n_item=2_000_000, n_edge=1_000_000_000),
    "lightgcn_100m": dict(n_user=2_000_000, n_item=400_000, n_edge=100_000_000),
    "lightgcn_10m": dict(n_user=400_000, n_item=80_000, n_edge=10_000_000),
}
NAMED_WORKLOADS = {
    # name: (model, shape in tagrec_b200.data.SHAPES, use_tag, sampler, config overrides)
    "lastfm": ("lightgcn", "lastfm", False, "BPR", {}),
    "delicious_tags_tgcn": ("tgcn", "delicious_tags", True, "BPR", {"dim_layer_list": [64, 64], "neighbor_k": 25}),
    "amazon_book_ngcf": ("ngcf", "amazon_book", False, "BPR", {}),
    "gowalla_dgcf": ("dgcf", "gowalla", False, "DGCF", {}),
}
DIM, LAYERS, BATCH = 64, 3, 2048
METRIC, UNIT = "bpr_train_edges_per_sec", "triples/s"
LR, REG = 0.001, 1e-4


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


def _finite(x):
    """NaN / inf are not JSON: replace them with null."""
    if isinstance(x, float):
        return x if x == x and abs(x) != float("inf") else None
    if isinstance(x, dict):
        return {k: _finite(v) for k, v in x.items()}
    if isinstance(x, (list, tuple)):
        return [_finite(v) for v in x]
    return x


def workload_text(name):
    if name in SCALE_WORKLOADS:
        s = SCALE_WORKLOADS[name]
        return (f"LightGCN {LAYERS}-layer dim-{DIM} BPR batch {BATCH}, synthetic {s['n_user']} users x {s['n_item']} items, "
                f"~{s['n_edge']} interactions ({name}); one reference training step per batch (model.loss: propagation over the whole "
                f"graph, BPR loss; backward; Adam over all parameters) with the reference's loss, gradients and parameters")
    model, shape, use_tag, _, over = NAMED_WORKLOADS[name]
    from importlib import import_module  # noqa: F401
    return (f"{model.upper()} dim-{DIM} BPR batch {BATCH} on the {shape}-shaped synthetic graph ({name}"
            f"{', tripartite user-tag-item' if use_tag else ''}); full-graph propagation fwd+bwd + Adam every step")


def config_of(name):
    """Identical in both arms (the driver compares the two `config` objects)."""
    small = name in NAMED_WORKLOADS or SCALE_WORKLOADS[name]["n_edge"] < 10_000_000
    return {"workload": workload_text(name), "batch": BATCH, "dim": DIM, "lr": LR, "reg": REG,
            "l2": "L2 flushed between timed steps" if small else "inputs larger than L2 (tables >> 126 MB)"}


# ======================================================================================================================
#  CPU arm: the reference's implementation of the path on the host cores
# ======================================================================================================================
def find_reference():
    for p in (os.path.join(ROOT, "baseline", "_ref"), "/root/reference"):
        if os.path.isdir(os.path.join(p, "model")) and os.path.isdir(os.path.join(p, "training")):
            return p
    return None


def import_reference(path):
    """SURVEY Appendix C shims (all outside the reference tree); returns the reference's CFG dict."""
    import collections
    import collections.abc
    import types

    import numpy as np
    collections.Iterable = collections.abc.Iterable
    np.int = int
    tb = types.ModuleType("tensorboardX")
    tb.SummaryWriter = object
    sys.modules.setdefault("tensorboardX", tb)
    if path not in sys.path:
        sys.path.insert(0, path)
    argv, sys.argv = sys.argv, ["bench", "--model", "lightgcn", "--use_tag", "", "--cpu_core", "1",
                                "--dim_layer_list", "[64,64,64]", "--topks", "[20]"]
    cwd = os.getcwd()
    os.chdir("/tmp")                       # utility/word.py creates run/<model>/... relative to the CWD
    try:
        from utility.word import CFG
    finally:
        sys.argv = argv
        os.chdir(cwd)
    return CFG


class CpuArm:
    """One model of the reference on torch-CPU with all host threads: `.step(batch)`, `.tables()` (propagated user /
    item tables for evaluation).  kind = "reference" (unmodified classes) or "port" (oracle/train_step.py)."""

    def __init__(self, model, ds, over, threads):
        import numpy as np
        import torch
        torch.set_num_threads(threads)
        self.model_name, self.ds, self.threads = model, ds, threads
        ref = find_reference() if os.environ.get("TAGREC_BENCH_PORT") != "1" else None
        self.kind = "reference" if ref else "port"
        U, I = ds.num["user"], ds.num["item"]
        if ref:
            CFG = import_reference(ref)
            from utility.config import dict_map
            CFG.update(dict(train_batch=BATCH, test_batch=512, has_val=False, use_tag=bool(ds.num.get("tag")) and model in
                            ("tgcn",), topks=[20], lr=LR, reg=REG, cor_reg=0, dim_latent=DIM, dim_layer_list=[DIM] * LAYERS,
                            message_drop_list=[0., 0., 0.], node_drop=0., seed=2020, cpu_core=threads, split_adj_k=1,
                            device=torch.device("cpu"), model=model))
            CFG.update(dict_map[model])
            CFG.update(over)
            from utility.utils import init_seed
            init_seed(2020)
            import importlib
            cls = getattr(importlib.import_module(f"model.{model}"), {"lightgcn": "LightGCN", "ngcf": "NGCF", "dgcf": "DGCF",
                                                                      "tgcn": "TGCN"}[model])
            with np.errstate(divide="ignore"):
                self.m = cls(ds)
            self.m.train()
            self.opt = torch.optim.Adam(self.m.parameters(), lr=LR)                    # com.py:25
            self.impl = f"unmodified reference classes from {ref}"
        else:
            from oracle import adjacency as OA
            from oracle import train_step as TS
            e = ds.edge_index["train"]
            if model == "tgcn":
                np.random.seed(2020)
                tables = ds.get_all_neighbor()
                self.m = TS.TGCNStep((U, I, ds.num["tag"], ds.num["weight"]), tables, n_layer=len(over.get("dim_layer_list", [64, 64])),
                                     neighbor_k=over.get("neighbor_k", 25), reg=REG, lr=LR)
            else:
                norm = {"lightgcn": "bi_norm", "ngcf": "ngcf", "dgcf": "plain"}[model]
                csr = OA.creat_adj(U, I, (e[:, 0], e[:, 1]), norm)
                cls = {"lightgcn": TS.LightGCNStep, "ngcf": TS.NGCFStep, "dgcf": TS.DGCFStep}[model]
                self.m = cls(U, I, csr, reg=REG, lr=LR)
            self.impl = "oracle/train_step.py (op-for-op torch-CPU port of the reference step)"

    def step(self, batch):
        if self.kind == "port":
            return self.m.step(batch)
        lossx = self.m.loss((batch, None) if self.model_name == "dgcf" else batch)         # basic_train.py:15-27
        parts = [x.cpu().item() for x in lossx]
        loss = sum(lossx)
        self.opt.zero_grad()
        loss.backward()
        self.opt.step()
        return parts, loss.cpu().item()

    def tables(self):
        import torch
        with torch.no_grad():
            return [t.detach() for t in self.m.forward()[:2]]


def host_batches(ds, n, seed=0):
    """n fixed (BATCH, 3) triple batches (positives from the train edges, uniform negatives) for the CPU arm."""
    import numpy as np
    import torch
    rng = np.random.RandomState(seed)
    e = ds.edge_index["train"]
    out = []
    for _ in range(n):
        sel = rng.randint(0, len(e), BATCH)
        out.append(torch.from_numpy(np.stack([e[sel, 0], e[sel, 1], rng.randint(0, ds.num["item"], BATCH)], 1)))
    return out


def cpu_train_leg(arm, ds, steps, warmup, budget_s=150.0):
    """`warmup` + `steps` steps of the CPU arm; fewer when one step is so slow that the run would not end in minutes
    (the number actually run is what is reported)."""
    batches = host_batches(ds, warmup + steps)
    times = []
    t_start = time.perf_counter()
    for s, b in enumerate(batches):
        t0 = time.perf_counter()
        arm.step(b)
        dt = time.perf_counter() - t0
        if s >= warmup:
            times.append(dt)
        elapsed = time.perf_counter() - t_start
        if times and elapsed + dt > budget_s:
            break
    return times


def cpu_named(name, steps, warmup, threads=None, eval_users=512, with_eval=True):
    """Both the `cpu_baseline` object of our line and the `--impl reference` line of a NAMED workload: the full config."""
    import numpy as np
    import tagrec_b200 as T
    threads = threads or os.cpu_count()
    model, shape, use_tag, _, over = NAMED_WORKLOADS[name]
    ds = T.data.synth_named(shape)
    if model == "tgcn":
        ds.get_all_neighbor = lambda: T.data.get_all_neighbor(ds, width=over.get("neighbor_k", 25))
    t0 = time.perf_counter()
    arm = CpuArm(model, ds, over, threads)
    setup_s = time.perf_counter() - t0
    times = cpu_train_leg(arm, ds, steps, warmup)
    t_step = sum(times) / len(times)
    out = {"value": BATCH / t_step, "unit": UNIT, "cores": threads, "kind": arm.kind,
           "sample": f"{arm.impl}; the FULL {name} configuration ({ds.num['user']} x {ds.num['item']}, "
                     f"{len(ds.edge_index['train'])} train interactions), {len(times)} timed steps after {warmup} warm-up",
           "ms_per_step": t_step * 1e3, "steps_run": len(times), "setup_s": round(setup_s, 1), "extrapolated": False}
    if with_eval:
        from oracle import metrics as OM
        from oracle import sampler as OS
        # sampler reset() the reference's way (forked workers over equal chunks, Python loop per edge)
        np.random.seed(2020)
        t0 = time.perf_counter()
        OS.reference_reset(ds.edge_index["train"], ds.user_items["train"], ds.num["item"], min(threads, 8))
        out["sampler_reset_s"] = time.perf_counter() - t0
        ut, it = arm.tables()
        n_eval = min(eval_users, len(ds.user_items["test"]))
        ev = {}
        OM.reference_epoch_test(ut, it, ds.user_items["train"], ds.user_items["test"], [20], 512, with_auc=True, max_users=16)
        for with_auc in (False, True):
            t0 = time.perf_counter()
            OM.reference_epoch_test(ut, it, ds.user_items["train"], ds.user_items["test"], [20], 512, with_auc=with_auc,
                                    max_users=n_eval)
            dt = time.perf_counter() - t0
            ev["with_auc" if with_auc else "topk_only"] = {"users_per_s": n_eval / dt, "s": dt}
        ev["users"] = n_eval
        ev["note"] = ("oracle/metrics.reference_epoch_test: dense sigmoid(U I^T) per 512-user batch, torch.topk, "
                      "sklearn roc_auc_score per user (training/basic_test.py:30-80) on the first users of the test dict; "
                      "excludes the reference's re-propagation per user batch (lightgcn.py:85)")
        out["eval"] = ev
    return out, t_step


def sample_shape(shape, target_edges=3_000_000):
    f = min(1.0, target_edges / shape["n_edge"])
    return dict(n_user=max(1000, int(shape["n_user"] * f)), n_item=max(1000, int(shape["n_item"] * f)),
                n_edge=int(shape["n_edge"] * f))


def cpu_scale(name, steps, warmup, full_nnz, threads=None):
    """lightgcn_* workloads: the CPU arm on a bounded sample graph of the same family (same mean user / item degree),
    `value` scaled linearly in nnz to the full workload (the step is SpMM-bound: 74-88 % of the reference's CPU step is
    aten::addmm).  Every reported step was actually run, on the sample."""
    import tagrec_b200 as T
    threads = threads or os.cpu_count()
    shape = sample_shape(SCALE_WORKLOADS[name])
    ds = T.data.synth_bipartite(shape["n_user"], shape["n_item"], shape["n_edge"], seed=2020)
    t0 = time.perf_counter()
    arm = CpuArm("lightgcn", ds, {}, threads)
    setup_s = time.perf_counter() - t0
    times = cpu_train_leg(arm, ds, steps, warmup)
    t_sample = sum(times) / len(times)
    nnz = 2 * len(ds.edge_index["train"])
    scale = full_nnz / nnz
    return {"value": BATCH / (t_sample * scale), "unit": UNIT, "cores": threads, "kind": arm.kind,
            "sample": f"{arm.impl} on a {ds.num['user']} x {ds.num['item']}, nnz={nnz} graph of the same degree profile: "
                      f"{len(times)} steps of {t_sample * 1e3:.1f} ms measured; value = sample triples/s / {scale:.1f} "
                      f"(linear in nnz) for the full workload",
            "ms_per_step": t_sample * 1e3, "sample_value": BATCH / t_sample, "steps_run": len(times), "extrapolated": True,
            "scale": scale, "setup_s": round(setup_s, 1)}, t_sample


def run_reference(args):
    """--impl reference (rank 0 only; the other ranks exit 0 without work)."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    name = args.workload
    if name in NAMED_WORKLOADS:
        cb, t_step = cpu_named(name, args.steps, args.warmup)
    else:
        shape = SCALE_WORKLOADS[name]
        cb, t_step = cpu_scale(name, args.steps, args.warmup, 2 * int(shape["n_edge"] * 0.94))
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": cb["steps_run"], "warmup": args.warmup, "ms_per_step": cb["ms_per_step"], "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config_of(name),
            "extrapolated": cb["extrapolated"], "scale": cb.get("scale", 1.0), "steps_requested": args.steps,
            "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(_finite(line)))


# ======================================================================================================================
#  our arm, named shapes (C1-C4): the drop-in classes end to end
# ======================================================================================================================
def build_named(T, name, dev):
    import torch
    model_name, shape, use_tag, sampler, over = NAMED_WORKLOADS[name]
    ds = T.data.synth_named(shape)
    cfg = dict(use_tag=use_tag, reg=REG, dim_latent=DIM, dim_layer_list=[DIM] * LAYERS, train_batch=BATCH, device=dev,
               lr=LR, sampler="device", topks=[20], test_batch=512, has_val=False)
    cfg.update(over)
    T.set_config(model_name, **cfg)
    torch.manual_seed(2020)
    if model_name == "tgcn":
        ds.get_all_neighbor = lambda ds=ds: T.data.get_all_neighbor(ds, width=over.get("neighbor_k", 25))
    cls = {"lightgcn": T.LightGCN, "ngcf": T.NGCF, "dgcf": T.DGCF, "tgcn": T.TGCN}[model_name]
    model = cls(ds).to(dev)
    data = (T.DGCF_training_data if sampler == "DGCF" else T.BPR_training_data)(ds, None)
    return ds, model, data


def named_ours(T, name, dev, steps, warmup, flush):
    """Train steps (eager + CUDA-graph), e2e, sampler reset(), Basic_test.run with / without AUC on a named shape."""
    import torch
    t0 = time.time()
    ds, model, data = build_named(T, name, dev)
    opt = T.FusedAdam(model.parameters(), lr=LR)
    test = T.Basic_test(ds, None)
    setup_s = time.time() - t0
    # sampler reset() (train_data/bpr_training_data.py:29-45) — one full re-sample of every positive edge
    data.reset()
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    data.reset()
    torch.cuda.synchronize()
    reset_s = time.perf_counter() - t1
    batches = []
    while len(batches) < warmup + steps:
        for b in data.mini_batch():
            batches.append(b)
            if len(batches) >= warmup + steps:
                break
    model.train()

    def step(b):
        lossx = model.loss(b)
        opt.zero_grad()
        sum(lossx).backward()
        opt.step()
        return lossx

    for b in batches[:warmup]:
        step(b)
    launches0 = T.launch_count()
    evs = []
    for b in batches[warmup:]:
        flush.zero_()
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); step(b); e.record()
        evs.append((a, e))
    torch.cuda.synchronize()
    launches = T.launch_count() - launches0
    ms = sum(a.elapsed_time(e) for a, e in evs)
    # e2e: batches from pinned host memory, loss read back every step
    host = [(tuple(t.cpu().pin_memory() if torch.is_tensor(t) else t for t in b) if isinstance(b, (tuple, list))
             else b.cpu().pin_memory()) for b in batches[warmup:]]
    h2d = sum(t.numel() * t.element_size() for t in (host[0] if isinstance(host[0], tuple) else (host[0],)) if torch.is_tensor(t))
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    last = None
    for hb in host:
        flush.zero_()
        b = tuple(t.to(dev, non_blocking=True) if torch.is_tensor(t) else t for t in hb) if isinstance(hb, tuple) \
            else hb.to(dev, non_blocking=True)
        last = [x.cpu().item() for x in step(b)]
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t1
    # one CUDA graph per step (T.GraphedStep): the launch-bound small graphs
    graphed = None
    try:
        gopt = T.FusedAdam(model.parameters(), lr=LR, capturable=True)
        gs = T.GraphedStep(model, gopt, warmup=2)
        for b in batches[:3]:
            gs.loss(b)
        gevs = []
        for b in batches[warmup:]:
            flush.zero_()
            a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); gs.loss(b); e.record()
            gevs.append((a, e))
        torch.cuda.synchronize()
        gms = sum(a.elapsed_time(e) for a, e in gevs) / len(gevs)
        graphed = {"ms_per_step": gms, "value": BATCH / (gms / 1e3)}
    except Exception as ex:                                   # reported, not hidden
        graphed = {"error": f"{type(ex).__name__}: {ex}"[:200]}
    # evaluation through the drop-in loop (training/basic_test.py:94-111), all test users
    n_test = len(ds.user_items["test"])
    ev = {"users": n_test, "items": ds.num["item"]}
    for with_auc in (False, True):
        T.CFG["eval_auc"] = with_auc
        test.run(model, istest=True)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        res = test.run(model, istest=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t1
        ev["with_auc" if with_auc else "topk_only"] = {"users_per_s": n_test / dt, "s": dt}
    ev["result"] = {k: [float(x) for x in v] for k, v in res.items()}
    ev["note"] = "T.Basic_test(ds).run(model, istest=True): forward() + K3 top-20 + metric sums (+ K3b AUC), wall clock"
    params = torch.cat([p.detach().flatten() for p in model.parameters()]).double()
    return {"ms_per_step": ms / steps, "value": BATCH * steps / (ms / 1e3), "gpu_launches": launches,
            "e2e": {"value": BATCH * steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 8},
            "check": {"last_loss": last, "param_abs_sum": float(params.abs().sum()), "param_sq_sum": float((params * params).sum())},
            "graphed_step": graphed, "sampler_reset_s": reset_s, "eval": ev, "setup_s": round(setup_s, 1),
            "nnz": model.norm_adj._nnz() if hasattr(model, "norm_adj") and hasattr(model.norm_adj, "_nnz") else None,
            "nodes": int(sum(ds.num[k] for k in (("user", "item", "tag") if NAMED_WORKLOADS[name][2] else ("user", "item"))))}


def run_named(args):
    import torch
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        if int(os.environ.get("RANK", "0")) != 0:
            return                       # these graphs are a few MB: replicas only; rank 0 reports
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    import __graft_entry__ as G
    G.build()
    import tagrec_b200 as T
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    with ClockSampler(local) as clocks:
        r = named_ours(T, args.workload, dev, args.steps, args.warmup, flush)
    line = {"metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": r["ms_per_step"], "check": r["check"], "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config_of(args.workload),
            "run": {"optimizer": "tagrec_b200.FusedAdam", "parallelism": "single", "nnz": r["nnz"], "nodes": r["nodes"]},
            "e2e": r["e2e"], "gpu_launches": r["gpu_launches"], "clocks": clocks.summary(), "graphed_step": r["graphed_step"],
            "sampler_reset_s": r["sampler_reset_s"], "eval": r["eval"], "setup_s": r["setup_s"],
            "roofline": None}
    if not args.no_cpu_baseline:
        cb, _ = cpu_named(args.workload, 3, 1)
        line["cpu_baseline"] = cb
    print(json.dumps(_finite(line)))


# ======================================================================================================================
#  our arm, lightgcn_* (C5 family): device-built graph, single GPU or node-range row blocks
# ======================================================================================================================
def run_scale(args):
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

    shape = SCALE_WORKLOADS[args.workload]
    small = shape["n_edge"] < 10_000_000
    t0 = time.time()
    T.set_config("lightgcn", use_tag=False, reg=REG, dim_latent=DIM, dim_layer_list=[DIM] * LAYERS, train_batch=BATCH,
                 device=dev, init_device=dev, lr=LR, sampler="device")
    W, K = args.warmup, args.steps
    need = BATCH * (K + W)
    if world > 1:
        from tagrec_b200 import distributed as D
        model, triples, info = D.build_sharded_lightgcn(shape, dev, rank, world, need, eval_users_per_rank=args.eval_users)
    else:
        ui_row, ui_col = T.data.synth_bipartite_device(shape["n_user"], shape["n_item"], int(shape["n_edge"]), dev, seed=2020)
        n_train = ui_row.numel()
        graph = T.build_csr(shape["n_user"], shape["n_item"], (ui_row, ui_col), "bi_norm", dev)

        class Data:
            num = {"user": shape["n_user"], "item": shape["n_item"]}
            prebuilt_adj = graph
        torch.manual_seed(2020)
        model = T.LightGCN(Data)
        # batch stream: device sampler over a random subset of the positive edges (all steps' triples)
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
        info = {"nnz": graph._nnz(), "n": graph.n, "n_long_rows": graph.n_long, "parallelism": "single",
                "plan": graph.col_block}
    # com.py:25 composes optim.Adam(model.parameters(), lr); FusedAdam is this package's drop-in for it (same update
    # rule, one kernel per tensor; on a sharded graph each rank updates the rows it owns).  --optimizer torch runs the
    # reference's own choice.
    if args.optimizer == "fused":
        opt = T.make_optimizer(model, lr=LR) if hasattr(T, "make_optimizer") else T.FusedAdam(model.parameters(), lr=LR)
    else:
        opt = torch.optim.Adam(model.parameters(), lr=LR)
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
    e2e = {"value": BATCH * K / e2e_s, "unit": UNIT, "h2d_bytes_per_step": BATCH * 3 * 8, "d2h_bytes_per_step": 8}
    # the same W + 2K batches are trained at every N: loss and parameter checksums are comparable across rank counts
    if hasattr(opt, "consolidate"):
        opt.consolidate()
    pflat = torch.cat([p.detach() for p in model.embed]).double()
    check = {"last_loss": last, "param_abs_sum": float(pflat.abs().sum()), "param_sq_sum": float((pflat * pflat).sum()),
             "steps_trained": W + 2 * K}
    del pflat

    eval_sharded = None
    if world > 1 and args.eval_users > 0 and info.get("eval_mask") is not None:
        eval_sharded = eval_leg_sharded(T, model, shape, dev, info["eval_mask"], world, dist)
    per_rank = None
    if world > 1:
        mine = {"rank": rank, "fwd_ms": timer.mean_ms("spmm_fwd"), "fwd_last_layer_rows_ms": timer.mean_ms("spmm_fwd_rows"),
                "bwd_ms": timer.mean_ms("spmm_bwd"),
                "all_gather_ms": timer.mean_ms("all_gather"), "all_gathers_per_step": timer.count("all_gather") // K,
                "barrier_ms": timer.mean_ms("barrier"), "adam_ms": timer.mean_ms("adam"),
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
                "bpr_ms": timer.mean_ms("bpr"), "bwd_first_ms": timer.mean_ms("bwd_first"),
                # the backward launches of the last timed step, in order: G_{L-1} (source non-zero on the batch rows
                # only), G_{L-2} (batch rows + neighbours), ..., dE0 (dense source); zero source rows are skipped
                "bwd_launch_ms": [round(x.elapsed_time(y), 3) for x, y in timer.pairs.get("spmm_bwd", [])[-LAYERS:]],
                "adam_ms": timer.mean_ms("adam"),
                # the last forward layer of a training step runs on the batch's rows only (the loss reads nothing else of
                # it): its launch is timed separately and is NOT part of ms_per_launch above
                "fwd_last_layer_rows_ms": timer.mean_ms("spmm_fwd_rows")}
    if traffic:
        # `achieved` counts ALGORITHMIC bytes (every gathered 256-byte row, SURVEY §8 d); with the column-blocked plan most
        # of the item-row half's gathers are served by L2, so it exceeds the DRAM peak.  The DRAM-side figure, from the
        # ncu-measured bytes of the same launch (profiles/spmm_traffic.json) and the live launch time:
        roofline["dram_achieved"] = traffic / (fwd_ms * 1e-3) / 1e9
        roofline["dram_frac"] = roofline["dram_achieved"] / peak
        roofline["note"] = ("frac > 1: algorithmic bytes / time against the DRAM copy peak; the launch moves `traffic` DRAM bytes "
                            "(ncu), i.e. dram_frac of the peak; ncu: 69 % of stalls long-scoreboard, no unit saturated")
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "check": check, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": config_of(args.workload),
            "run": {"optimizer": type(opt).__name__, "parallelism": info["parallelism"], "nnz": info["nnz"],
                    "nodes": info["n"], "long_rows": info["n_long_rows"], "plan": info.get("plan"),
                    "step": ("model.loss(batch) / backward() / optimizer.step() of the drop-in LightGCN: layers 1..L-1 on all "
                             "rows, layer L on the rows the loss reads (the batch's users and items: nothing else of it "
                             "reaches the loss or any gradient; TAGREC_LAST_LAYER_ROWS=0 computes it everywhere), backward "
                             "over all rows, Adam over all rows — loss, gradients and parameters equal the reference's "
                             "(tests/test_gpu_parity.py, tests/test_gpu_shapes.py)")},
            "e2e": e2e, "gpu_launches": launches, "clocks": clocks.summary(), "roofline": roofline,
            "setup_s": round(setup_s, 1)}
    if per_rank:
        line["per_rank"] = per_rank
        line["partition"] = {k: info.get(k) for k in ("bounds", "bounds_bwd", "type_weight_s_per_nnz", "balance_feedback")}
    if world == 1 and not args.no_cpu_baseline:
        cb, _ = cpu_scale(args.workload, 3, 1, info["nnz"])
        line["cpu_baseline"] = cb
    if args.eval_users > 0 and world == 1:
        line["eval"] = eval_leg(T, model, shape, dev, args.eval_users)
    if eval_sharded is not None:
        line["eval"] = eval_sharded
    if world == 1 and not args.no_c1 and not small:
        # BASELINE configs[0] — "the reference's own CPU-runnable case" — both arms on the FULL config in the same run
        del model, opt, triples
        torch.cuda.empty_cache()
        fl = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
        ours = named_ours(T, "lastfm", dev, 20, 5, fl)
        c1 = {"config": config_of("lastfm"), "ours": ours}
        if not args.no_cpu_baseline:
            cpu, _ = cpu_named("lastfm", 10, 2)
            c1["cpu"] = cpu
        line["c1"] = c1
    print(json.dumps(_finite(line)))
    if world > 1:
        dist.destroy_process_group()


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


def measure_tf32_peak(dev, n=8192, reps=10):
    """Measured dense TF32 throughput of this GPU: cuBLAS fp32 GEMM with TF32 tensor cores allowed, 8192^3, best of
    `reps` (CUDA events) — the denominator for the K3-TC tensor fraction (the driver's MEASURED_PEAKS.json has bf16 only)."""
    import torch
    old = torch.backends.cuda.matmul.allow_tf32
    try:
        torch.backends.cuda.matmul.allow_tf32 = True
        a = torch.randn(n, n, device=dev)
        b = torch.randn(n, n, device=dev)
        torch.matmul(a, b)
        best = float("inf")
        for _ in range(reps):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(); torch.matmul(a, b); e.record()
            torch.cuda.synchronize()
            best = min(best, s.elapsed_time(e))
        return 2.0 * n ** 3 / (best * 1e-3) / 1e12
    except Exception:
        return None
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old


def eval_leg(T, model, shape, dev, n_users):
    """Secondary metric of BASELINE.json: full-rank eval users/s (K3: scoring + mask + top-20 + metric sums) on the
    benchmark graph.  The drop-in Basic_test takes the reference's dict-of-lists data object, which cannot hold 10 M
    users; this leg calls the same kernels Basic_test._sums_device calls, on device CSR masks ("c1" in the line times
    Basic_test.run itself).  Both scoring paths are timed: the tcgen05 TF32-filter path (default for dim 64) and the
    exact-fp32 CUDA-core path; they return identical lists (checked here on the benchmark inputs)."""
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
    from tagrec_b200.eval_ops import eval_plan
    plan = eval_plan(n_users, shape["n_item"], DIM, 20)
    tf32_peak = 1100.0      # nominal dense TF32 TFLOP/s (B200_PROFILING.md); MEASURED_PEAKS.json only has bf16
    measured_tf32 = measure_tf32_peak(dev)
    return {"users_per_s": out["tf32"]["users_per_s"], "users": n_users, "items": shape["n_item"], "k": 20,
            "ms": out["tf32"]["ms"], "tflops": out["tf32"]["tflops"],
            "tensor_frac_of_nominal_tf32": out["tf32"]["tflops"] / tf32_peak,
            "tf32_peak_measured_tflops": measured_tf32,
            "tensor_frac_of_measured_tf32": out["tf32"]["tflops"] / measured_tf32 if measured_tf32 else None,
            "kernel": ("eval_tc2_kernel (CTA pairs: tcgen05.mma.cta_group::2 M256xN256xK8 kind::tf32 filter + exact fp32 re-score)"
                       if plan["cta_pairs"] else "eval_tc_kernel (tcgen05.mma kind::tf32 filter + exact fp32 re-score)"),
            "plan": plan,
            "fp32_cuda_core_path": out["fp32"], "paths_identical": same, "auc": auc_out}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="lightgcn_1b", choices=sorted(SCALE_WORKLOADS) + sorted(NAMED_WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-c1", action="store_true", help="skip the LastFM-shaped (configs[0]) sub-benchmark of the default line")
    ap.add_argument("--eval-users", type=int, default=16384)
    ap.add_argument("--optimizer", default="fused", choices=["fused", "torch"])
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    elif args.workload in NAMED_WORKLOADS:
        run_named(args)
    else:
        run_scale(args)


if __name__ == "__main__":
    main()
