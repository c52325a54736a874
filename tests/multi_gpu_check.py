"""Run under torchrun on N >= 2 GPUs: the node-range sharded LightGCN training step equals the single-GPU step on the
same inputs, for EVERY exchange path the product has:

    nccl          K1 on row blocks + NCCL all-gather of each layer's rows
    peer-stores   all-gather fused into K1's epilogue through per-peer mappings (TAGREC_MULTICAST=0)
    multicast     the same through the NVLS multicast address (what bench.py runs on an NVSwitch box)
    + "rebalanced": the fused path after the measured re-partition bench.py applies
    + "sharded-adam": owner-sharded FusedAdam (each rank updates its row block and stores the new rows to all ranks)
    + "sharded-adam-epilogue": the same update inside the epilogue of the last backward launch
      (tagrec_lightgcn_bwd_layer_adam): losses of every step and the parameters after K steps
    + "split-partition[-epilogue]": forward and backward launches on DIFFERENT row blocks (graph.bwd_graph), with the
      replicated torch optimizer and with the owner-sharded Adam epilogue
    + "ngcf-sharded": NGCF (K1 row blocks + K6 on the local rows, layer outputs all-gathered, DENSE weight gradients
      all-reduced): loss, propagated tables and the gradient of every parameter vs the single-GPU model

Checked per mode, against the single-GPU run of the same K steps (SURVEY §4 / §8 e: within 1e-5):
loss and reg of every step, the gradient and the propagated tables of step 1, the parameters after K Adam steps, and
that the replicas (gradients, parameters) are BIT-identical across ranks.  Prints one JSON line per mode and
'MULTI_GPU_OK' / 'MULTI_GPU_FAIL'; exit code 0 only on success.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tests/multi_gpu_check.py [--out profiles/r2_multi_gpu_parity_n2.jsonl]
"""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

STEPS = 3
TOL = 1e-5
# the check graph is small: switch the big-graph paths on (last forward layer on the loss's rows, each rank its share)
os.environ.setdefault("TAGREC_LAST_LAYER_ROWS", "force")
os.environ.setdefault("TAGREC_PUSH_BWD", "force")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--users", type=int, default=30000)
    ap.add_argument("--items", type=int, default=6000)
    args = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    import tagrec_b200 as T
    from tagrec_b200.distributed import rebalance_by_measurement, shard_graph
    U, I = args.users, args.items
    rng = np.random.RandomState(0)
    ne = 13 * U
    u = np.r_[rng.randint(0, U, ne), rng.permutation(U)[:9000]]
    i = np.r_[(rng.zipf(1.3, ne) - 1) % I, np.zeros(9000, dtype=np.int64)]         # item 0 is a long row
    key = np.unique(u.astype(np.int64) * I + i)
    ui = (key // I, key % I)
    T.set_config("lightgcn", use_tag=False, reg=1e-3, dim_layer_list=[64, 64, 64], device=dev, init_device=dev)
    full = T.build_csr(U, I, ui, "bi_norm", dev)
    sel = rng.randint(0, len(key), (STEPS, 2048))
    batches = [torch.tensor(np.stack([ui[0][s], ui[1][s], rng.randint(0, I, 2048)], 1), device=dev) for s in sel]

    def run(graph, optimizer="torch"):
        class D:
            num = {"user": U, "item": I}
            prebuilt_adj = graph
        torch.manual_seed(5)
        m = T.LightGCN(D)
        m.train()
        if optimizer in ("sharded", "sharded-epilogue"):
            opt = T.ShardedFusedAdam(m, lr=0.001, fused_backward=(optimizer == "sharded-epilogue"))
        elif optimizer == "fused":
            opt = T.FusedAdam(m.parameters(), lr=0.001)
        else:
            opt = torch.optim.Adam(m.parameters(), lr=0.001)
        losses, g1 = [], None
        m.eval()                             # the propagated tables of the INITIAL parameters (before any Adam step)
        with torch.no_grad():
            f1 = torch.cat([t.detach() for t in m.forward()]).clone()
        m.train()
        for s in range(STEPS):
            lossx = m.loss(batches[s])
            opt.zero_grad()
            sum(lossx).backward()
            if s == 0 and optimizer != "sharded-epilogue":       # the epilogue form never materialises a gradient
                g1 = torch.cat([p.grad for p in m.embed]).clone()
            opt.step()
            losses.append([x.item() for x in lossx])
        if optimizer == "sharded-epilogue":
            assert all(p.grad is None for p in m.embed)
        params = torch.cat([p.detach() for p in m.embed]).clone()
        og = getattr(graph, "bwd_graph", None) or graph
        own = (og.comm.lo, og.comm.hi) if (optimizer.startswith("sharded") and graph.comm is not None) else None
        return np.array(losses), g1, f1, params, own

    rel = lambda a, b: float((a - b).abs().max() / b.abs().max())

    # Parameters after K Adam steps: Adam divides by sqrt(v) + eps, and the gradients of this model are 1e-6 ... 1e-9
    # (mean BPR loss over 2048 triples x 0.01-scale rows) — the eps = 1e-8 regime, where du/dg = eps/(|g|+eps)^2 ~ 1e7
    # turns the last-bit differences between two summation orders (K1's long rows meet through atomics: even two
    # single-GPU runs differ) into 1e-6 ... 1e-4 relative differences of the update.  So the parameter check is a loose
    # sanity bound (PARAM_TOL); the trajectory is pinned through the per-step losses and the step-1 gradient at 1e-5, and
    # the optimizer arithmetic itself in tests/test_gpu_parity.py::test_fused_adam_vs_torch on identical gradients.
    PARAM_TOL = 2e-3
    ref = {o: run(full, o) for o in ("torch", "fused")}
    fused_vs_torch = rel(ref["fused"][3], ref["torch"][3])
    rerun_vs_run = rel(run(full, "torch")[3], ref["torch"][3])   # the noise floor: the same single-GPU run twice
    assert fused_vs_torch <= PARAM_TOL, fused_vs_torch

    def sharded(mode):
        os.environ["TAGREC_MULTICAST"] = "0" if mode == "peer-stores" else "1"
        g = shard_graph(full, rank, world)
        kind = "nccl all-gather"
        if mode != "nccl":
            g.comm.enable_p2p(dev).table("probe", (8, 64))
            kind = g.comm.peer.kind
            if mode.startswith("split-partition"):
                # forward rows on the calibrated cut, backward rows on the plain nnz cut: two different row blocks
                from tagrec_b200.distributed import partition_rows, shard_with_bounds
                gb = shard_with_bounds(full, partition_rows(full.rowptr, world), rank, world, g.comm.group, g.comm.peer)
                assert gb.comm.bounds != g.comm.bounds, "the two cuts coincide: the mode would test nothing"
                g.bwd_graph = gb
            if mode in ("rebalanced", "sharded-adam", "sharded-adam-epilogue"):
                def make_model(gr):
                    class D:
                        num = {"user": U, "item": I}
                        prebuilt_adj = gr
                    torch.manual_seed(5)
                    return T.LightGCN(D)
                g = rebalance_by_measurement(full, g, rank, world, make_model=make_model, batch=batches[0], rounds=2,
                                             tol=0.0)
        return g, kind

    ok_all, lines = True, []
    for mode in ("nccl", "peer-stores", "multicast", "rebalanced", "sharded-adam", "sharded-adam-epilogue",
                 "split-partition", "split-partition-epilogue"):
        g, kind = sharded(mode)
        optimizer = {"sharded-adam": "sharded", "sharded-adam-epilogue": "sharded-epilogue",
                     "split-partition-epilogue": "sharded-epilogue"}.get(mode, "torch")
        want = ref["fused" if optimizer.startswith("sharded") else "torch"]
        losses, g1, f1, params, own = run(g, optimizer)
        errs = {"loss": float(np.abs(losses[:, 0] - want[0][:, 0]).max() / np.abs(want[0][:, 0]).max()),
                "reg": float(np.abs(losses[:, 1] - want[0][:, 1]).max() / np.abs(want[0][:, 1]).max()),
                "final": rel(f1, want[2])}
        if g1 is not None:
            if own is not None:             # sharded optimizer: a rank's gradient is defined on its own rows only
                lo, hi = own
                g1c, w1c = g1[lo:hi], want[1][lo:hi]
            else:
                g1c, w1c = g1, want[1]
            errs["grad"] = float((g1c - w1c).abs().max() / want[1].abs().max())
        errs["params_after_steps"] = rel(params, want[3])
        ok = all(v < (PARAM_TOL if k == "params_after_steps" else TOL) for k, v in errs.items())
        # replicas must be bit-identical across ranks (each row is produced by exactly one rank)
        same = True
        for t in ([params, f1] if own is not None else [params, f1, g1]):
            if t is None:
                continue
            r0 = t.clone()
            dist.broadcast(r0, src=0)
            same = same and bool(torch.equal(r0, t))
        flags = torch.tensor([int(ok), int(same)], device=dev)
        dist.all_reduce(flags, op=dist.ReduceOp.MIN)
        ok_all = ok_all and bool(flags.min().item() == 1)
        line = {"mode": mode, "exchange": kind, "world": world, "steps": STEPS, "errs_vs_single_gpu": errs,
                "tolerance": TOL, "param_tolerance": PARAM_TOL,
                "single_gpu_noise": {"fused_vs_torch_adam_params": fused_vs_torch, "same_run_twice_params": rerun_vs_run},
                "within_tolerance_all_ranks": bool(flags[0].item()),
                "replicas_bit_identical": bool(flags[1].item()), "bounds": g.comm.bounds,
                "losses": losses[:, 0].tolist(), "graph": {"users": U, "items": I, "nnz": full._nnz(),
                                                           "long_rows": full.n_long}}
        lines.append(line)
        if rank == 0:
            print(json.dumps(line), flush=True)
        del g
        torch.cuda.synchronize()
        dist.barrier()
    # ---- NGCF on the sharded graph: K1 row blocks + K6 on local rows, all-gather of the layer outputs, all-reduce of
    # the dense weight gradients (distributed.ShardedSpMMFn & co.) vs the single-GPU model ----
    T.set_config("ngcf", use_tag=False, reg=1e-3, dim_layer_list=[64, 64, 64], device=dev, init_device=dev)
    full_n = T.build_csr(U, I, ui, "ngcf", dev)

    def run_ngcf(graph):
        class D:
            num = {"user": U, "item": I}
            prebuilt_adj = graph
        torch.manual_seed(5)
        m = T.NGCF(D).to(dev)
        m.train()
        lossx = m.loss(batches[0])
        sum(lossx).backward()
        m.eval()
        with torch.no_grad():
            fin = torch.cat([t.detach() for t in m.forward()]).clone()
        return [x.item() for x in lossx], {k: p.grad.clone() for k, p in m.named_parameters()}, fin

    l1, g1n, f1n = run_ngcf(full_n)
    _, g1b, _ = run_ngcf(full_n)                           # the same single-GPU step again: its own run-to-run noise
    gs = shard_graph(full_n, rank, world)
    l2, g2n, f2n = run_ngcf(gs)
    gmax = max(float(v.abs().max()) for v in g1n.values())
    terr = lambda a, b, k: float((a[k] - b[k]).abs().max()) / max(float(b[k].abs().max()), 1e-7 * gmax)   # noqa: E731
    emb = [k for k in g1n if k.startswith("embed.")]
    dense = [k for k in g1n if not k.startswith("embed.")]
    # Embedding-gradient rows are produced by the same kernels in the same order on one rank: 1e-5.  The DENSE weight /
    # bias gradients are float32 sums over all N rows whose order differs by construction (partial sums per rank, then an
    # all-reduce; on one GPU per block, then atomics — two single-GPU runs already differ): their bar is 1e-4 of the
    # tensor's largest entry, reported next to the single-GPU run-to-run figure.  (In float64 the sharded and the
    # unsharded gradients agree to 1e-10: tests/test_distributed_cpu.py::test_sharded_ngcf_world2_gloo.)
    DENSE_TOL = 1e-4
    errs = {"loss": abs(l2[0] - l1[0]) / abs(l1[0]), "reg": abs(l2[1] - l1[1]) / abs(l1[1]), "final": rel(f2n, f1n),
            "grad_embed": max(terr(g2n, g1n, k) for k in emb), "grad_dense": max(terr(g2n, g1n, k) for k in dense)}
    noise = {"grad_embed": max(terr(g1b, g1n, k) for k in emb), "grad_dense": max(terr(g1b, g1n, k) for k in dense)}
    ok = all(v < (DENSE_TOL if k == "grad_dense" else TOL) for k, v in errs.items())
    same = True
    for t in list(g2n.values()) + [f2n]:
        r0 = t.clone()
        dist.broadcast(r0, src=0)
        same = same and bool(torch.equal(r0, t))
    flags = torch.tensor([int(ok), int(same)], device=dev)
    dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    ok_all = ok_all and bool(flags.min().item() == 1)
    line = {"mode": "ngcf-sharded", "exchange": "nccl all-gather of layer rows + all-reduce of dense weight gradients",
            "world": world, "errs_vs_single_gpu": errs, "tolerance": TOL, "dense_grad_tolerance": DENSE_TOL,
            "single_gpu_run_to_run": noise, "within_tolerance_all_ranks": bool(flags[0].item()), "replicas_bit_identical": bool(flags[1].item()),
            "bytes_moved_per_step": gs.comm.bytes_moved, "params": sorted(g1n)}
    lines.append(line)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if rank == 0:
        print("MULTI_GPU_OK" if ok_all else "MULTI_GPU_FAIL", flush=True)
        if args.out:
            os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
            with open(args.out, "w") as f:
                for ln in lines:
                    f.write(json.dumps(ln) + "\n")
                f.write(json.dumps({"result": "MULTI_GPU_OK" if ok_all else "MULTI_GPU_FAIL", "world": world}) + "\n")
    dist.destroy_process_group()
    sys.exit(0 if ok_all else 1)


if __name__ == "__main__":
    main()
