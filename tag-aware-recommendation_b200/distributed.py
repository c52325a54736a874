"""Multi-GPU (one process per GPU, torch.distributed over NCCL / NVLink) node-range sharding of the hot path.

The reference is single-process / single-device; its only seam is ``split_adj_k`` row folding
(model/help/adj.py:114-140).  Here the N x N adjacency is cut into P contiguous ROW BLOCKS balanced by nnz; rank p
stores rows R_p of the CSR (global column ids) and computes rows R_p of every propagated table with the same K1
kernel (``row_offset`` = first global row).  The exchange step is one all-gather of the freshly computed row block
per layer, forward and backward (A is symmetric for bi_norm, so the backward uses the same blocks):

    forward : E^{k+1}[R_p] = A[R_p,:] E^k        -> all-gather E^{k+1}   (k+1 < L; the last raw layer stays local)
              final[R_p] (fused mean epilogue)   -> all-gather final      (the BPR kernel reads rows of any rank)
    BPR     : every rank runs the (tiny) fused BPR kernel on the WHOLE batch — no gradient all-reduce is needed
    backward: G_k[R_p] = nb(.)[R_p] + A[R_p,:] G_{k+1} -> all-gather G_k; dL/dE0[R_p] -> all-gather
Parameters and optimizer state stay replicated (state_dict / external optim.Adam unchanged, SURVEY §8 e): every row
of the gradient is computed by exactly one rank and broadcast, so the replicas stay bit-identical.
"""
import os

import torch
import torch.distributed as dist

from .adj import CsrGraph


def partition_rows(rowptr, world, type_bounds=None, type_weight=None, row_cost=3.0, range_scale=None):
    """Contiguous row ranges with (almost) equal COST.  cost(row) = w[type(row)] * nnz(row) + row_cost * min(w):
    ``range_scale`` = (bounds, factors): multiply the modelled cost of the rows of each old range by a measured
    correction (one feedback step on the real fused kernel, see build_sharded_lightgcn).
    ``type_bounds`` = cumulative node counts [0, n_user, n_user+n_item, ...] and ``type_weight`` = measured seconds
    per nnz of each node type (rows of popular-column blocks hit L2 and are cheaper than rows whose neighbours are
    spread over a table far larger than L2); both None -> plain nnz balance.  The per-row term stands for the
    epilogue traffic (776 B/row vs 264 B/nnz).  Returns a python list of world + 1 row indices."""
    rp = torch.as_tensor(rowptr)
    n = rp.numel() - 1
    deg = (rp[1:] - rp[:-1]).to(torch.float64)
    if type_weight is not None:
        w = torch.empty(n, dtype=torch.float64, device=rp.device)
        for t, wt in enumerate(type_weight):
            w[type_bounds[t]:type_bounds[t + 1]] = float(wt)
        cost = deg * w + row_cost * float(min(type_weight))
    else:
        cost = deg + row_cost
    if range_scale is not None:                          # feedback: (old_bounds, measured/predicted factor per range)
        old, fac = range_scale
        for p in range(len(fac)):
            cost[old[p]:old[p + 1]] *= float(fac[p])
    cum = torch.cat([torch.zeros(1, dtype=torch.float64, device=rp.device), torch.cumsum(cost, 0)])
    total = float(cum[-1])
    targets = torch.tensor([total * p / world for p in range(1, world)], dtype=torch.float64, device=rp.device)
    cuts = torch.searchsorted(cum, targets, right=False).clamp_(0, n).tolist() if world > 1 else []
    bounds = [0] + [int(c) for c in cuts] + [n]
    for i in range(1, len(bounds)):                      # keep the bounds monotone even for degenerate inputs
        bounds[i] = max(bounds[i], bounds[i - 1])
    return bounds


def calibrate_type_weights(full: CsrGraph, dim=64, group=None):
    """Seconds per nnz of K1 on the rows of each node type (one timed launch per type on this rank, averaged over
    ranks so that every rank derives the same partition)."""
    from .adj import spmm_raw
    tb = [0]
    for k in full.num_list:
        tb.append(tb[-1] + k)
    x = torch.randn(full.n, dim, device=full.device)
    y = torch.empty_like(x)
    w = []
    for t in range(len(full.num_list)):
        lo, hi = tb[t], tb[t + 1]
        rp, col, val = slice_csr(full.rowptr, full.col, full.val, lo, hi)
        blk = CsrGraph(full.n, rp, col, val, None, None, full.norm_type, full.num_list, row_offset=lo)
        spmm_raw(blk, x, out=y)                          # warm-up
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        spmm_raw(blk, x, out=y)
        b.record()
        torch.cuda.synchronize()
        w.append(a.elapsed_time(b) * 1e-3 / max(1, int(rp[-1])))
        del blk, rp, col, val
    wt = torch.tensor(w, dtype=torch.float64, device=full.device)
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(wt, op=dist.ReduceOp.SUM, group=group)
        wt /= dist.get_world_size(group)
    return tb, wt.tolist()


def slice_csr(rowptr, col, val, lo, hi):
    """Row block [lo, hi) of a CSR with a re-based rowptr (global column ids are kept)."""
    a, b = int(rowptr[lo]), int(rowptr[hi])
    return (rowptr[lo:hi + 1] - rowptr[lo]).clone(), col[a:b].clone(), (val[a:b].clone() if val is not None else None)


class PeerTables:
    """Full-size tables in symmetric memory (torch.distributed._symmetric_memory: cuMem allocations mapped into every
    rank of the NVSwitch domain).  A table allocated here can be the target of K1's fused epilogue stores
    (tagrec_mirror_t): every rank writes its row block straight into every rank's copy — through the NVLS multicast
    address when the fabric offers one, else through the per-peer mappings — and ``barrier()`` (a device-side
    signal exchange on the current stream) separates producers from the next kernel that gathers from the table."""

    def __init__(self, group, device):
        import torch.distributed._symmetric_memory as symm_mem
        self.symm_mem, self.group, self.device = symm_mem, group or dist.group.WORLD, device
        self.tables = {}
        self.use_multicast = os.environ.get("TAGREC_MULTICAST", "1") != "0"
        self.kind = None

    def table(self, name, shape):
        """(tensor, MirrorDesc) — allocated and rendezvoused once (collective: all ranks must ask in the same order)."""
        from ._lib import MirrorDesc
        ent = self.tables.get(name)
        if ent is None or tuple(ent[0].shape) != tuple(shape):
            t = self.symm_mem.empty(*shape, dtype=torch.float32, device=self.device)
            hdl = self.symm_mem.rendezvous(t, self.group)
            m = MirrorDesc()
            m.self = hdl.rank
            mc = int(hdl.multicast_ptr or 0) if self.use_multicast else 0
            if mc:
                m.n = 1
                m.base[0] = mc
                self.kind = "nvls-multicast"
            else:
                m.n = hdl.world_size
                for r in range(hdl.world_size):
                    m.base[r] = int(hdl.buffer_ptrs[r])
                self.kind = "peer-stores"
            ent = (t, hdl, m)
            self.tables[name] = ent
        return ent[0], ent[2]

    def barrier(self, name):
        self.tables[name][1].barrier(channel=0)


class RowComm:
    """Reassembles row blocks of a full-size [N, dim] table: NCCL all-gather (in place), or — ``peer`` set — the
    fused peer-store path where the all-gather has already happened inside the producing kernel."""

    def __init__(self, bounds, rank, world, group=None):
        self.bounds, self.rank, self.world, self.group = list(bounds), rank, world, group
        self.bytes_moved = 0
        self.peer = None

    def enable_p2p(self, device):
        """Switch to the fused path; raises if symmetric memory cannot be set up (caller decides about NCCL)."""
        self.peer = PeerTables(self.group, device)
        return self.peer

    @property
    def lo(self):
        return self.bounds[self.rank]

    @property
    def hi(self):
        return self.bounds[self.rank + 1]

    def all_gather_rows(self, table):
        """Every rank has written rows [lo, hi) of ``table``; afterwards every rank holds the whole table."""
        if self.world == 1:
            return table
        views = [table[self.bounds[p]:self.bounds[p + 1]] for p in range(self.world)]
        from . import functional as Fn
        t = Fn.KERNEL_TIMER if table.is_cuda else None
        if t:
            t.start("all_gather")
        if dist.get_backend(self.group) == "nccl":
            dist.all_gather(views, views[self.rank], group=self.group)
        else:                                            # gloo: uneven all_gather is not available
            for p in range(self.world):
                if views[p].numel():
                    dist.broadcast(views[p], src=dist.get_global_rank(self.group, p) if self.group else p,
                                   group=self.group)
        if t:
            t.stop("all_gather")
        self.bytes_moved += table.numel() * table.element_size()
        return table


# ----------------------------------------------------------------------------------------------------------------------
#  Autograd building blocks for models composed from split_mm + row-wise layers (NGCF) on a row-sharded graph.
#
#  Every [N, d] activation is REPLICATED in value; a rank computes the rows it owns and the owners' rows are
#  all-gathered.  Gradient convention: on rank p the gradient of a replicated activation is correct on p's own rows
#  (other rows are not used) — row-wise layers need nothing else, the SpMM backward all-gathers the upstream gradient
#  first (A^T needs every row of it), row-table parameters all-gather their gradient rows at the end, and DENSE
#  parameters (NGCF's W1 / W2 / biases: partial sums over the local rows) are all-reduced.
# ----------------------------------------------------------------------------------------------------------------------
class ShardedSpMMFn(torch.autograd.Function):
    """rows R_p of A @ x from the replicated x.  Returns a full-size table of which only this rank's rows are filled
    (the row-wise layer that follows reads only those).  Backward: all-gather the upstream gradient rows, then
    rows R_p of A^T @ g with K1 on the transposed values of the row block."""

    @staticmethod
    def forward(ctx, graph, x):
        from .adj import spmm_raw
        ctx.graph = graph
        out = torch.zeros_like(x)
        spmm_raw(graph, x.detach().contiguous(), out=out)
        return out

    @staticmethod
    def backward(ctx, g):
        from .adj import spmm_raw
        graph = ctx.graph
        g = g.contiguous().clone()
        graph.comm.all_gather_rows(g)
        gx = torch.zeros_like(g)
        spmm_raw(graph, g, out=gx, transposed=True)
        return None, gx


class GatherRowsFn(torch.autograd.Function):
    """[n_local, d] rows computed by this rank -> the replicated [N, d] table (all-gather).  Backward: this rank's rows
    of the incoming gradient (correct by the convention above), no communication."""

    @staticmethod
    def forward(ctx, comm, n, local):
        ctx.lo, ctx.hi = comm.lo, comm.hi
        full = torch.empty((n, local.shape[1]), dtype=local.dtype, device=local.device)
        full[comm.lo:comm.hi] = local
        comm.all_gather_rows(full)
        return full

    @staticmethod
    def backward(ctx, g):
        return None, None, g[ctx.lo:ctx.hi].contiguous()


class RowOwnedParamFn(torch.autograd.Function):
    """Identity on a replicated row table built from parameters; backward all-gathers the gradient rows from their
    owners so that every replica of the parameters receives the full gradient (replicated optimizer)."""

    @staticmethod
    def forward(ctx, comm, x):
        ctx.comm = comm
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        g = g.contiguous().clone()
        ctx.comm.all_gather_rows(g)
        return None, g


class AllReduceGradFn(torch.autograd.Function):
    """Identity on a dense (replicated) parameter; backward sums the partial gradients of the ranks — the all-reduce of
    dense gradients of the node-range design."""

    @staticmethod
    def forward(ctx, comm, w):
        ctx.comm = comm
        return w.view_as(w)

    @staticmethod
    def backward(ctx, g):
        g = g.contiguous().clone()
        dist.all_reduce(g, op=dist.ReduceOp.SUM, group=ctx.comm.group)
        ctx.comm.bytes_moved += g.numel() * g.element_size()
        return None, g


def is_sharded(graph):
    comm = getattr(graph, "comm", None)
    return comm is not None and comm.world > 1


def shard_graph(full: CsrGraph, rank, world, group=None, calibrate=True, weights=None, range_scale=None, peer=None):
    """Row block of ``full`` for this rank (plus the communicator that reassembles tables).  With ``calibrate`` the
    cut points equalise MEASURED cost (see partition_rows), otherwise nnz."""
    if weights is not None:
        tb, tw = weights
        bounds = partition_rows(full.rowptr, world, tb, tw, range_scale=range_scale)
    elif calibrate and world > 1 and full.device.type == "cuda":
        tb, tw = calibrate_type_weights(full, group=group)
        bounds = partition_rows(full.rowptr, world, tb, tw)
    else:
        tb, tw = None, None
        bounds = partition_rows(full.rowptr, world)
    comm = RowComm(bounds, rank, world, group)
    comm.peer = peer
    lo, hi = comm.lo, comm.hi
    rp, col, val = slice_csr(full.rowptr, full.col, full.val, lo, hi)
    val_t = None
    if full.val_t is not full.val:
        a, b = int(full.rowptr[lo]), int(full.rowptr[hi])
        val_t = full.val_t[a:b].clone()
    g = CsrGraph(full.n, rp, col, val, val_t, None, full.norm_type, full.num_list, row_offset=lo, comm=comm)
    g.type_weight = tw
    g.type_bounds = tb
    g.nnz_global = full._nnz()
    return g


def row_costs(full, type_bounds=None, type_weight=None, row_cost=3.0):
    """Modelled cost of every row (see partition_rows) as a float64 device vector — the state the measured
    re-partitioning below refines."""
    rp = full.rowptr
    deg = (rp[1:] - rp[:-1]).to(torch.float64)
    if type_weight is None:
        return deg + row_cost
    w = torch.empty(deg.numel(), dtype=torch.float64, device=rp.device)
    for t, wt in enumerate(type_weight):
        w[type_bounds[t]:type_bounds[t + 1]] = float(wt)
    return deg * w + row_cost * float(min(type_weight))


def cut_by_cost(cost, world):
    """world + 1 row indices splitting ``cost`` into contiguous ranges of (almost) equal total."""
    n = cost.numel()
    cum = torch.cat([torch.zeros(1, dtype=torch.float64, device=cost.device), torch.cumsum(cost, 0)])
    total = float(cum[-1])
    targets = torch.tensor([total * p / world for p in range(1, world)], dtype=torch.float64, device=cost.device)
    cuts = torch.searchsorted(cum, targets, right=False).clamp_(0, n).tolist() if world > 1 else []
    bounds = [0] + [int(c) for c in cuts] + [n]
    for i in range(1, len(bounds)):
        bounds[i] = max(bounds[i], bounds[i - 1])
    return bounds


def shard_with_bounds(full, bounds, rank, world, group=None, peer=None):
    comm = RowComm(bounds, rank, world, group)
    comm.peer = peer
    lo, hi = comm.lo, comm.hi
    rp, col, val = slice_csr(full.rowptr, full.col, full.val, lo, hi)
    val_t = None
    if full.val_t is not full.val:
        a, b = int(full.rowptr[lo]), int(full.rowptr[hi])
        val_t = full.val_t[a:b].clone()
    g = CsrGraph(full.n, rp, col, val, val_t, None, full.norm_type, full.num_list, row_offset=lo, comm=comm)
    g.nnz_global = full._nnz()
    return g


def measured_step_ms(model, batch, reps=2):
    """Per-rank K1 time of one real training step (forward layers + backward gather launches, CUDA events around every
    launch, mirrored stores and sparse first table included; no optimizer) — what the partition must equalise."""
    from . import functional as Fn
    model.train()
    out = []
    for it in range(reps + 1):
        timer = Fn.KernelTimer()
        Fn.KERNEL_TIMER = timer if it > 0 else None
        lossx = model.loss(batch)
        sum(lossx).backward()
        for p in model.parameters():
            p.grad = None
        torch.cuda.synchronize()
        Fn.KERNEL_TIMER = None
        if it > 0:
            out.append(sum(a.elapsed_time(b) for name in ("spmm_fwd", "spmm_fwd_rows", "spmm_bwd")
                           for a, b in timer.pairs.get(name, [])))
    return min(out)


def measured_fwd_bwd_ms(model, batch, reps=2):
    """(forward, backward) K1 time of one real training step on this rank, separately (see measured_step_ms)."""
    from . import functional as Fn
    model.train()
    best = None
    for it in range(reps + 1):
        timer = Fn.KernelTimer()
        Fn.KERNEL_TIMER = timer if it > 0 else None
        lossx = model.loss(batch)
        sum(lossx).backward()
        for p in model.parameters():
            p.grad = None
        torch.cuda.synchronize()
        Fn.KERNEL_TIMER = None
        if it > 0:
            # (the last layer's row-list launch lands on the ranks that own the batch's hub items: part of "forward")
            f = sum(a.elapsed_time(b) for name in ("spmm_fwd", "spmm_fwd_rows") for a, b in timer.pairs.get(name, []))
            b_ = sum(a.elapsed_time(b) for a, b in timer.pairs.get("spmm_bwd", []))
            best = (f, b_) if best is None else (min(best[0], f), min(best[1], b_))
    return best


def split_partition_by_measurement(full, graph, rank, world, make_model, batch, rounds=2, tol=0.02):
    """Separate row partitions for the forward and the backward launches (collective).  Every exchanged table is
    full-size on every rank, so a launch may cut the rows any way it likes — and a row block does not cost the same in
    both directions: on the 1 B-edge graph a user-row block is the cheaper one forward and the dearer one backward
    (8 GPUs, one partition balanced on the sum: forward 6.4-7.3 ms, backward 6.6-5.9 ms per launch; every launch ends in
    a barrier, so a step pays the MAXIMUM of each).  Starting from ``graph`` (both directions on its bounds), each round
    times a real step per rank, scales the modelled cost of every rank's forward rows by its forward time / mean and
    of its backward rows by its backward time / mean, and cuts both again.  Returns the forward graph with
    ``.bwd_graph`` set (functional.lightgcn_backward_layers and optim.ShardedFusedAdam pick it up)."""
    comm = graph.comm
    tb, tw = getattr(graph, "type_bounds", None), getattr(graph, "type_weight", None)
    cost_f = row_costs(full, tb, tw)
    cost_b = cost_f.clone()
    gf, gbk = graph, graph
    history = []
    for rnd in range(rounds):
        gf.bwd_graph = gbk if gbk is not gf else None
        model = make_model(gf)
        f_ms, b_ms = measured_fwd_bwd_ms(model, batch)
        del model
        mine = torch.tensor([f_ms, b_ms], dtype=torch.float64, device=full.device)
        allt = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allt, mine, group=comm.group)
        tf, tbw = [float(x[0]) for x in allt], [float(x[1]) for x in allt]
        mf, mb = sum(tf) / world, sum(tbw) / world
        history.append({"bounds_fwd": list(gf.comm.bounds), "bounds_bwd": list(gbk.comm.bounds),
                        "fwd_ms": [round(x, 3) for x in tf], "bwd_ms": [round(x, 3) for x in tbw]})
        if max(tf) / mf - 1.0 < tol and max(tbw) / mb - 1.0 < tol:
            break
        for p in range(world):
            cost_f[gf.comm.bounds[p]:gf.comm.bounds[p + 1]] *= tf[p] / mf
            cost_b[gbk.comm.bounds[p]:gbk.comm.bounds[p + 1]] *= tbw[p] / mb
        bf, bb = cut_by_cost(cost_f, world), cut_by_cost(cost_b, world)
        peer = comm.peer
        gf.bwd_graph = None
        del gf, gbk
        torch.cuda.empty_cache()
        gf = shard_with_bounds(full, bf, rank, world, comm.group, peer)
        gbk = shard_with_bounds(full, bb, rank, world, comm.group, peer)
        gf.type_bounds, gf.type_weight = tb, tw
        comm = gf.comm
    gf.bwd_graph = gbk if gbk is not gf else None
    gf.balance_feedback = (getattr(graph, "balance_feedback", None) or []) + history
    return gf


def rebalance_by_measurement(full, graph, rank, world, make_model=None, batch=None, rounds=2, tol=0.03):
    """Measured feedback on the partition (collective).  Each round times one REAL training step per rank (forward and
    backward K1 launches: the backward of a user-row block costs more than its forward, so balancing the forward alone
    leaves ranks 20 % apart), scales the modelled cost of every rank's rows by measured / mean and cuts again.
    ``make_model(graph)`` builds the model on a candidate partition; without it one forward layer is timed (the
    round-1 behaviour, kept for callers that have no model yet)."""
    comm = graph.comm
    cost = row_costs(full, getattr(graph, "type_bounds", None), getattr(graph, "type_weight", None))
    history = []
    for rnd in range(rounds):
        if make_model is not None:
            model = make_model(graph)
            mine_ms = measured_step_ms(model, batch)
            del model
        else:
            mine_ms = _forward_layer_ms(graph)
        mine = torch.tensor([mine_ms], dtype=torch.float64, device=full.device)
        allt = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allt, mine, group=comm.group)
        tms = [float(x.item()) for x in allt]
        mean = sum(tms) / world
        history.append({"bounds": list(comm.bounds), "k1_ms_per_step": [round(x, 3) for x in tms]})
        if max(tms) / mean - 1.0 < tol:
            break
        for p in range(world):
            cost[comm.bounds[p]:comm.bounds[p + 1]] *= tms[p] / mean
        bounds = cut_by_cost(cost, world)
        peer, tb, tw = comm.peer, getattr(graph, "type_bounds", None), getattr(graph, "type_weight", None)
        del graph
        torch.cuda.empty_cache()
        graph = shard_with_bounds(full, bounds, rank, world, comm.group, peer)
        graph.type_bounds, graph.type_weight = tb, tw
        comm = graph.comm
    graph.balance_feedback = history
    return graph


def _forward_layer_ms(graph, dim=64):
    """One fused forward layer on this rank's block (mirrored stores included), best of 2 after a warm-up."""
    from ._lib import check, lib, ptr, stream_ptr
    import ctypes as C
    dev, n, comm = graph.device, graph.n, graph.comm
    m = None
    if comm.peer is not None:
        raw0, m = comm.peer.table("raw0", (n, dim))
        final, _ = comm.peer.table("final", (n, dim))
    else:
        raw0, final = torch.empty((n, dim), device=dev), torch.empty((n, dim), device=dev)
    e0 = torch.randn(n, dim, device=dev) * 0.1
    d = graph.desc(dim)
    times = []
    for it in range(3):
        torch.cuda.synchronize()
        dist.barrier(group=comm.group)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        check(lib().tagrec_lightgcn_fwd_layer_p2p(C.byref(d), ptr(e0), ptr(raw0), ptr(final), dim, 1, 0, 0.25,
                                                  C.byref(m) if m is not None else None, None, stream_ptr(dev)),
              "calibration layer")
        b.record()
        torch.cuda.synchronize()
        times.append(a.elapsed_time(b))
    return min(times[1:])


def build_sharded_lightgcn(shape, dev, rank, world, n_triples, seed=2020, eval_users_per_rank=0):
    """bench.py helper for N > 1: every rank generates the same synthetic graph (same device RNG stream), builds the
    CSR, keeps its row block, and samples the same batch stream.  Returns (model, triples, info).
    ``eval_users_per_rank`` > 0 also keeps the train rows (evaluation masks) of this rank's share of the evaluation
    users — users [rank*E, (rank+1)*E) — as a local CSR in ``info["eval_mask"] = (first_user, ptr, items)``."""
    import tagrec_b200 as T
    ui_row, ui_col = T.data.synth_bipartite_device(shape["n_user"], shape["n_item"], int(shape["n_edge"]), dev, seed=seed)
    n_train = ui_row.numel()
    full = T.build_csr(shape["n_user"], shape["n_item"], (ui_row, ui_col), "bi_norm", dev)
    g = torch.Generator(device=dev)
    g.manual_seed(1)
    idx = torch.randint(0, n_train, (n_triples,), device=dev, generator=g)
    edges = torch.stack([ui_row[idx], ui_col[idx]], 1).contiguous()
    del ui_row, ui_col, idx
    U = shape["n_user"]
    train_ptr = full.rowptr[:U + 1].contiguous()
    train_items = (full.col[:n_train] - U).contiguous()
    triples = torch.empty((n_triples, 3), dtype=torch.int64, device=dev)
    T._lib.check(T._lib.lib().tagrec_sample_bpr_device(T._lib.ptr(edges), n_triples, T._lib.ptr(train_ptr),
                                                       T._lib.ptr(train_items), shape["n_item"], seed, 0,
                                                       T._lib.ptr(triples), T._lib.stream_ptr(dev)), "sampler")
    eval_mask = None
    if eval_users_per_rank > 0:
        e_n = min(int(eval_users_per_rank), U // world)
        lo = rank * e_n
        a, b = int(train_ptr[lo]), int(train_ptr[lo + e_n])
        eval_mask = (lo, (train_ptr[lo:lo + e_n + 1] - a).contiguous(), train_items[a:b].contiguous())
    del edges, train_items, train_ptr
    graph = shard_graph(full, rank, world)
    mode = "nccl all-gather per layer"
    if os.environ.get("TAGREC_P2P", "1") != "0":
        try:
            graph.comm.enable_p2p(dev).table("probe", (8, 64))
            mode = f"all-gather fused into the K1 epilogue ({graph.comm.peer.kind} over NVLink)"
        except Exception as e:                       # loud, not silent: the JSON line names the path that ran
            print(f"[tagrec_b200] symmetric memory unavailable ({type(e).__name__}: {e}); using NCCL all-gather",
                  flush=True)
            graph.comm.peer = None
    def make_model(g):
        class Data:
            num = {"user": shape["n_user"], "item": shape["n_item"]}
            prebuilt_adj = g
        torch.manual_seed(seed)
        return T.LightGCN(Data)

    if os.environ.get("TAGREC_REBALANCE", "1") != "0":
        graph = rebalance_by_measurement(full, graph, rank, world, make_model=make_model, batch=triples[:2048],
                                         rounds=int(os.environ.get("TAGREC_REBALANCE_ROUNDS", "3")))
    if os.environ.get("TAGREC_SPLIT_PARTITION", "1") != "0" and graph.comm.peer is not None:
        graph = split_partition_by_measurement(full, graph, rank, world, make_model, triples[:2048],
                                               rounds=int(os.environ.get("TAGREC_SPLIT_ROUNDS", "3")))
    nnz_full, n_long_full = full._nnz(), full.n_long
    del full
    torch.cuda.empty_cache()

    class Data:
        num = {"user": shape["n_user"], "item": shape["n_item"]}
        prebuilt_adj = graph
    torch.manual_seed(seed)
    model = T.LightGCN(Data)
    info = {"nnz": graph._nnz(), "n": graph.n_rows, "n_long_rows": graph.n_long, "nnz_global": nnz_full,
            "parallelism": f"node-range row blocks x{world}, {mode}, replicated parameters",
            "rows_local": graph.n_rows, "bounds": graph.comm.bounds,
            "bounds_bwd": graph.bwd_graph.comm.bounds if getattr(graph, "bwd_graph", None) is not None else None,
            "type_weight_s_per_nnz": graph.type_weight,
            "balance_feedback": getattr(graph, "balance_feedback", None), "eval_mask": eval_mask,
            "plan": graph.col_block}
    return model, triples, info
