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
import numpy as np
import torch
import torch.distributed as dist

from .adj import CsrGraph


def partition_rows(rowptr, world, type_bounds=None, type_weight=None, row_cost=3.0):
    """Contiguous row ranges with (almost) equal COST.  cost(row) = w[type(row)] * nnz(row) + row_cost * min(w):
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
        cost = deg
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


class RowComm:
    """In-place all-gather of row blocks of a full-size [N, dim] table."""

    def __init__(self, bounds, rank, world, group=None):
        self.bounds, self.rank, self.world, self.group = list(bounds), rank, world, group
        self.bytes_moved = 0

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


def shard_graph(full: CsrGraph, rank, world, group=None, calibrate=True):
    """Row block of ``full`` for this rank (plus the communicator that reassembles tables).  With ``calibrate`` the
    cut points equalise MEASURED cost (see partition_rows), otherwise nnz."""
    if calibrate and world > 1 and full.device.type == "cuda":
        tb, tw = calibrate_type_weights(full, group=group)
        bounds = partition_rows(full.rowptr, world, tb, tw)
    else:
        tb, tw = None, None
        bounds = partition_rows(full.rowptr, world)
    comm = RowComm(bounds, rank, world, group)
    lo, hi = comm.lo, comm.hi
    rp, col, val = slice_csr(full.rowptr, full.col, full.val, lo, hi)
    val_t = None
    if full.val_t is not full.val:
        a, b = int(full.rowptr[lo]), int(full.rowptr[hi])
        val_t = full.val_t[a:b].clone()
    g = CsrGraph(full.n, rp, col, val, val_t, None, full.norm_type, full.num_list, row_offset=lo, comm=comm)
    g.type_weight = tw
    return g


def build_sharded_lightgcn(shape, dev, rank, world, n_triples, seed=2020):
    """bench.py helper for N > 1: every rank generates the same synthetic graph (same device RNG stream), builds the
    CSR, keeps its row block, and samples the same batch stream.  Returns (model, triples, info)."""
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
    del edges, train_items, train_ptr
    graph = shard_graph(full, rank, world)
    nnz_full, n_long_full = full._nnz(), full.n_long
    del full
    torch.cuda.empty_cache()

    class Data:
        num = {"user": shape["n_user"], "item": shape["n_item"]}
        prebuilt_adj = graph
    torch.manual_seed(seed)
    model = T.LightGCN(Data)
    info = {"nnz": graph._nnz(), "n": graph.n_rows, "n_long_rows": graph.n_long, "nnz_global": nnz_full,
            "parallelism": f"node-range row blocks x{world} (all-gather per layer, replicated parameters)",
            "rows_local": graph.n_rows, "bounds": graph.comm.bounds, "type_weight_s_per_nnz": graph.type_weight}
    return model, triples, info
