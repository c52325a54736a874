"""Glue over the K5 routing kernels (csrc/routing.cu): DGCF's intent-aware propagation (model/dgcf.py:49-110) and
DisenGCN's neighbour routing (model/disengcn.py:29-43) as autograd nodes.

No gradient flows through the routing weights in the reference (`.detach()`, dgcf.py:92, disengcn.py:36), so the
backward of a layer is the transposed per-factor operator of its LAST routing iteration.  Those weights are not
symmetric, hence the reverse-edge permutation.
"""
import ctypes as C

import torch

from ._lib import RoutePlan, check, lib, ptr, stream_ptr

FACTORS = 4


def _st(t):
    return stream_ptr(t.device)


def reverse_perm(graph):
    """rev[e(h,t)] = e(t,h), cached on the graph."""
    rev = getattr(graph, "_rev_perm", None)
    if rev is None:
        rev = torch.empty(graph._nnz(), dtype=torch.int32, device=graph.device)
        missing = torch.zeros(1, dtype=torch.int32, device=graph.device)
        check(lib().tagrec_csr_reverse_perm(ptr(graph.rowptr), ptr(graph.col), graph.n_rows, ptr(rev), ptr(missing),
                                            stream_ptr(graph.device)), "tagrec_csr_reverse_perm")
        if int(missing.item()) != 0:
            raise RuntimeError("adjacency structure is not symmetric: no reverse edge for some entries")
        graph._rev_perm = rev
    return rev


def edge_rows(graph):
    """Row id of every CSR entry (int32 [nnz]) + the long-row plan of R3 (tagrec_route_plan_t); cached on the graph."""
    cached = getattr(graph, "_edge_rows", None)
    if cached is None:
        L = lib()
        dev = graph.device
        er = graph.row_ids().to(torch.int32).contiguous()
        deg = graph.rowptr[1:] - graph.rowptr[:-1]
        long_rows = torch.nonzero(deg > int(L.tagrec_spmm4_long_threshold())).flatten()
        plan, keep = None, None
        if long_rows.numel():
            piece = int(L.tagrec_spmm4_piece())
            npieces = (deg[long_rows] + piece - 1) // piece
            slot = torch.repeat_interleave(torch.arange(long_rows.numel(), device=dev), npieces)
            first = torch.cumsum(npieces, 0) - npieces
            k = torch.arange(slot.numel(), device=dev) - first[slot]
            begin = (graph.rowptr[long_rows][slot] + k * piece).contiguous()
            end = torch.minimum(begin + piece, graph.rowptr[long_rows + 1][slot]).contiguous()
            keep = (long_rows.to(torch.int32).contiguous(), slot.to(torch.int32).contiguous(), begin, end,
                    torch.zeros((long_rows.numel(), 16 * FACTORS), dtype=torch.float32, device=dev))
            plan = RoutePlan()
            plan.long_rows, plan.n_long = ptr(keep[0]), long_rows.numel()
            plan.piece_slot, plan.piece_begin, plan.piece_end = ptr(keep[1]), ptr(keep[2]), ptr(keep[3])
            plan.n_pieces, plan.scratch = slot.numel(), ptr(keep[4])
        cached = graph._edge_rows = (er, plan, keep)
    return cached[0], cached[1]


def edge_softmax_rowsum(graph, logit, w, dinv):
    er, _ = edge_rows(graph)
    check(lib().tagrec_edge_softmax_rowsum(ptr(er), er.numel(), graph.n_rows, ptr(logit), ptr(w), ptr(dinv), _st(logit)),
          "tagrec_edge_softmax_rowsum")


def edge_scale(graph, w, dinv, val):
    er, _ = edge_rows(graph)
    check(lib().tagrec_edge_scale(ptr(er), ptr(graph.col), er.numel(), ptr(w), ptr(dinv), ptr(val), _st(w)),
          "tagrec_edge_scale")


def spmm4(graph, val, x, perm=None, res=None, y_raw=None, y_norm=None, mean_acc=None, mean_x0=None, mean_first=False,
          mean_last=False, mean_scale=1.0):
    _, plan = edge_rows(graph)
    check(lib().tagrec_spmm4(ptr(graph.rowptr), ptr(graph.col), graph.n_rows, C.byref(plan) if plan is not None else None,
                             ptr(val), ptr(perm), ptr(x), ptr(res), ptr(y_raw), ptr(y_norm), ptr(mean_acc), ptr(mean_x0),
                             int(mean_first), int(mean_last), float(mean_scale), _st(x)), "tagrec_spmm4")


def edge_dot4(graph, a, b, out, softmax):
    er, _ = edge_rows(graph)
    check(lib().tagrec_edge_dot4(ptr(er), ptr(graph.col), er.numel(), ptr(a), ptr(b), ptr(out), int(bool(softmax)),
                                 _st(a)), "tagrec_edge_dot4")


def chunk_normalize(x, tanh=False, out=None):
    out = torch.empty_like(x) if out is None else out
    check(lib().tagrec_chunk_normalize(ptr(x), x.shape[0], int(bool(tanh)), ptr(out), _st(x)), "tagrec_chunk_normalize")
    return out


def chunk_normalize_bwd(g, x, out=None):
    out = torch.empty_like(x) if out is None else out
    check(lib().tagrec_chunk_normalize_bwd(ptr(g), ptr(x), x.shape[0], ptr(out), _st(x)), "tagrec_chunk_normalize_bwd")
    return out


def _check_table(t):
    assert t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() and t.shape[1] == 16 * FACTORS, \
        "routing kernels need contiguous float32 CUDA tables of width 64 (4 factors x 16)"


def dgcf_propagate(graph, n_layer, iterate_k, ego, collect_weights=False):
    """The launch sequence of DGCF.forward (dgcf.py:49-110).  Returns (mean table, per-layer operator values of the
    last routing iteration, per-layer un-normalised outputs, per-layer softmax weights if ``collect_weights``)."""
    ego = ego.detach().contiguous()
    _check_table(ego)
    dev, n, nnz = ego.device, ego.shape[0], graph._nnz()
    logits = torch.ones((nnz, FACTORS), dtype=torch.float32, device=dev)          # A_values, dgcf.py:50
    w = torch.empty_like(logits)
    dinv = torch.empty((n, FACTORS), dtype=torch.float32, device=dev)
    mean = torch.empty_like(ego)
    tn = torch.empty_like(ego)
    fnorm = torch.empty_like(ego)
    vals, raws, weights = [], [], []
    x = ego
    for layer in range(n_layer):
        chunk_normalize(x, tanh=True, out=tn)                  # tanh(normalize(ego_split[tail])), dgcf.py:106-108
        val = torch.empty_like(logits)
        raw = torch.empty_like(ego)
        nxt = torch.empty_like(ego)
        for t in range(iterate_k):
            edge_softmax_rowsum(graph, logits, w, dinv)
            edge_scale(graph, w, dinv, val)
            last_it = t == iterate_k - 1
            if last_it:
                spmm4(graph, val, x, y_raw=raw, y_norm=nxt, mean_acc=mean, mean_x0=ego, mean_first=layer == 0,
                      mean_last=layer == n_layer - 1, mean_scale=1.0 / (n_layer + 1))
                head = nxt          # normalize(factor_emb[head]) — the chunk-normalised layer output itself
                if collect_weights:
                    weights.append(w.clone())
            else:
                spmm4(graph, val, x, y_norm=fnorm)
                head = fnorm
            if not (last_it and layer == n_layer - 1):         # the very last score update is never read
                edge_dot4(graph, head, tn, logits, softmax=False)
        vals.append(val)
        raws.append(raw)
        x = nxt
    return mean, vals, raws, weights


class DgcfPropagateFn(torch.autograd.Function):
    """DGCF.forward (dgcf.py:49-65): L layers x iterate_k routing iterations over 4 intents, mean over layers.
    Returns the [N, 64] mean table.  Per (layer, iteration): R1 softmax+rowsum, R2 edge values, R3 SpMM over the four
    factor chunks at once, R5+R4 edge-score update of the [nnz, 4] logits — 5 launches instead of >= 12 sparse-tensor
    constructions, 12 sparse.mm and 4 device->host copies."""

    @staticmethod
    def forward(ctx, graph, n_layer, iterate_k, ego):
        mean, vals, raws, _ = dgcf_propagate(graph, n_layer, iterate_k, ego)
        ctx.graph, ctx.n_layer = graph, n_layer
        ctx.vals, ctx.raws = vals, raws
        return mean

    @staticmethod
    def backward(ctx, g_mean):
        graph, n_layer = ctx.graph, ctx.n_layer
        rev = reverse_perm(graph)
        gm = (g_mean / (n_layer + 1)).contiguous()
        g = gm                                   # gradient w.r.t. the (normalised) output of layer L
        for layer in range(n_layer - 1, -1, -1):
            gf = chunk_normalize_bwd(g, ctx.raws[layer])
            g_prev = torch.empty_like(gm)
            spmm4(graph, ctx.vals[layer], gf, perm=rev, res=gm, y_raw=g_prev)     # gm + (D A_k D)^T gf
            g = g_prev
        return None, None, None, g


class DisenRouteFn(torch.autograd.Function):
    """Neighbour routing of one DisenGCN layer (disengcn.py:29-43) on the projected, chunk-normalised factors
    ``fac`` [N, 64]: iterate_k rounds of  p = softmax_k <new[head], fac[tail]>;  new = normalize(fac + P_k fac)."""

    @staticmethod
    def forward(ctx, graph, iterate_k, fac):
        fac = fac.detach().contiguous()
        _check_table(fac)
        nnz = graph._nnz()
        w = torch.empty((nnz, FACTORS), dtype=torch.float32, device=fac.device)
        new = fac
        raw = torch.empty_like(fac)
        bufs = [torch.empty_like(fac), torch.empty_like(fac)]
        for t in range(iterate_k):
            edge_dot4(graph, new, fac, w, softmax=True)
            out = bufs[t % 2]
            spmm4(graph, w, fac, res=fac, y_raw=raw if t == iterate_k - 1 else None, y_norm=out)
            new = out
        ctx.graph = graph
        ctx.save_for_backward(w, raw)
        return new

    @staticmethod
    def backward(ctx, g):
        w, raw = ctx.saved_tensors
        graph = ctx.graph
        gr = chunk_normalize_bwd(g.contiguous(), raw)
        g_fac = torch.empty_like(gr)
        spmm4(graph, w, gr, perm=reverse_perm(graph), res=gr, y_raw=g_fac)         # (I + P^T) gr
        return None, None, g_fac


class ChunkNormFn(torch.autograd.Function):
    """F.normalize over each 16-d factor chunk of a [N, 64] table (disengcn.py:26) — R5 forward, R6 backward."""

    @staticmethod
    def forward(ctx, x):
        x = x.detach().contiguous()
        _check_table(x)
        ctx.save_for_backward(x)
        return chunk_normalize(x)

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        return chunk_normalize_bwd(g.contiguous(), x)
