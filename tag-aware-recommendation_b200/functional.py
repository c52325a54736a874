"""torch.autograd glue over the C ABI: LightGCN propagation (K1) and the fused BPR step (K2).

Everything numeric happens in libtagrec_b200.so; this file only owns buffers (torch allocates, the library never
keeps a pointer) and the order of launches on torch's current stream.
"""
import ctypes as C

import torch

from ._lib import check, lib, ptr, stream_ptr

LOSS_KIND = {"softplus": 0, "logsigmoid": 1}
MASK_DEPTH = 1           # backward gather launches that skip all-zero source rows (see lightgcn_backward_layers)
import os as _os


def _structural_cut(env, graph):
    """The two structural cuts of a LightGCN training step — last forward layer on the loss's rows only
    (TAGREC_LAST_LAYER_ROWS), push form of the first backward launch's item-row half (TAGREC_PUSH_BWD) — save whole
    launches on graphs where a launch is tens of milliseconds and cost a few small kernels per step, which is a loss on
    the launch-bound small graphs (LastFM-shaped: 0.17 -> 0.35 ms per graphed step).  "0": off, "1" / "force": on,
    unset: on from 2^25 stored entries up (counted over the whole graph of a sharded one)."""
    v = _os.environ.get(env, "auto")
    if v == "0":
        return False
    if v in ("1", "force"):
        return True
    total = getattr(graph, "nnz_global", None) or graph._nnz()
    return total >= (1 << 25)


class KernelTimer:
    """CUDA-event pairs around individual launches on torch's current stream (bench.py's live roofline)."""

    def __init__(self):
        self.pairs = {}

    def start(self, name):
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        self.pairs.setdefault(name, []).append([ev, None])

    def stop(self, name):
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        self.pairs[name][-1][1] = ev

    def mean_ms(self, name):
        p = self.pairs.get(name, [])
        return sum(a.elapsed_time(b) for a, b in p) / len(p) if p else float("nan")

    def count(self, name):
        return len(self.pairs.get(name, []))


KERNEL_TIMER = None      # set by bench.py for the timed region only


def _buf(ws, name, shape, device, dtype=torch.float32):
    t = ws.get(name)
    if t is None or t.shape != torch.Size(shape) or t.device != device or t.dtype != dtype:
        t = torch.empty(shape, dtype=dtype, device=device)
        ws[name] = t
    return t


def _mref(m):
    return C.byref(m) if m is not None else None


def lightgcn_forward_layers(graph, e0, n_layer, raw, final, mirrors=None, last_rows=None):
    """lightgcn.py:52-60 — L launches of K1 with the fused normalise + running-mean epilogue.
    raw[k] receives the un-normalised E^{k+1}; ``final`` the mean table.
    Sharded graphs: ``mirrors`` (dict id(tensor) -> MirrorDesc, tables in symmetric memory) selects the fused
    peer-store all-gather + barrier; without it the row blocks are all-gathered with NCCL after each launch.
    ``last_rows`` (int32 local row ids, ascending, unique): the LAST layer is produced on these rows only — a training
    step reads the mean table at the batch's rows and nothing else of the last layer (lightgcn.py:68-75; its backward
    needs E^L at the same rows), so raw[L-1] and ``final`` are defined on those rows only afterwards."""
    L, st, dim = lib(), stream_ptr(e0.device), e0.shape[1]
    d = graph.desc(dim)
    comm = graph.comm
    x = e0
    t = KERNEL_TIMER
    for k in range(n_layer):
        last = k == n_layer - 1
        my = mirrors.get(id(raw[k])) if (mirrors and not last) else None
        ma = mirrors.get(id(final)) if (mirrors and last) else None
        dk, undo, label = d, None, "spmm_fwd"
        if last and last_rows is not None:
            dk, undo = graph.subset_desc(dim, last_rows)
            label = "spmm_fwd_rows"
        if t:
            t.start(label)
        check(L.tagrec_lightgcn_fwd_layer_p2p(C.byref(dk), ptr(x), ptr(raw[k]), ptr(final), dim, int(k == 0), int(last),
                                              1.0 / (n_layer + 1), _mref(my), _mref(ma), st),
              "tagrec_lightgcn_fwd_layer")
        if undo is not None:
            undo()
        if t:
            t.stop(label)
        x = raw[k]
        if comm is not None:
            if mirrors:
                if t:
                    t.start("barrier")
                comm.peer.barrier(mirrors["name", id(final if last else x)])
                if t:
                    t.stop("barrier")
            else:
                if not last:
                    comm.all_gather_rows(x)                 # the next layer gathers rows of every rank
                else:
                    comm.all_gather_rows(final)
    return final


def lightgcn_backward_layers(graph, raw, g_final, n_layer, bufs, g_out, reg_grad=None, upstream=None, mirrors=None,
                             batch_nodes=None, mask_depth=0, local_out=False, adam=None):
    """Closed-form backward of the above: the first table G_L (no gather) + L launches of K1 on A^T with
    the normalise-Jacobian epilogue.  ``bufs`` = two scratch tables, ``g_out`` receives dL/dE0.

    ``batch_nodes`` (int64 global row ids of the batch's users and items): dL/dF — hence G_L — is non-zero on those rows
    only, so G_L is produced by the O(batch) sparse kernel into a table this graph keeps all-zero, flagged in a byte
    map, consumed by the masked K1 launch (the 256 B gathers of all-zero rows are never issued) and cleared again.
    Without it (a generic upstream gradient) G_L is an elementwise pass over all rows.
    Sharded graphs: ``mirrors`` selects the fused peer-store exchange (outputs stored to every rank by the kernels);
    G_L's few rows are summed across ranks (each row is non-zero on its owner only).  ``local_out``: dL/dE0 is needed on
    this rank's rows only (owner-sharded optimizer) — the last launch is not exchanged at all.  ``adam`` (an
    ``_lib.AdamDesc``): the optimizer step runs in the epilogue of the last launch (tagrec_lightgcn_bwd_layer_adam): the
    gradient rows are consumed where they are produced and the new parameter rows go to every rank; ``g_out`` may then
    be None."""
    L, st, dim = lib(), stream_ptr(g_final.device), g_final.shape[1]
    # Sharded graphs may carry a SECOND row block for the backward launches (graph.bwd_graph, see
    # distributed.split_partition_by_measurement): every table is full-size on every rank, so each launch may cut the
    # rows its own way, and forward and backward launches of a row block do not cost the same.  The first (sparse)
    # table stays with the forward owner: it reads the last raw layer, which is not exchanged.
    gb = getattr(graph, "bwd_graph", None) or graph
    d = gb.desc(dim, transposed=True)
    d_masked = gb.desc(dim, transposed=True, plain=True)         # the masked launch keeps the plain plan (issue-bound)
    comm, comm_b = graph.comm, gb.comm
    inv = 1.0 / (n_layer + 1)
    t = KERNEL_TIMER
    ws = graph.__dict__.setdefault("_bwd_ws", {})
    n = g_final.shape[0]
    sparse = batch_nodes is not None
    mk = None
    if sparse:
        g_next = _zero_buf(ws, "g_sparse", (n, dim), g_final.device)
        mk = _zero_buf(ws, "nz_mask", (n,), g_final.device, torch.uint8) if mask_depth > 0 else None
        lo, hi = (comm.lo, comm.hi) if comm is not None else (0, n)
        if t:
            t.start("bwd_first")
        check(L.tagrec_lightgcn_bwd_first_sparse(ptr(batch_nodes), batch_nodes.numel(), lo, hi, ptr(raw[n_layer - 1]),
                                                 ptr(g_final), ptr(upstream), inv, ptr(g_next), ptr(mk), dim, st),
              "tagrec_lightgcn_bwd_first_sparse")
        if comm is not None and comm.world > 1:
            rows = g_next.index_select(0, batch_nodes)            # zero where another rank owns the node
            torch.distributed.all_reduce(rows, group=comm.group)
            g_next.index_copy_(0, batch_nodes, rows)
        if t:
            t.stop("bwd_first")
    else:
        g_next = bufs[n_layer % 2]
        m = mirrors.get(id(g_next)) if mirrors else None
        if t:
            t.start("bwd_first")
        d_first = graph.desc(dim, transposed=True)               # forward owner's rows (see above)
        check(L.tagrec_lightgcn_bwd_layer_ex(C.byref(d_first), None, None, ptr(raw[n_layer - 1]), ptr(g_final), None,
                                             ptr(upstream), inv, ptr(g_next), dim, _mref(m), st),
              "tagrec_lightgcn_bwd_layer")
        if t:
            t.stop("bwd_first")
        if comm is not None:
            if mirrors:
                comm.peer.barrier(mirrors["name", id(g_next)])
            else:
                comm.all_gather_rows(g_next)

    def clear_sparse():
        check(L.tagrec_rows_zero(ptr(batch_nodes), batch_nodes.numel(), ptr(ws["g_sparse"]), ptr(mk), dim, st),
              "tagrec_rows_zero")

    # Unsharded bipartite graph, first gather launch of a BPR step: its source table is non-zero on the batch's rows
    # only.  The USER-row half sums over item sources — the batch's items, popular ones among them — and stays a masked
    # gather (on the user-row block); the ITEM-row half sums over USER sources, of which only the batch's <= B users
    # count: instead of scanning every stored entry of the item rows for them (half of the launch), those users PUSH
    # their rows into an accumulation table (B x ~100 entries) and the item rows run the epilogue alone.
    halves = graph.halves() if (sparse and mk is not None and comm is None and _structural_cut("TAGREC_PUSH_BWD", graph)) else None
    first_gather = True
    for k in range(n_layer - 1, 0, -1):
        out = bufs[k % 2]
        m = mirrors.get(id(out)) if mirrors else None
        if t:
            t.start("spmm_bwd")
        use_mask = first_gather and mk is not None
        if use_mask and halves is not None:
            ub, ib = halves
            # the pushing rows = the batch's USER nodes, each once: sorted node list with a keep flag (fixed size, no
            # host sync: the step stays CUDA-graph capturable)
            srt = torch.sort(batch_nodes).values
            keep = srt < int(graph.num_list[0])
            keep[1:] &= srt[1:] != srt[:-1]
            keep = keep.to(torch.uint8)
            acc_tab = _zero_buf(ws, "push_acc", (n, dim), g_final.device)
            check(L.tagrec_spmm_push_rows(ptr(graph.rowptr), ptr(graph.col), ptr(graph.val), ptr(srt), ptr(keep),
                                          srt.numel(), ptr(g_next), ptr(acc_tab), dim, st), "tagrec_spmm_push_rows")
            du = ub.desc(dim, transposed=True, plain=True)
            check(L.tagrec_lightgcn_bwd_layer_ex(C.byref(du), ptr(g_next), ptr(mk), ptr(raw[k - 1]), ptr(g_final), None,
                                                 ptr(upstream), inv, ptr(out), dim, None, st), "tagrec_lightgcn_bwd_layer")
            di = ib.desc(dim, transposed=True)
            check(L.tagrec_lightgcn_bwd_layer_acc(C.byref(di), ptr(acc_tab), ptr(raw[k - 1]), ptr(g_final), ptr(upstream),
                                                  inv, ptr(out), dim, None, st), "tagrec_lightgcn_bwd_layer_acc")
        else:
            check(L.tagrec_lightgcn_bwd_layer_ex(C.byref(d_masked if use_mask else d), ptr(g_next), ptr(mk) if use_mask else None, ptr(raw[k - 1]),
                                                 ptr(g_final), None, ptr(upstream), inv, ptr(out), dim, _mref(m), st),
                  "tagrec_lightgcn_bwd_layer")
        if t:
            t.stop("spmm_bwd")
        if first_gather and sparse:
            clear_sparse()
        first_gather = False
        g_next = out
        if comm is not None:
            if mirrors:
                comm.peer.barrier(mirrors["name", id(out)])
            else:
                comm_b.all_gather_rows(g_next)
    m = mirrors.get(id(g_out)) if (mirrors and not local_out) else None
    if t:
        t.start("spmm_bwd")
    use_mask = first_gather and mk is not None
    if adam is not None:
        check(L.tagrec_lightgcn_bwd_layer_adam(C.byref(d_masked if use_mask else d), ptr(g_next), ptr(mk) if use_mask else None,
                                               ptr(g_final), ptr(reg_grad), ptr(upstream), inv, ptr(g_out), dim,
                                               C.byref(adam), st), "tagrec_lightgcn_bwd_layer_adam")
    else:
        check(L.tagrec_lightgcn_bwd_layer_ex(C.byref(d_masked if use_mask else d), ptr(g_next), ptr(mk) if use_mask else None, None, ptr(g_final),
                                             ptr(reg_grad), ptr(upstream), inv, ptr(g_out), dim, _mref(m), st),
              "tagrec_lightgcn_bwd_layer")
    if t:
        t.stop("spmm_bwd")
    if first_gather and sparse:
        clear_sparse()
    if comm is not None and not local_out:
        if mirrors:
            comm.peer.barrier(mirrors["name", id(g_out)])
        else:
            comm_b.all_gather_rows(g_out)
    return g_out


def _zero_buf(ws, name, shape, device, dtype=torch.float32):
    """A persistent table that is all-zero between uses (its users restore that state themselves)."""
    tns = ws.get(name)
    if tns is None or tns.shape != torch.Size(shape) or tns.device != device or tns.dtype != dtype:
        tns = torch.zeros(shape, dtype=dtype, device=device)
        ws[name] = tns
    return tns


def bpr_fwd_bwd(batch, item_offset, final, reg_src, reg, loss_kind, g_final, g_reg, loss_out):
    """K2: loss_out[0:2] = (loss, reg * reg_loss); g_final / g_reg += gradients (caller zeroes them)."""
    assert batch.dtype == torch.int64 and batch.dim() == 2 and batch.shape[1] == 3
    batch = batch.contiguous()
    t = KERNEL_TIMER
    if t:
        t.start("bpr")
    check(lib().tagrec_bpr_fwd_bwd(ptr(batch), batch.shape[0], item_offset, ptr(final), ptr(reg_src), final.shape[1],
                                   float(reg), LOSS_KIND[loss_kind], ptr(g_final), ptr(g_reg), ptr(loss_out),
                                   stream_ptr(final.device)), "tagrec_bpr_fwd_bwd")
    if t:
        t.stop("bpr")


def _tables(model, names, n, dim, dev):
    """Persistent [n, dim] work tables of a model.  On a sharded graph with the peer path enabled, the tables that
    other ranks write into live in symmetric memory and come with their MirrorDesc."""
    ws, comm = model._ws, model.norm_adj.comm
    peer = comm.peer if comm is not None else None
    out, mirrors = {}, ({} if peer is not None else None)
    for name, shared in names:
        if peer is not None and shared:
            tns, m = peer.table(name, (n, dim))
            mirrors[id(tns)] = m
            mirrors["name", id(tns)] = name
        else:
            tns = _buf(ws, name, (n, dim), dev)
        out[name] = tns
    return out, mirrors


class LightGCNLossFn(torch.autograd.Function):
    """model.loss(batch) of LightGCN (lightgcn.py:68-82) as ONE autograd node: L fused SpMM launches, one fused BPR
    launch in forward; 1 + L launches in backward."""

    @staticmethod
    def forward(ctx, model, batch, *embeds):
        ws, graph, nl = model._ws, model.norm_adj, model.num_layer
        dev = embeds[0].device
        n, dim = graph.n, embeds[0].shape[1]
        e0 = model._flat_params()                                   # [n, dim] storage the parameters are views of
        names = [(f"raw{k}", k < nl - 1) for k in range(nl)] + [("final", True)]
        tabs, mirrors = _tables(model, names, n, dim, dev)
        raw, final = [tabs[f"raw{k}"] for k in range(nl)], tabs["final"]
        ws["raw_list"] = raw
        batch = batch.contiguous()
        nodes = torch.cat([batch[:, 0], batch[:, 1] + model.num_list[0], batch[:, 2] + model.num_list[0]])
        last_rows = None
        if _structural_cut("TAGREC_LAST_LAYER_ROWS", graph):
            # the last layer on the batch's rows only (this rank's share of them).  Fixed-size list, no host sync (also
            # valid under CUDA-graph capture): sorted nodes, duplicates and other ranks' rows blanked with -1
            srt = torch.sort(nodes).values
            lo, hi = (graph.comm.lo, graph.comm.hi) if graph.comm is not None else (0, n)
            bad = (srt < lo) | (srt >= hi)
            bad[1:] |= srt[1:] == srt[:-1]
            last_rows = torch.where(bad, torch.full_like(srt, -1), srt - lo).to(torch.int32)
        lightgcn_forward_layers(graph, e0, nl, raw, final, mirrors, last_rows=last_rows)
        # The gradient tables are all-zero between steps: K2 scatters into the batch's rows, backward() consumes them
        # and re-zeroes exactly those rows (no state crosses a step, so eager and CUDA-graph steps can alternate).
        # If a previous forward was never followed by its backward, its rows are cleared here first.
        for name in ("g_final", "g_reg"):
            if name == "g_reg" and model.reg == 0:
                continue
            tns = ws.get(name)
            if tns is None or tns.shape != (n, dim) or tns.device != dev:
                ws[name] = torch.zeros((n, dim), dtype=torch.float32, device=dev)
            elif ws.get("pending_nodes") is not None:
                tns.index_fill_(0, ws["pending_nodes"], 0.0)
        ws["pending_nodes"] = nodes
        ws["generation"] = ws.get("generation", 0) + 1             # the work tables now belong to THIS forward
        g_final = ws["g_final"]
        g_reg = ws["g_reg"] if model.reg != 0 else None
        loss_out = torch.empty(2, dtype=torch.float32, device=dev)
        bpr_fwd_bwd(batch, model.num_list[0], final, e0, model.reg, model.loss_func, g_final, g_reg, loss_out)
        ctx.model, ctx.has_reg, ctx.nodes, ctx.generation = model, g_reg is not None, nodes, ws["generation"]
        ctx.sizes = [e.shape[0] for e in embeds]
        return loss_out[0], loss_out[1]

    @staticmethod
    def backward(ctx, g_loss, g_regterm):
        model = ctx.model
        ws, graph, nl = model._ws, model.norm_adj, model.num_layer
        if ws.get("generation") != ctx.generation:
            raise RuntimeError("LightGCN.loss() was called again before this loss was back-propagated: the saved layer "
                               "tables and gradient buffers are shared work space and now hold the later call's data "
                               "(run backward() right after loss(), as training/basic_train.py:18-24 does)")
        g_final = ws["g_final"]
        dev, (n, dim) = g_final.device, g_final.shape
        upstream = torch.stack([g_loss.reshape(()), g_regterm.reshape(())]).to(torch.float32)
        raw = ws["raw_list"]
        p2p = graph.comm is not None and graph.comm.peer is not None
        local_out = bool(ws.get("local_grad_only")) and graph.comm is not None
        tabs, mirrors = _tables(model, [("gbuf0", True), ("gbuf1", True)] + ([("g_e0", True)] if p2p and not local_out else []),
                                n, dim, dev)
        # single GPU / NCCL: a fresh table per step (the parameters' .grad are views of it and may outlive the step);
        # fused exchange: the symmetric-memory table other ranks store into
        # optimizer folded into the last launch (optim.ShardedFusedAdam(fused_backward=True)): no gradient table at all
        hook = ws.get("adam_epilogue")
        adam = hook.begin_fused_step() if hook is not None else None
        if adam is not None:
            g_e0 = None
        else:
            g_e0 = tabs["g_e0"] if (p2p and not local_out) else torch.empty((n, dim), dtype=torch.float32, device=dev)
        lightgcn_backward_layers(graph, raw, g_final, nl, [tabs["gbuf0"], tabs["gbuf1"]], g_e0,
                                 ws["g_reg"] if ctx.has_reg else None, upstream, mirrors,
                                 batch_nodes=ctx.nodes, mask_depth=MASK_DEPTH, local_out=local_out, adam=adam)
        g_final.index_fill_(0, ctx.nodes, 0.0)
        if ctx.has_reg:
            ws["g_reg"].index_fill_(0, ctx.nodes, 0.0)
        ws["pending_nodes"] = None
        if adam is not None:
            return (None, None) + (None,) * len(ctx.sizes)          # the parameters were updated in place; no .grad
        return (None, None) + tuple(torch.split(g_e0, ctx.sizes, dim=0))


class LightGCNPropagateFn(torch.autograd.Function):
    """model.forward() of LightGCN (lightgcn.py:49-63) as a differentiable [N, dim] table."""

    @staticmethod
    def forward(ctx, model, *embeds):
        graph, nl = model.norm_adj, model.num_layer
        e0 = torch.cat([e.detach() for e in embeds], dim=0)
        raw = [torch.empty_like(e0) for _ in range(nl)]
        final = torch.empty_like(e0)
        lightgcn_forward_layers(graph, e0, nl, raw, final)
        ctx.model, ctx.raw = model, raw
        ctx.sizes = [e.shape[0] for e in embeds]
        return final

    @staticmethod
    def backward(ctx, g):
        model = ctx.model
        g = g.contiguous()
        bufs = [torch.empty_like(g), torch.empty_like(g)]
        g_e0 = torch.empty_like(g)
        lightgcn_backward_layers(model.norm_adj, ctx.raw, g, model.num_layer, bufs, g_e0)
        return (None,) + tuple(torch.split(g_e0, ctx.sizes, dim=0))


class BprLossFn(torch.autograd.Function):
    """mul_loss + l2reg_loss (loss.py:4-12,27-32) on row tables, for models whose propagation is composed from
    primitives (NGCF, ...).  ``reg_src`` is the table the L2 term reads; may be ``final`` itself."""

    @staticmethod
    def forward(ctx, batch, item_offset, reg, loss_kind, final, reg_src):
        final_c = final.detach().contiguous()
        same = reg_src is final
        src_c = final_c if same else reg_src.detach().contiguous()
        g_final = torch.zeros_like(final_c)
        g_reg = None
        if reg != 0:
            g_reg = g_final if same else torch.zeros_like(src_c)
        loss_out = torch.empty(2, dtype=torch.float32, device=final.device)
        bpr_fwd_bwd(batch, item_offset, final_c, src_c, reg, loss_kind, g_final, g_reg, loss_out)
        ctx.same = same
        ctx.merged = same and reg != 0
        ctx.save_for_backward(g_final, g_reg if (g_reg is not None and not same) else None)
        return loss_out[0], loss_out[1]

    @staticmethod
    def backward(ctx, g_loss, g_regterm):
        g_final, g_reg = ctx.saved_tensors
        # when the L2 term reads the final table both parts were scattered into ONE buffer, which is only right when
        # both upstream gradients are the same scalar (sum(lossx).backward(), basic_train.py:18).  A caller that weights
        # the two parts differently gets NaN gradients instead of silently wrong ones (no host sync for the check).
        if ctx.merged and g_regterm is not None and g_regterm is not g_loss:
            g_loss = torch.where(g_loss == g_regterm, g_loss, torch.full_like(g_loss, float("nan")))
        gf = g_final * g_loss
        gr = None if (g_reg is None or ctx.same) else g_reg * g_regterm
        return None, None, None, None, gf, gr


def xty(x, y):
    """x[n, a]^T y[n, b] on K8 (csrc/xty.cu): the weight gradient of a small dense layer as a row reduction over all
    SMs (cuBLAS runs the [a x n] x [n x b] form on one tile's worth of SMs)."""
    x, y = x.contiguous(), y.contiguous()
    out = torch.empty((x.shape[1], y.shape[1]), dtype=torch.float32, device=x.device)
    check(lib().tagrec_xty(ptr(x), ptr(y), x.shape[0], x.shape[1], y.shape[1], ptr(out), stream_ptr(x.device)),
          "tagrec_xty")
    return out


class SkinnyMmFn(torch.autograd.Function):
    """x[n, a] @ w[a, b] for tall x and a, b <= 64 (attention projections tgcn.py:26-31, factor projection
    disengcn.py:25).  Forward and input gradient are plain GEMMs (the latter on a contiguous copy of w^T: cuBLAS picks
    a 6x slower kernel for the transposed-operand form with K = 32); the weight gradient runs on K8."""

    @staticmethod
    def forward(ctx, x, w):
        ctx.save_for_backward(x, w)
        return torch.mm(x, w)

    @staticmethod
    def backward(ctx, g):
        x, w = ctx.saved_tensors
        g = g.contiguous()
        gx = torch.mm(g, w.t().contiguous()) if ctx.needs_input_grad[0] else None
        gw = xty(x.detach(), g) if ctx.needs_input_grad[1] else None
        return gx, gw


def skinny_mm(x, w):
    ok = x.is_cuda and x.dtype == torch.float32 and x.dim() == 2 and all(4 <= d <= 64 and d % 4 == 0 for d in w.shape)
    return SkinnyMmFn.apply(x, w) if ok else torch.mm(x, w)


class NgcfDenseFn(torch.autograd.Function):
    """K6: the dense half of one NGCF layer (ngcf.py:77-86) as one autograd node — one fused forward launch, one
    fused backward launch (input gradients AND the two 64x64 weight gradients)."""

    @staticmethod
    def forward(ctx, nei, e, w1, b1, w2, b2):
        nei_c, e_c = nei.detach().contiguous(), e.detach().contiguous()
        w1c, b1c, w2c, b2c = (t.detach().contiguous() for t in (w1, b1, w2, b2))
        out, nrm = torch.empty_like(e_c), torch.empty_like(e_c)
        s_act, t_act = torch.empty_like(e_c), torch.empty_like(e_c)
        n, dim = e_c.shape
        check(lib().tagrec_ngcf_dense_fwd(ptr(nei_c), ptr(e_c), ptr(w1c), ptr(b1c), ptr(w2c), ptr(b2c), n, dim, ptr(out),
                                          ptr(nrm), ptr(s_act), ptr(t_act), stream_ptr(e_c.device)),
              "tagrec_ngcf_dense_fwd")
        ctx.save_for_backward(nei_c, e_c, w1c, b1c, w2c, b2c, out, s_act, t_act)
        return out, nrm

    @staticmethod
    def backward(ctx, g_out, g_nrm):
        nei, e, w1, b1, w2, b2, out, s_act, t_act = ctx.saved_tensors
        n, dim = e.shape
        if g_nrm is None:
            g_nrm = torch.zeros_like(e)
        if g_nrm.stride(1) != 1 or g_nrm.stride(0) % 4 != 0 or g_nrm.data_ptr() % 16 != 0:
            g_nrm = g_nrm.contiguous()
        if g_out is not None:
            g_out = g_out.contiguous()
        g_nei, g_e = torch.empty_like(e), torch.empty_like(e)
        dw = torch.zeros((2, dim, dim), dtype=torch.float32, device=e.device)       # d(W1 + b1), d(W2 + b2)
        check(lib().tagrec_ngcf_dense_bwd(ptr(g_out), ptr(g_nrm), g_nrm.stride(0), ptr(out), ptr(s_act), ptr(t_act),
                                          ptr(nei), ptr(e), ptr(w1), ptr(b1), ptr(w2), ptr(b2), n, dim, ptr(g_nei),
                                          ptr(g_e), None, None, ptr(dw[0]), ptr(dw[1]), stream_ptr(e.device)),
              "tagrec_ngcf_dense_bwd")
        # the bias is broadcast over the ROWS of W (ngcf.py:78): its gradient is the column sum of dW
        gb = dw.sum(1, keepdim=True)
        return g_nei, g_e, dw[0], gb[0], dw[1], gb[1]
