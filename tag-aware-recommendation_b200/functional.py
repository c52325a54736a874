"""torch.autograd glue over the C ABI: LightGCN propagation (K1) and the fused BPR step (K2).

Everything numeric happens in libtagrec_b200.so; this file only owns buffers (torch allocates, the library never
keeps a pointer) and the order of launches on torch's current stream.
"""
import ctypes as C

import torch

from ._lib import check, lib, ptr, stream_ptr

LOSS_KIND = {"softplus": 0, "logsigmoid": 1}


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


def lightgcn_forward_layers(graph, e0, n_layer, raw, final):
    """lightgcn.py:52-60 — L launches of K1 with the fused normalise + running-mean epilogue.
    raw[k] receives the un-normalised E^{k+1}; ``final`` the mean table."""
    L, st, dim = lib(), stream_ptr(e0.device), e0.shape[1]
    d = graph.desc(dim)
    x = e0
    t = KERNEL_TIMER
    for k in range(n_layer):
        if t:
            t.start("spmm_fwd")
        check(L.tagrec_lightgcn_fwd_layer(C.byref(d), ptr(x), ptr(raw[k]), ptr(final), dim, int(k == 0),
                                          int(k == n_layer - 1), 1.0 / (n_layer + 1), st), "tagrec_lightgcn_fwd_layer")
        if t:
            t.stop("spmm_fwd")
        x = raw[k]
        if graph.comm is not None and k < n_layer - 1:
            graph.comm.all_gather_rows(x)                 # the next layer gathers rows of every rank
    if graph.comm is not None:
        graph.comm.all_gather_rows(final)
    return final


def lightgcn_backward_layers(graph, raw, g_final, n_layer, bufs, g_out, reg_grad=None, upstream=None):
    """Closed-form backward of the above: one elementwise launch (layer L) + L launches of K1 on
    A^T with the normalise-Jacobian epilogue.  ``bufs`` = two scratch tables, ``g_out`` receives dL/dE0."""
    L, st, dim = lib(), stream_ptr(g_final.device), g_final.shape[1]
    d = graph.desc(dim, transposed=True)
    inv = 1.0 / (n_layer + 1)
    g_next = None
    t = KERNEL_TIMER
    for k in range(n_layer, 0, -1):
        out = bufs[k % 2]
        name = "spmm_bwd" if g_next is not None else "bwd_elementwise"
        if t:
            t.start(name)
        check(L.tagrec_lightgcn_bwd_layer(C.byref(d), ptr(g_next), ptr(raw[k - 1]), ptr(g_final), None, ptr(upstream),
                                          inv, ptr(out), dim, st), "tagrec_lightgcn_bwd_layer")
        if t:
            t.stop(name)
        g_next = out
        if graph.comm is not None:
            graph.comm.all_gather_rows(g_next)
    if t:
        t.start("spmm_bwd")
    check(L.tagrec_lightgcn_bwd_layer(C.byref(d), ptr(g_next), None, ptr(g_final), ptr(reg_grad), ptr(upstream), inv,
                                      ptr(g_out), dim, st), "tagrec_lightgcn_bwd_layer")
    if t:
        t.stop("spmm_bwd")
    if graph.comm is not None:
        graph.comm.all_gather_rows(g_out)
    return g_out


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


class LightGCNLossFn(torch.autograd.Function):
    """model.loss(batch) of LightGCN (lightgcn.py:68-82) as ONE autograd node: L fused SpMM launches, one fused BPR
    launch in forward; 1 + L launches in backward."""

    @staticmethod
    def forward(ctx, model, batch, *embeds):
        ws, graph, nl = model._ws, model.norm_adj, model.num_layer
        dev = embeds[0].device
        n, dim = graph.n, embeds[0].shape[1]
        e0 = _buf(ws, "e0", (n, dim), dev)
        torch.cat([e.detach() for e in embeds], dim=0, out=e0)
        raw = [_buf(ws, f"raw{k}", (n, dim), dev) for k in range(nl)]
        final = _buf(ws, "final", (n, dim), dev)
        lightgcn_forward_layers(graph, e0, nl, raw, final)
        g_final = _buf(ws, "g_final", (n, dim), dev)
        g_final.zero_()
        g_reg = None
        if model.reg != 0:
            g_reg = _buf(ws, "g_reg", (n, dim), dev)
            g_reg.zero_()
        loss_out = torch.empty(2, dtype=torch.float32, device=dev)
        bpr_fwd_bwd(batch, model.num_list[0], final, e0, model.reg, model.loss_func, g_final, g_reg, loss_out)
        ctx.model, ctx.has_reg = model, g_reg is not None
        ctx.sizes = [e.shape[0] for e in embeds]
        return loss_out[0], loss_out[1]

    @staticmethod
    def backward(ctx, g_loss, g_regterm):
        model = ctx.model
        ws, graph, nl = model._ws, model.norm_adj, model.num_layer
        g_final = ws["g_final"]
        dev, (n, dim) = g_final.device, g_final.shape
        upstream = torch.stack([g_loss.reshape(()), g_regterm.reshape(())]).to(torch.float32)
        raw = [ws[f"raw{k}"] for k in range(nl)]
        bufs = [_buf(ws, "gbuf0", (n, dim), dev), _buf(ws, "gbuf1", (n, dim), dev)]
        g_e0 = torch.empty((n, dim), dtype=torch.float32, device=dev)
        lightgcn_backward_layers(graph, raw, g_final, nl, bufs, g_e0, ws["g_reg"] if ctx.has_reg else None, upstream)
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
        ctx.save_for_backward(g_final, g_reg if (g_reg is not None and not same) else None)
        return loss_out[0], loss_out[1]

    @staticmethod
    def backward(ctx, g_loss, g_regterm):
        g_final, g_reg = ctx.saved_tensors
        # when the L2 term reads the final table both parts were scattered into one buffer; both upstream
        # gradients are the same scalar in every caller (sum(lossx).backward()), so one multiply suffices
        gf = g_final * g_loss
        gr = None if (g_reg is None or ctx.same) else g_reg * g_regterm
        return None, None, None, None, gf, gr
