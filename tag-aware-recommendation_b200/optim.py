"""Fused Adam optimizer over the C ABI (tagrec_adam_step): one pass over param/grad/exp_avg/exp_avg_sq per tensor.

Optional replacement for the ``optim.Adam(model.parameters(), lr=CFG['lr'])`` the reference composes in com.py:25
(SURVEY §8 f-3).  Same update rule and state names as torch.optim.Adam (amsgrad=False), so ``state_dict()`` of the
optimizer stays loadable by torch's Adam; checked against it in tests/test_gpu_parity.py.
"""
import torch

from ._lib import check, lib, ptr, stream_ptr


class FusedAdam(torch.optim.Optimizer):
    """``capturable=True`` keeps the step counter and the bias corrections in device memory (one extra 1-thread launch
    per step) so that ``step()`` can be recorded into a CUDA graph (``graph_step.GraphedStep``)."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, capturable=False):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, capturable=capturable))

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        self.__dict__.pop("_dev", None)         # device-side counters are re-seeded from the loaded ``step``

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        L = lib()
        dev_state = self.__dict__.setdefault("_dev", {})
        for gi, group in enumerate(self.param_groups):
            b1, b2 = group["betas"]
            cap = group.get("capturable", False)
            scal = None
            for p in group["params"]:
                if p.grad is None:
                    continue
                if not (p.is_cuda and p.dtype == torch.float32 and p.is_contiguous()):
                    raise RuntimeError("FusedAdam needs contiguous float32 CUDA parameters (no CPU fallback)")
                st = self.state[p]
                if not st:
                    st["step"] = torch.zeros((), dtype=torch.int64, device=p.device) if cap else 0
                    st["exp_avg"] = torch.zeros_like(p)
                    st["exp_avg_sq"] = torch.zeros_like(p)
                for key in ("exp_avg", "exp_avg_sq"):          # a state loaded from a checkpoint may live elsewhere
                    if st[key].device != p.device or st[key].dtype != torch.float32 or not st[key].is_contiguous():
                        st[key] = st[key].to(device=p.device, dtype=torch.float32).contiguous()
                g = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
                if cap:
                    if scal is None:
                        # one device-side counter per group (every tensor of the group steps together)
                        gs = dev_state.setdefault(gi, {})
                        if "step" not in gs:
                            # resume: start from the counter a loaded state_dict carries (torch's Adam stores a
                            # float tensor, this class an int or an int64 tensor), not from zero
                            gs["step"] = torch.full((), int(st["step"]), dtype=torch.int64, device=p.device)
                            gs["scal"] = torch.zeros(2, dtype=torch.float32, device=p.device)
                        check(L.tagrec_adam_advance(ptr(gs["step"]), group["lr"], b1, b2, ptr(gs["scal"]),
                                                    stream_ptr(p.device)), "tagrec_adam_advance")
                        scal = gs["scal"]
                    st["step"] = dev_state[gi]["step"]
                    check(L.tagrec_adam_step_dev(ptr(p), ptr(g), ptr(st["exp_avg"]), ptr(st["exp_avg_sq"]), p.numel(), b1,
                                                 b2, group["eps"], group["weight_decay"], ptr(scal), stream_ptr(p.device)),
                          "tagrec_adam_step_dev")
                else:
                    st["step"] = int(st["step"]) + 1           # int(): a state loaded from torch's Adam holds a tensor
                    check(L.tagrec_adam_step(ptr(p), ptr(g), ptr(st["exp_avg"]), ptr(st["exp_avg_sq"]), p.numel(),
                                             group["lr"], b1, b2, group["eps"], group["weight_decay"], st["step"],
                                             stream_ptr(p.device)), "tagrec_adam_step")
                # the kernel wrote the parameter through a raw pointer: tell autograd / version-keyed caches
                torch._C._increment_version([p])
        return loss


class ShardedFusedAdam(torch.optim.Optimizer):
    """Owner-sharded Adam for a LightGCN on a node-range sharded graph with the fused exchange (multi-GPU only; no
    reference equivalent — the reference is single-device).

    The embedding tables stay REPLICATED (``model.state_dict()`` is full-size on every rank, SURVEY §8 e), but each row
    is updated by the rank that owns it: rank p runs the Adam step on rows [lo_p, hi_p) only and the kernel stores the
    new parameter rows into every rank's replica (tagrec_adam_step_mirror: NVLS multicast / peer stores), followed by
    one device-side barrier.  Per step this replaces 7 x N x dim x 4 bytes of replicated optimizer traffic on every
    GPU by 1/P of it, and the last backward launch no longer has to exchange dL/dE0 at all — a rank only needs the
    gradient rows it owns (``model._ws['local_grad_only']``; ``p.grad`` is defined on the owned rows only).
    ``exp_avg`` / ``exp_avg_sq`` are full-size tensors whose owned rows are live; ``consolidate()`` all-gathers them so
    that ``state_dict()`` is complete on every rank before a checkpoint.

    ``fused_backward`` (default; TAGREC_ADAM_EPILOGUE=0 switches it off): the update runs in the EPILOGUE of the last
    backward launch (tagrec_lightgcn_bwd_layer_adam) — each dL/dE0 row is consumed where K1 produces it and the new
    parameter row is stored to every rank from there, so the parameter exchange (N x dim x 4 bytes into every GPU per
    step: 4 ms of exposed NVLink time at 8 GPUs as a separate pass) overlaps the gathers of that launch.  ``backward()``
    then applies step t + 1 and ``step()`` only closes it (counter, barrier); ``p.grad`` stays None.  One ``backward()``
    per ``step()`` — a second one raises."""

    def __init__(self, model, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, fused_backward=None):
        params = list(model.parameters())
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        graph = model.norm_adj
        graph = getattr(graph, "bwd_graph", None) or graph       # rows are owned by the rank whose BACKWARD produces them
        comm = getattr(graph, "comm", None)
        single = comm is None                    # one GPU: the same update in the epilogue, nothing to exchange
        if not single and (comm.peer is None or comm.world < 2):
            raise ValueError("ShardedFusedAdam needs a model on one GPU or on a sharded graph with the fused exchange "
                             "enabled (graph.comm.enable_p2p); use FusedAdam otherwise")
        if [id(p) for p in model.parameters()] != [id(p) for p in model.embed]:
            raise ValueError("ShardedFusedAdam handles row tables only (LightGCN)")
        self.model, self.comm = model, comm
        flat = model._flat_params()
        n, dim = flat.shape
        if single:
            from ._lib import MirrorDesc
            if fused_backward is False:
                raise ValueError("on one GPU ShardedFusedAdam only exists as the backward epilogue; use FusedAdam")
            fused_backward = True
            mirror = MirrorDesc()                # n == 0: local table only
        else:
            table, mirror = comm.peer.table("e0", (n, dim))      # the parameters move into symmetric memory
            table.copy_(flat)
            off = 0
            for p in model.embed:
                p.data = table[off:off + p.shape[0]]
                off += p.shape[0]
            model._ws["flat"] = table
            model._ws["local_grad_only"] = True
        self._mirror = mirror
        self._m = torch.zeros((n, dim), dtype=torch.float32, device=flat.device)
        self._v = torch.zeros((n, dim), dtype=torch.float32, device=flat.device)
        self._step = 0
        import os
        if fused_backward is None:
            fused_backward = os.environ.get("TAGREC_ADAM_EPILOGUE", "1") != "0"
        self._applied = False                    # backward() already applied the update of the step being taken
        if fused_backward:
            model._ws["adam_epilogue"] = self
        off = 0
        for p in model.embed:
            self.state[p] = {"step": 0, "exp_avg": self._m[off:off + p.shape[0]], "exp_avg_sq": self._v[off:off + p.shape[0]]}
            off += p.shape[0]
        if comm is not None:
            comm.peer.barrier("e0")

    def load_state_dict(self, state_dict):
        """Resume: torch's loader replaces the state tensors by copies — move their contents into the flat exp_avg /
        exp_avg_sq tables the kernels use (the state entries are views of those again afterwards) and restart the step
        counter (bias corrections) from the loaded one."""
        super().load_state_dict(state_dict)
        off, step = 0, 0
        with torch.no_grad():
            for p in self.model.embed:
                st = self.state.get(p, {})
                rows = slice(off, off + p.shape[0])
                if "exp_avg" in st and st["exp_avg"].data_ptr() != self._m[rows].data_ptr():
                    self._m[rows].copy_(st["exp_avg"])
                    self._v[rows].copy_(st["exp_avg_sq"])
                step = max(step, int(st.get("step", 0)))
                self.state[p] = {"step": step, "exp_avg": self._m[rows], "exp_avg_sq": self._v[rows]}
                off += p.shape[0]
        self._step = step
        self._applied = False

    def begin_fused_step(self):
        """Called by LightGCNLossFn.backward: the tagrec_adam_t of step t + 1 for the epilogue of its last launch."""
        from ._lib import AdamDesc
        if self._applied:
            raise RuntimeError("ShardedFusedAdam(fused_backward=True) applies the update inside backward(): call step() "
                               "after every backward() (gradient accumulation needs fused_backward=False)")
        group = self.param_groups[0]
        d = AdamDesc()
        d.param, d.exp_avg, d.exp_avg_sq = ptr(self.model._flat_params()), ptr(self._m), ptr(self._v)
        d.lr, (d.beta1, d.beta2), d.eps, d.weight_decay = group["lr"], group["betas"], group["eps"], group["weight_decay"]
        d.step = self._step + 1
        d.param_mirror = self._mirror
        self._applied = True
        return d

    @torch.no_grad()
    def step(self, closure=None):
        import ctypes as C
        from . import functional as Fn
        from ._lib import MirrorDesc
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        L, comm, model = lib(), self.comm, self.model
        group = self.param_groups[0]
        b1, b2 = group["betas"]
        self._step += 1
        if self._applied:                        # the epilogue of the last backward launch did the arithmetic
            self._applied = False
            t = Fn.KERNEL_TIMER
            if t:
                t.start("adam")
            for p in model.embed:
                self.state[p]["step"] = self._step
                torch._C._increment_version([p])
            if comm is not None:
                comm.peer.barrier("e0")          # every rank's replica is complete before the next forward gathers it
            if t:
                t.stop("adam")
            return loss
        if comm is None:
            raise RuntimeError("ShardedFusedAdam on one GPU: step() without a backward() through model.loss()")
        table = model._ws["flat"]
        dim = table.shape[1]
        t = Fn.KERNEL_TIMER
        if t:
            t.start("adam")
        off = 0
        for p in model.embed:
            lo, hi = max(comm.lo, off), min(comm.hi, off + p.shape[0])       # owned rows of this parameter
            if hi > lo and p.grad is not None:
                g = p.grad
                if not (g.is_contiguous() and g.dtype == torch.float32):
                    raise RuntimeError("ShardedFusedAdam needs contiguous float32 gradients")
                a, cnt = lo - off, (hi - lo) * dim
                seg = MirrorDesc()
                seg.n, seg.self = self._mirror.n, self._mirror.self
                for r in range(self._mirror.n):
                    seg.base[r] = self._mirror.base[r] + lo * dim * 4
                check(L.tagrec_adam_step_mirror(ptr(p[a:]), ptr(g[a:]), ptr(self._m[lo:]), ptr(self._v[lo:]), cnt,
                                                group["lr"], b1, b2, group["eps"], group["weight_decay"], self._step,
                                                C.byref(seg), stream_ptr(p.device)), "tagrec_adam_step_mirror")
            self.state[p]["step"] = self._step
            torch._C._increment_version([p])
            off += p.shape[0]
        comm.peer.barrier("e0")                  # every rank's replica is complete before the next forward gathers it
        if t:
            t.stop("adam")
        return loss

    def consolidate(self):
        """All-gather the optimizer state row blocks (collective) — call before ``state_dict()`` / a checkpoint."""
        if self.comm is None:
            return
        self.comm.all_gather_rows(self._m)
        self.comm.all_gather_rows(self._v)


def make_optimizer(model, lr=1e-3, **kw):
    """FusedAdam over ``model.parameters()``, or the owner-sharded form when the model sits on a sharded graph with the
    fused exchange (and TAGREC_SHARDED_ADAM != 0)."""
    import os
    comm = getattr(getattr(model, "norm_adj", None), "comm", None)
    if (comm is not None and comm.peer is not None and comm.world > 1 and hasattr(model, "embed")
            and [id(p) for p in model.parameters()] == [id(p) for p in model.embed] and os.environ.get("TAGREC_SHARDED_ADAM", "1") != "0"):
        return ShardedFusedAdam(model, lr=lr, **kw)
    if (comm is None and hasattr(model, "embed") and hasattr(model, "_flat_params")
            and os.environ.get("TAGREC_ADAM_EPILOGUE") in ("1", "force")
            and [id(p) for p in model.parameters()] == [id(p) for p in model.embed] and not model._dropout_active()):
        # one GPU, on request only: the update in the epilogue of the last backward launch (no gradient table; p.grad stays
        # None).  Measured on the 1 B-edge graph: 241.6 vs 241.4 ms per step — the launch grows by what the separate pass
        # cost (both are bound by the same parameter / state traffic), so FusedAdam stays the default there; on several
        # GPUs the epilogue hides the NVLink ingest of the new parameters, which a separate pass cannot.
        return ShardedFusedAdam(model, lr=lr, **kw)
    return FusedAdam(model.parameters(), lr=lr, **kw)
