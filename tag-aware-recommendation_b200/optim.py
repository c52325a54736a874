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
