"""Fused Adam optimizer over the C ABI (tagrec_adam_step): one pass over param/grad/exp_avg/exp_avg_sq per tensor.

Optional replacement for the ``optim.Adam(model.parameters(), lr=CFG['lr'])`` the reference composes in com.py:25
(SURVEY §8 f-3).  Same update rule and state names as torch.optim.Adam (amsgrad=False), so ``state_dict()`` of the
optimizer stays loadable by torch's Adam; checked against it in tests/test_gpu_parity.py.
"""
import torch

from ._lib import check, lib, ptr, stream_ptr


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        L = lib()
        for group in self.param_groups:
            b1, b2 = group["betas"]
            for p in group["params"]:
                if p.grad is None:
                    continue
                if not (p.is_cuda and p.dtype == torch.float32 and p.is_contiguous()):
                    raise RuntimeError("FusedAdam needs contiguous float32 CUDA parameters (no CPU fallback)")
                st = self.state[p]
                if not st:
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros_like(p)
                    st["exp_avg_sq"] = torch.zeros_like(p)
                st["step"] += 1
                g = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
                check(L.tagrec_adam_step(ptr(p), ptr(g), ptr(st["exp_avg"]), ptr(st["exp_avg_sq"]), p.numel(),
                                         group["lr"], b1, b2, group["eps"], group["weight_decay"], st["step"],
                                         stream_ptr(p.device)), "tagrec_adam_step")
        return loss
