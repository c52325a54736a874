"""CUDA-graph capture of one training step.

On the small graphs (C1/C2: tables of a few MB, everything L2-resident) a LightGCN step is ~30 kernel launches of a
few microseconds each: it is LAUNCH-bound, not bandwidth-bound.  ``GraphedStep`` records
``lossx = model.loss(batch); sum(lossx).backward(); opt.step()`` once and replays it as one graph launch.

    step = T.GraphedStep(model, opt)            # opt: T.FusedAdam(..., capturable=True) or torch Adam(capturable=True)
    train = T.Basic_train([sampler], [step.loss], [step.opt], test, args)     # epoch_training recognises the pair

The first ``warmup`` calls run eagerly (real steps: nothing is executed twice or with fake data), the next call is
captured (capture records, it does not execute) and replayed.  Batches whose shape differs from the captured one —
the merged tail of an epoch (train_data/abstract.py:17-23) — run eagerly.  Losses are returned as tensors, like
``model.loss`` does; reading them is the caller's only synchronisation.
"""
import torch


class _NoopOpt:
    """What ``epoch_training`` gets as optimizer for a graphed step: zero_grad / step already happened in the graph."""

    def __init__(self, real):
        self.real = real
        self.param_groups = real.param_groups

    def zero_grad(self, set_to_none=True):
        pass

    def step(self):
        pass

    def state_dict(self):
        return self.real.state_dict()


class GraphedStep:
    def __init__(self, model, opt, warmup=3):
        for g in opt.param_groups:
            if not g.get("capturable", False):
                raise ValueError("GraphedStep needs a capturable optimizer (T.FusedAdam(..., capturable=True) or "
                                 "torch.optim.Adam(..., capturable=True)): the step counter must live on the device")
        self.model, self.real_opt, self.warmup = model, opt, warmup
        self.opt = _NoopOpt(opt)
        self.calls = 0
        self.graph = None
        self.static_batch = None
        self.static_out = None
        self.loss.__func__.graphed = True

    def _eager(self, batch):
        lossx = self.model.loss(batch)
        self.real_opt.zero_grad(set_to_none=True)
        sum(lossx).backward()
        self.real_opt.step()
        return tuple(x.detach() for x in lossx)

    # a batch is a LongTensor, or the (data, cor) tuple of DGCF_training_data (cor may be None)
    @staticmethod
    def _tensors(batch):
        return [t for t in (batch if isinstance(batch, (tuple, list)) else (batch,)) if torch.is_tensor(t)]

    @staticmethod
    def _clone(batch):
        if isinstance(batch, (tuple, list)):
            return tuple(t.clone() if torch.is_tensor(t) else t for t in batch)
        return batch.clone()

    def _capture(self, batch):
        self.static_batch = self._clone(batch)
        self.real_opt.zero_grad(set_to_none=True)          # gradients are (re)created inside the graph's memory pool
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            lossx = self.model.loss(self.static_batch)
            sum(lossx).backward()
            self.real_opt.step()
            self.static_out = tuple(x.detach() for x in lossx)

    def loss(self, batch):
        """Drop-in for ``model.loss`` in ``Basic_train(..., loss_func=[step.loss], opt=[step.opt], ...)``: runs the WHOLE
        step and returns the loss tuple (already detached)."""
        self.calls += 1
        new = self._tensors(batch)
        if self.calls <= self.warmup or not all(t.is_cuda for t in new):
            return self._eager(batch)
        if self.graph is None:
            torch.cuda.synchronize()
            self._capture(batch)
        old = self._tensors(self.static_batch)
        if len(old) != len(new) or any(a.shape != b.shape for a, b in zip(old, new)):
            return self._eager(batch)
        for a, b in zip(old, new):
            a.copy_(b)
        self.graph.replay()
        # the replay rewrote the parameters without going through torch: drop the model's cached inference table and
        # bump the versions that version-keyed caches look at
        if hasattr(self.model, "_cache"):
            self.model._cache = None
        for g in self.real_opt.param_groups:
            for p in g["params"]:
                torch._C._increment_version([p])
        return tuple(x.clone() for x in self.static_out)
