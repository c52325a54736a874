"""Train loop with the call contract of training/basic_train.py:10-85.

``Basic_train(train_data, loss_func, opt, test, args).run(model)`` drives, per epoch and per training phase,
``reset() -> mini_batch() -> loss_func(batch) -> sum -> zero_grad / backward / step``, evaluates every
``test_interval`` epochs through ``test.run(model)``, feeds the result to ``Early_stop`` and logs to ``args.writer``
(tensorboardX-style, may be None).  ``epoch_training`` returns the per-step total losses as python floats, like the
reference.

What is different on purpose: the reference reads three scalars back to the host on EVERY step
(``x.cpu().item()`` per loss part and ``loss.cpu().item()``, basic_train.py:16,27), which serialises the host with the
device.  Here every step's losses stay on the device in one pre-sized buffer and are copied back ONCE at the end of
the epoch; the numbers returned and printed are the same.  A ``graph_step.GraphedStep`` passed as ``loss_func`` has
already run backward + optimizer inside its CUDA graph, so the loop only collects its losses.
"""
import time

import torch

from . import config
from .early_stop import Early_stop


class _LossLog:
    """Device-side log of (parts..., total) per step; one device->host copy per epoch."""

    def __init__(self):
        self.rows = []

    def add(self, lossx, total):
        self.rows.append(torch.stack([x.detach().reshape(()) for x in lossx] + [total.detach().reshape(())]))

    def fetch(self):
        if not self.rows:
            return [], []
        host = torch.stack(self.rows).to("cpu", torch.float64).numpy()
        return [float(v) for v in host[:, -1]], [float(v) for v in host[:, :-1].sum(0)]


def _optimizers(opt):
    return opt if isinstance(opt, (list, tuple)) else (opt,)


def epoch_training(training_data, loss_func, opt):
    """One pass over ``training_data`` (a fresh negative sample per epoch, basic_train.py:12)."""
    log = _LossLog()
    graphed = getattr(loss_func, "graphed", False)
    training_data.reset()
    for batch in training_data.mini_batch():
        lossx = loss_func(batch)
        total = sum(lossx)
        if not graphed:
            for o in _optimizers(opt):
                o.zero_grad()
            total.backward()
            for o in _optimizers(opt):
                o.step()
        log.add(lossx, total)
    loss_list, part_sums = log.fetch()
    print(f"[avg_loss of each part]:{part_sums}")
    return loss_list


def add_loss_to_writer(writer, values, i, ep):
    if not writer:
        return
    base = ep * len(values)
    for j, v in enumerate(values):
        writer.add_scalar(f'Train/loss_{i}', v, base + j)


def add_result_to_writer(writer, data_dict, epoch, name):
    if not writer:
        return
    for key, val in data_dict.items():
        if len(val) == 1:
            writer.add_scalar(f'test/{key}', val, epoch)
        else:
            writer.add_scalars(f'test/{key}', {f'@{k}': v for k, v in zip(name, val)}, epoch)


class Basic_train():
    def __init__(self, train_data: list, loss_func: list, opt: list, test, args=None):
        self.train_data, self.loss_func, self.opt = train_data, loss_func, opt
        self.train_sphase = len(train_data)          # (sic) attribute name kept for scripts that read it
        self.test = test
        self.args = args
        self.early_stop = Early_stop(args)

    def _writer(self):
        return getattr(self.args, "writer", None)

    def _train_epoch(self, ep):
        for i, (td, lf, op) in enumerate(zip(self.train_data, self.loss_func, self.opt)):
            t0 = time.time()
            losses = epoch_training(td, lf, op)
            mean = sum(losses) / max(1, len(losses))
            print(f"[Epoch:{ep}][Time:{(time.time() - t0) / 60:.2}]:avg_loss_{i} :{mean:.5}")
            add_loss_to_writer(self._writer(), losses, i, ep)

    def _evaluate(self, model, ep):
        t0 = time.time()
        results = self.test.run(model)
        print(f"[Epoch {ep}][Time:{(time.time() - t0) / 60:.2}] results: {results}")
        add_result_to_writer(self._writer(), results, ep, config.current()['topks'])
        return self.early_stop(model, results, ep)

    def run(self, model):
        cfg = config.current()
        for ep in range(cfg['epochs']):
            model.train()
            self._train_epoch(ep)
            if ep % cfg['test_interval'] == 0 and self._evaluate(model, ep):
                print(f"early stop trigger at epoch {ep}")
                break
        es = self.early_stop
        print(f"best result [{es.best_epoch}:{es.best_result}]")
        if self._writer():
            self._writer().add_text("LOG", f"best results: epoch-{es.best_epoch}:{es.best_result}")
