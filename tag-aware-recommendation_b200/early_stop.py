"""Patience-based early stopping with a best-checkpoint file.

Behavioural contract taken from training/early_stop.py:7-41 (what ``Basic_train`` and user scripts observe):
``Early_stop(args)(model, results, epoch) -> bool`` returns True once the monitored metric has failed to improve for
more than ``patient_epoch`` consecutive evaluations; every improvement stores ``model.state_dict()`` under
``<args.out_dir>/model.pth.tar`` and updates ``best_value / best_result / best_epoch / count_step``.  Ranking metrics
(precision, recall, ndcg) are maximised, anything else (a loss) is minimised.  The checkpoint is written to a
temporary name and renamed, so an interrupted run never leaves a truncated file behind.
"""
import os
from collections.abc import Iterable

import torch

from . import config

_MAXIMISED = ("precision", "recall", "ndcg")


class Early_stop:
    def __init__(self, args):
        cfg = config.current()
        self.key = cfg['early_stop_key']
        self.patient_step = cfg['patient_epoch']
        self.save_path = f"{args.out_dir}/model.pth.tar"
        self._sign = 1.0 if self.key in _MAXIMISED else -1.0
        self.best_value, self.best_result, self.best_epoch = None, None, 0
        self.count_step = 0

    def _monitored(self, results):
        v = results[self.key]
        return v[0] if isinstance(v, Iterable) else v          # first cut-off of CFG['topks']

    def _improved(self, value):
        return self.best_value is None or self._sign * (value - self.best_value) > 0

    def _checkpoint(self, model):
        tmp = self.save_path + ".tmp"
        torch.save(model.state_dict(), tmp)
        os.replace(tmp, self.save_path)

    def __call__(self, model, cur_results, epoch):
        value = self._monitored(cur_results)
        if not self._improved(value):
            self.count_step += 1
            return self.count_step > self.patient_step
        self.best_value, self.best_result, self.best_epoch = value, cur_results, epoch
        self.count_step = 0
        self._checkpoint(model)
        return False
