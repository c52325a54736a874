"""training/early_stop.py:7-41 — best-metric tracking + state_dict checkpoint (unchanged behaviour)."""
from collections.abc import Iterable

import torch

from . import config


class Early_stop:
    def __init__(self, args):
        cfg = config.current()
        self.best_value = None
        self.count_step = 0
        self.best_result = None
        self.best_epoch = 0
        self.patient_step = cfg['patient_epoch']
        self.save_path = f"{args.out_dir}/model.pth.tar"
        self.key = cfg['early_stop_key']
        if self.key in ['precision', 'recall', 'ndcg']:
            self.cmp = lambda x, y: x > y
        else:
            self.cmp = lambda x, y: x < y

    def __call__(self, model, cur_results, epoch):
        cur = cur_results[self.key]
        cur_res = cur[0] if isinstance(cur, Iterable) else cur
        if self.best_value is None or self.cmp(cur_res, self.best_value):
            self.best_value = cur_res
            self.count_step = 0
            torch.save(model.state_dict(), self.save_path)
            self.best_result = cur_results
            self.best_epoch = epoch
        else:
            self.count_step += 1
        return self.count_step > self.patient_step
