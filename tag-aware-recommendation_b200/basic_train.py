"""Train loop — drop-in for training/basic_train.py:10-85.

Same orchestration (reset, mini_batch, loss tuple, sum, zero_grad/backward/step, eval every test_interval, early
stop, returned ``loss_list``); the one change is that the three per-step host syncs of the reference
(``x.cpu().item()`` twice and ``loss.cpu().item()``, basic_train.py:16,27) are deferred to ONE device->host copy at
the end of the epoch — the returned numbers are the same.
"""
import time

import numpy as np
import torch

from . import config
from .early_stop import Early_stop


def epoch_training(training_data, loss_func, opt):
    training_data.reset()
    parts, totals = [], []
    for data in training_data.mini_batch():
        lossx = loss_func(data)
        parts.append(torch.stack([x.detach() for x in lossx]))
        loss = sum(lossx)
        if getattr(loss_func, "graphed", False):     # graph_step.GraphedStep: backward + optimizer ran in the graph
            totals.append(loss.detach())
            continue
        if isinstance(opt, list):
            [op.zero_grad() for op in opt]
            loss.backward()
            [op.step() for op in opt]
        else:
            opt.zero_grad()
            loss.backward()
            opt.step()
        totals.append(loss.detach())
    all_loss = torch.stack(parts).cpu().numpy() if parts else np.zeros((0, 2))
    loss_list = [float(x) for x in torch.stack(totals).cpu().numpy()] if totals else []
    print(f"[avg_loss of each part]:{list(all_loss.sum(0))}")
    return loss_list


def add_loss_to_writer(writer, values, i, ep):
    if writer:
        n = len(values)
        for j in range(n):
            writer.add_scalar(f'Train/loss_{i}', values[j], ep * n + j)


def add_result_to_writer(writer, data_dict, epoch, name):
    if writer:
        for key, val in data_dict.items():
            if len(val) > 1:
                writer.add_scalars(f'test/{key}', {f'@{name[i]}': val[i] for i in range(len(val))}, epoch)
            else:
                writer.add_scalar(f'test/{key}', val, epoch)


class Basic_train():
    def __init__(self, train_data: list, loss_func: list, opt: list, test, args=None):
        self.train_sphase = len(train_data)
        self.train_data = train_data
        self.loss_func = loss_func
        self.opt = opt
        self.test = test
        self.early_stop = Early_stop(args)
        self.args = args

    def run(self, model):
        cfg = config.current()
        for ep in range(cfg['epochs']):
            model.train()
            for i in range(self.train_sphase):
                start = time.time()
                loss_list = epoch_training(self.train_data[i], self.loss_func[i], self.opt[i])
                print(f"[Epoch:{ep}][Time:{(time.time()-start)/60:.2}]:"
                      f"avg_loss_{i} :{sum(loss_list)/len(loss_list):.5}")
                add_loss_to_writer(self.args.writer, loss_list, i, ep)
            if ep % cfg['test_interval'] == 0:
                start = time.time()
                results = self.test.run(model)
                print(f"[Epoch {ep}][Time:{(time.time()-start)/60:.2}] results: {results}")
                add_result_to_writer(self.args.writer, results, ep, cfg['topks'])
                if self.early_stop(model, results, ep):
                    print(f"early stop trigger at epoch {ep}")
                    break
        print(f"best result [{self.early_stop.best_epoch}:{self.early_stop.best_result}]")
        if self.args.writer:
            self.args.writer.add_text("LOG", f"best results: epoch-{self.early_stop.best_epoch}:{self.early_stop.best_result}")
