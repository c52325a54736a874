"""Full-sort evaluation loop — drop-in for training/basic_test.py:12-111 (+ training/utils.py).

``Basic_test(data, args).run(model, istest=False, group_k=0)`` returns the same dict
``{'recall': [per k], 'precision': [...], 'hr': [...], 'ndcg': [...], 'auc': [x]}`` (or a dict of such dicts keyed
``inter<{n}-{count}`` when ``group_k > 1``).  Per user chunk the model's ``eval_topk`` (K3: tcgen05 scoring + mask +
top-K) produces the masked top-K directly, ``eval_auc`` (K3b) the per-user AUC sums; everything stays on the device
until the end.  A model must offer ``eval_topk`` / ``eval_auc`` (``eval_ops.EvalMixin``: any model whose ``forward()``
returns the (user, item) tables that ``predict_rating`` multiplies); there is no host path.
Multi-GPU (new): users are sharded over the ranks of the default process group, sums are all-reduced.

Documented deviations: (i) ties are ordered by item id (the reference's torch.topk order is arbitrary);
(ii) no crash when ``len(users) % test_batch == 0`` (the reference yields an empty batch and raises IndexError,
training/utils.py:48-54).
"""
from collections import defaultdict

import numpy as np
import torch

from . import config
from .bpr_training_data import user_items_to_csr
from .eval_ops import metric_sums


def minibatch(data, batch_size):
    """training/utils.py:48-54 without the trailing empty batch."""
    for i in range(0, len(data), batch_size):
        yield data[i:i + batch_size]


def user_group_split(test_ui, train_ui, k, method="interaction"):
    """training/utils.py:62-109 (method 'interaction': equal shares of the total interaction count)."""
    num_inter = defaultdict(list)
    tot_inter = 0
    for u in test_ui.keys():
        n_inter = len(test_ui[u]) + (len(train_ui[u]) if u in train_ui else 0)
        num_inter[n_inter].append(u)
        tot_inter += n_inter
    step = tot_inter // k
    end = list(range(step, tot_inter + 1, step))
    end[-1] = tot_inter
    groups, count, i, temp = dict(), 0, 0, []
    for n in sorted(num_inter):
        temp += num_inter[n]
        count += n * len(num_inter[n])
        if count >= end[i]:
            groups[n] = temp
            temp = []
            i += 1
            print(f"interaction < {n} has {len(groups[n])} user")
    return groups


class Basic_test():
    def __init__(self, data, args=None):
        cfg = config.current()
        self.args = args
        self.pos_ui = data.user_items['train']
        self.true_ui = dict()
        if cfg['has_val'] == True:
            self.true_ui['val'] = data.user_items['val']
        self.true_ui['test'] = data.user_items['test']
        self.num_user = int(data.num['user'])
        self._dev_csr = {}
        print("Basic_test got ready!")

    def _csr(self, name, dic, device):
        key = (name, str(device))
        if key not in self._dev_csr:
            p, items = user_items_to_csr(dic, self.num_user)
            self._dev_csr[key] = (torch.as_tensor(p, device=device), torch.as_tensor(items, device=device).to(torch.int32))
        return self._dev_csr[key]

    @staticmethod
    def _world():
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            return dist.get_rank(), dist.get_world_size()
        return 0, 1

    def epoch_test(self, model, true_name, true_ui, all_users=None):
        """basic_test.py:30-80.  With torch.distributed initialised (one process per GPU) the users are sharded
        round-robin over the ranks and the metric sums are all-reduced — every rank returns the full result."""
        cfg = config.current()
        if all_users is None:
            all_users = list(true_ui.keys())
        topks = list(cfg['topks'])
        n = len(all_users)
        rank, world = self._world()
        mine = all_users[rank::world] if (world > 1 and cfg.get('eval_shard', True)) else all_users
        sharded = world > 1 and cfg.get('eval_shard', True)
        if not hasattr(model, "eval_topk"):
            raise TypeError(f"{type(model).__name__} offers no eval_topk(): this evaluation loop scores on the device "
                            "(K3) and has no host path.  Give the model tagrec_b200.eval_ops.EvalMixin — it only needs "
                            "forward() to return the (user, item) tables predict_rating multiplies")
        sums, auc = self._sums_device(model, true_name, true_ui, mine, topks)
        if sharded:
            import torch.distributed as dist
            packed = torch.cat([sums.flatten(), auc.flatten()])
            dist.all_reduce(packed)
            sums, auc = packed[:sums.numel()].reshape(sums.shape), packed[sums.numel():]
        sums = (sums / n).cpu().numpy()
        ret = {'recall': list(sums[0]), 'precision': [np.float32(x) for x in sums[1]], 'hr': list(sums[2]),
               'ndcg': list(sums[3])}
        if auc.numel():
            ret['auc'] = [float(auc[0].item()) / n]
        return ret

    def _sums_device(self, model, true_name, true_ui, users, topks):
        """K3 path: masked top-K (tcgen05 / fp32 kernels), metric sums and AUC sums stay on the device."""
        cfg = config.current()
        device = cfg['device']
        max_k = max(topks)
        train_ptr, train_items = self._csr("train", self.pos_ui, device)
        test_ptr, test_items = self._csr(true_name, true_ui, device)
        sums = torch.zeros((4, len(topks)), dtype=torch.float64, device=device)
        want_auc = cfg.get('eval_auc', True) and hasattr(model, "eval_auc")
        auc = torch.zeros(2 if want_auc else 0, dtype=torch.float64, device=device)
        users_t = torch.as_tensor(np.asarray(users, dtype=np.int64), device=device)
        # no [B, n_item] matrix is ever materialised, so the user chunk is not bounded by test_batch (a memory knob of
        # the reference): larger chunks keep all SMs on one wave of item splits
        chunk = max(int(cfg['test_batch']), int(cfg.get('eval_chunk', 16384)))
        for s in range(0, len(users), chunk):
            ub = users_t[s:s + chunk]
            ids, _ = model.eval_topk(ub, max_k, train_ptr, train_items)
            metric_sums(ub, ids, test_ptr, test_items, topks, out=sums)
            if want_auc:
                model.eval_auc(ub, train_ptr, train_items, test_ptr, test_items, out=auc)
        return sums, auc

    def run(self, model, istest=False, group_k=0):
        cfg = config.current()
        model.eval()
        if istest == False and cfg['has_val']:
            name = 'val'
        else:
            name = 'test'
        true_ui = self.true_ui[name]
        if group_k > 1:
            all_result = dict()
            for key, all_user in user_group_split(true_ui, self.pos_ui, group_k).items():
                all_result[f"inter<{key}-{len(all_user)}"] = self.epoch_test(model, name, true_ui, all_user)
            return all_result
        return self.epoch_test(model, name, true_ui)
