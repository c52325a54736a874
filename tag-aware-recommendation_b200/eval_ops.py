"""K3 wrappers: fused scoring + masking + top-K, and the metric sums."""
import torch

from ._lib import check, lib, ptr, stream_ptr


PATHS = {"auto": 0, "fp32": 1, "tf32": 2}     # TAGREC_EVAL_*


def topk_scores(users, user_table, item_table, train_ptr, train_items, k, path="auto"):
    """users int64 [nu] (rows of user_table); returns (ids int32 [nu,k], scores fp32 [nu,k]) ordered by
    (-score, item id); train items of each user rank below everything else with score -1024.
    ``path``: "auto" (tcgen05 TF32 filter + exact fp32 re-score when dim == 64, else the fp32 CUDA-core tiles),
    "fp32" or "tf32" — all return the same lists."""
    L = lib()
    dev = user_table.device
    users = users.to(device=dev, dtype=torch.int64).contiguous()
    ut, it = user_table.contiguous(), item_table.contiguous()
    nu, n_item, dim = users.numel(), it.shape[0], it.shape[1]
    ids = torch.empty((nu, k), dtype=torch.int32, device=dev)
    scores = torch.empty((nu, k), dtype=torch.float32, device=dev)
    ws = torch.empty(int(L.tagrec_eval_workspace_bytes(nu, n_item, k)), dtype=torch.uint8, device=dev)
    check(L.tagrec_eval_topk_ex(ptr(users), nu, ptr(ut), ptr(it), n_item, dim, ptr(train_ptr), ptr(train_items), k,
                                ptr(ids), ptr(scores), ptr(ws), ws.numel(), PATHS[path], stream_ptr(dev)),
          "tagrec_eval_topk")
    return ids, scores


def eval_plan(nu, n_item, dim, k):
    """Launch shape of the tensor-core top-K path for this problem size (host only): dict with tensor_core, cta_pairs
    (eval_tc2_kernel, tcgen05.mma.cta_group::2), halves, splits, stages, lists."""
    import ctypes as C
    plan = (C.c_int32 * 6)()
    check(lib().tagrec_eval_plan(int(nu), int(n_item), int(dim), int(k), plan), "tagrec_eval_plan")
    keys = ("tensor_core", "cta_pairs", "halves", "splits", "stages", "lists")
    return {k_: int(v) for k_, v in zip(keys, plan)}


def metric_sums(users, topk_ids, test_ptr, test_items, ks, out=None):
    """training/utils.py:15-35 summed over ``users``: returns float64 [4, len(ks)] = recall|precision|hr|ndcg."""
    dev = topk_ids.device
    ks_t = torch.as_tensor(list(ks), dtype=torch.int32, device=dev)
    if out is None:
        out = torch.zeros((4, len(ks)), dtype=torch.float64, device=dev)
    users = users.to(device=dev, dtype=torch.int64).contiguous()
    check(lib().tagrec_eval_metrics(ptr(users), users.numel(), ptr(topk_ids), topk_ids.shape[1], ptr(test_ptr),
                                    ptr(test_items), ptr(ks_t), len(ks), ptr(out), stream_ptr(dev)),
          "tagrec_eval_metrics")
    return out


def auc_sums(users, user_table, item_table, train_ptr, train_items, test_ptr, test_items, out=None, path="auto"):
    """training/utils.py:37-45 summed over ``users``: float64 [2] = (sum of per-user AUC, users with both classes).
    ``path``: "auto" (tensor cores for 64-d tables), "fp32", "tf32" — identical sums."""
    L = lib()
    dev = user_table.device
    users = users.to(device=dev, dtype=torch.int64).contiguous()
    ut, it = user_table.contiguous(), item_table.contiguous()
    if out is None:
        out = torch.zeros(2, dtype=torch.float64, device=dev)
    n_test = int(test_items.numel())
    ws = torch.empty(int(L.tagrec_eval_auc_workspace_bytes(users.numel(), n_test)), dtype=torch.uint8, device=dev)
    check(L.tagrec_eval_auc_ex(ptr(users), users.numel(), ptr(ut), ptr(it), it.shape[0], it.shape[1], ptr(train_ptr),
                               ptr(train_items), ptr(test_ptr), ptr(test_items), n_test, ptr(ws), ws.numel(), ptr(out),
                               {"auto": 0, "fp32": 1, "tf32": 2}[path], stream_ptr(dev)), "tagrec_eval_auc_ex")
    return out


class EvalMixin:
    """K3 entry points shared by the drop-in models: everything the evaluation loop needs from ``forward()``'s
    (user, item) tables without materialising predict_rating's [B, n_item] matrix."""

    def train(self, mode=True):
        """The inference tables are cached between user batches of ONE evaluation run (the reference re-propagates
        for every batch, lightgcn.py:85).  Every mode switch drops the cache: ``Basic_train.run`` calls ``train()`` at
        each epoch and both evaluation loops call ``eval()`` first, so a table never survives a parameter update even
        when the optimizer writes through raw pointers (FusedAdam, CUDA-graph replays) and leaves ``_version`` alone."""
        self._cache = None
        return super().train(mode)

    def _eval_tables(self):
        with torch.no_grad():
            all_users, all_items = self.forward()[:2]
        pad = (-all_users.shape[1]) % 32
        if pad:      # K3 tiles the embedding dimension in 32s (176-d tables of the default [64, 32, 16] widths): zero
            # columns leave every dot product unchanged
            all_users = torch.nn.functional.pad(all_users, (0, pad))
            all_items = torch.nn.functional.pad(all_items, (0, pad))
        return all_users.contiguous(), all_items.contiguous()

    def eval_topk(self, users, k, train_ptr, train_items, path="auto"):
        """Top-k item ids / scores per user with the user's train items masked (basic_test.py:40-48)."""
        all_users, all_items = self._eval_tables()
        return topk_scores(users, all_users, all_items, train_ptr, train_items, k, path=path)

    def eval_auc(self, users, train_ptr, train_items, test_ptr, test_items, out=None):
        """Sum of the per-user AUC (training/utils.py:37-45) and the number of users it is defined for."""
        all_users, all_items = self._eval_tables()
        return auc_sums(users, all_users, all_items, train_ptr, train_items, test_ptr, test_items, out=out)
