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
