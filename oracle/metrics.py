"""Oracle: full-sort evaluation — masking, top-K by (-score, item id), recall/precision/hr/ndcg/auc.

Restates training/basic_test.py:12-111 and training/utils.py:7-54 in vectorised numpy (float64 sums).
The reference's own ``torch.topk`` has arbitrary tie order (SURVEY §4 iii), so the canonical order here is the
stable (-score, id) order; metrics are identical whenever the K-boundary has no exact tie.
TEST INFRASTRUCTURE — see oracle/__init__.py.
"""
import numpy as np

MASK_VALUE = -float(1 << 10)        # basic_test.py:47


def mask_train(scores, users, train_ptr, train_items):
    """basic_test.py:42-47: rating[row, train_items(u)] = -1024 (users without train items: no mask, :37)."""
    s = np.array(scores, copy=True)
    for r, u in enumerate(users):
        s[r, train_items[train_ptr[u]:train_ptr[u + 1]]] = MASK_VALUE
    return s


def topk_ids(scores, k):
    """Top-k item ids per row ordered by (-score, id)."""
    n = scores.shape[1]
    ids = np.broadcast_to(np.arange(n), scores.shape)
    order = np.lexsort((ids, -scores.astype(np.float64)), axis=1)
    return order[:, :k]


def minibatch_slices(n, batch):
    """training/utils.py:48-54 — note the trailing EMPTY batch when n % batch == 0 (crashes the reference,
    SURVEY A13); returned here so callers can decide."""
    step = n // batch + 1
    return [(i * batch, n if (i + 1) * batch > n else (i + 1) * batch) for i in range(step)]


def ranking_metrics(topk, users, test_ptr, test_items, ks):
    """training/utils.py:7-35 get_label / pre_rec_k / ndcg_k summed over users (NOT yet divided)."""
    nu, kmax = topk.shape
    label = np.zeros((nu, kmax), dtype=np.float64)
    n_true = np.zeros(nu, dtype=np.float64)
    for r, u in enumerate(users):
        t = test_items[test_ptr[u]:test_ptr[u + 1]]
        n_true[r] = len(t)
        label[r] = np.isin(topk[r], t)
    out = {"recall": [], "precision": [], "hr": [], "ndcg": []}
    for k in ks:
        right = label[:, :k].sum(1)
        out["precision"].append(right.sum() / k)
        out["recall"].append((right / n_true).sum())
        out["hr"].append(float((right > 0).sum()))
        disc = 1.0 / np.log2(np.arange(2, k + 2))
        ideal_len = np.minimum(k, n_true).astype(np.int64)
        idcg = np.array([disc[:m].sum() for m in ideal_len])
        idcg[idcg == 0.0] = 1.0
        dcg = (label[:, :k] * disc).sum(1)
        out["ndcg"].append((dcg / idcg).sum())
    return out


def auc_one(masked_row, test_items_u):
    """training/utils.py:37-45: roc_auc_score over the un-masked items (score >= 0) == Mann-Whitney U with
    average ranks for ties."""
    keep = masked_row >= 0
    y = np.zeros(len(masked_row), dtype=bool)
    y[test_items_u] = True
    y, s = y[keep], masked_row[keep].astype(np.float64)
    order = np.argsort(s, kind="stable")
    ss = s[order]
    ranks = np.empty(len(s), dtype=np.float64)
    # average ranks of tie groups
    bounds = np.flatnonzero(np.r_[True, ss[1:] != ss[:-1], True])
    for a, b in zip(bounds[:-1], bounds[1:]):
        ranks[order[a:b]] = 0.5 * (a + 1 + b)
    npos = y.sum()
    nneg = len(y) - npos
    return (ranks[y].sum() - npos * (npos + 1) / 2.0) / (npos * nneg)


def epoch_test(masked_scores, users, test_ptr, test_items, ks, with_auc=True):
    """basic_test.py:30-80 epoch_test given the masked score rows of ``users`` (dict-key order)."""
    kmax = max(ks)
    top = topk_ids(masked_scores, kmax)
    sums = ranking_metrics(top, users, test_ptr, test_items, ks)
    n = float(len(users))
    res = {k: [x / n for x in v] for k, v in sums.items()}
    if with_auc:
        tot = sum(auc_one(masked_scores[r], test_items[test_ptr[u]:test_ptr[u + 1]]) for r, u in enumerate(users))
        res["auc"] = [tot / n]
    return res, top


# ---------------------------------------------------------------------------------------------------------------
# bench.py's CPU arm: epoch_test AS THE REFERENCE EXECUTES IT (training/basic_test.py:30-80): per user batch a dense
# sigmoid(U I^T) score matrix, Python lists for the mask, torch.topk, and — the expensive part — a full score row
# copied out and sklearn.metrics.roc_auc_score per user (training/utils.py:37-45).
def reference_epoch_test(user_table, item_table, pos_ui, true_ui, topks, test_batch, with_auc=True, max_users=None):
    """user_table / item_table: torch CPU tensors = model.forward()[:2].  Returns the reference's result dict.
    ``max_users`` bounds the number of evaluated users (bench.py's bounded sample; metrics are then over that subset)."""
    import torch
    from sklearn.metrics import roc_auc_score
    all_users = list(true_ui.keys())
    if max_users is not None:
        all_users = all_users[:max_users]
    max_k = max(topks)
    n_item = item_table.shape[0]
    tot = {k: np.zeros(len(topks)) for k in ("precision", "recall", "hr", "ndcg")}
    auc = []
    with torch.no_grad():
        for s in range(0, len(all_users), test_batch):            # training/utils.py:48-54 (without the empty batch)
            user = all_users[s:s + test_batch]
            allpos = [pos_ui[u] if u in pos_ui else [] for u in user]
            truth = [true_ui[u] for u in user]
            rating = torch.sigmoid(torch.matmul(user_table[torch.tensor(user, dtype=torch.long)], item_table.t()))
            rows, cols = [], []
            for i, item in enumerate(allpos):
                rows.extend([i] * len(item))
                cols.extend(item)
            rating[rows, cols] = -(1 << 10)
            _, top = torch.topk(rating, k=max_k)
            top = top.cpu().numpy()
            if with_auc:
                for i, t in enumerate(truth):                     # training/utils.py:37-45
                    all_item_scores = rating[i].detach().cpu().numpy()
                    r_all = np.zeros((n_item,))
                    r_all[t] = 1
                    keep = all_item_scores >= 0
                    auc.append(roc_auc_score(r_all[keep], all_item_scores[keep]))
            label = np.array([[float(x in set(t)) for x in row] for row, t in zip(top, truth)])   # training/utils.py:7-13
            for q, k in enumerate(topks):                         # training/utils.py:15-35
                right = label[:, :k].sum(1)
                n_true = np.array([len(t) for t in truth], dtype=np.float64)
                tot["precision"][q] += right.sum() / k
                tot["recall"][q] += (right / n_true).sum()
                tot["hr"][q] += (right > 0).sum()
                disc = 1.0 / np.log2(np.arange(2, k + 2))
                idcg = np.array([disc[:int(min(k, m))].sum() for m in n_true])
                idcg[idcg == 0.0] = 1.0
                tot["ndcg"][q] += ((label[:, :k] * disc).sum(1) / idcg).sum()
    n = float(len(all_users))
    res = {k: list(v / n) for k, v in tot.items()}
    if with_auc:
        res["auc"] = [float(np.sum(auc)) / n]
    return res
