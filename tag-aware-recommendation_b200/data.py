"""Data containers the hot path consumes + deterministic synthetic graphs of the BASELINE shapes.

The reference's loaders (data/cf_load.py, data/tgcn_load.py — out of scope, host I/O) hand the models an object with
``num`` (dict user/item/tag/weight), ``ui_adj`` / ``ut_adj`` / ``it_adj`` (scipy COO float32), ``user_items``
(dict split -> dict u -> list) and ``edge_index`` (dict split -> E x 2 ndarray).  :class:`Dataset` is that contract;
``load_text`` reads the reference's file formats (train.txt / test.txt: ``u i1 i2 ...``; user_item_tag.txt:
``u i t``; data/utils.py:9-46) and the ``synth_*`` functions generate graphs of the named shapes (SURVEY §8 d):
user activity ~ lognormal, item popularity ~ Zipf(s), duplicates removed per user, 80/20 split per user.
"""
import os

import numpy as np
import scipy.sparse as sp
import torch


class Dataset:
    def __init__(self):
        self.num, self.user_items, self.edge_index = {}, {}, {}
        self.ui_adj = self.ut_adj = self.it_adj = None
        self.uit_data = None

    def create_edge(self):
        return create_edge(self)


def create_edge(ds):
    """data/tgcn_load.py:55-70 TGCN_load.create_edge: the six directed relations (ui iu ut tu it ti) as [2, E] arrays of
    global node ids (users, then items, then tags) — the layout KGAT_training_data expects."""
    U, I = int(ds.num["user"]), int(ds.num["item"])
    edge = {}
    user, item = ds.ui_adj.row, ds.ui_adj.col + U
    edge[0], edge[1] = np.stack([user, item]), np.stack([item, user])
    user, tag = ds.ut_adj.row, ds.ut_adj.col + I + U
    edge[2], edge[3] = np.stack([user, tag]), np.stack([tag, user])
    item, tag = ds.it_adj.row + U, ds.it_adj.col + I + U
    edge[4], edge[5] = np.stack([item, tag]), np.stack([tag, item])
    return edge


def _coo(rows, cols, shape):
    # data/utils.py:50-53 to_sparse_adj
    return sp.coo_matrix((np.ones_like(rows), (rows, cols)), dtype=np.float32, shape=shape)


def dict_edges(d):
    """data/utils.py:121-129 dict_info: (E,2) in dict / list order."""
    u = np.concatenate([np.full(len(v), k, dtype=np.int64) for k, v in d.items()]) if d else np.zeros(0, np.int64)
    i = np.concatenate([np.asarray(v, dtype=np.int64) for v in d.values()]) if d else np.zeros(0, np.int64)
    return np.stack([u, i], 1)


def finish(ds, n_user, n_item, uit=None, n_tag=0):
    ds.edge_index = {k: dict_edges(v) for k, v in ds.user_items.items()}
    ds.num = {"user": int(n_user), "item": int(n_item)}
    e = ds.edge_index["train"]
    ds.ui_adj = _coo(e[:, 0], e[:, 1], (n_user, n_item))
    if uit is not None:
        ds.uit_data = np.unique(np.asarray(uit, dtype=np.int32), axis=0)           # data/utils.py:11-13
        ds.num["tag"] = int(n_tag)
        ds.ut_adj = _coo(ds.uit_data[:, 0], ds.uit_data[:, 2], (n_user, n_tag))
        ds.it_adj = _coo(ds.uit_data[:, 1], ds.uit_data[:, 2], (n_item, n_tag))
        ds.num["weight"] = int(max(ds.ui_adj.max(), ds.ut_adj.tocsr().max(), ds.it_adj.tocsr().max()))
    return ds


def load_text(file_dir, has_val=False):
    """data/cf_load.py:8-28 + data/tgcn_load.py:11-25 on the reference's own file formats."""
    def read(name):
        out = {}
        with open(os.path.join(file_dir, name)) as f:
            for line in f:
                x = [int(t) for t in line.strip().split(' ') if t]
                if len(x) > 1:
                    out[x[0]] = list(set(out.get(x[0], []) + x[1:]))
        return out
    ds = Dataset()
    ds.user_items["train"] = read("train.txt")
    if has_val:
        ds.user_items["val"] = read("val.txt")
    ds.user_items["test"] = read("test.txt")
    mx_u = max(max(d) for d in ds.user_items.values())
    mx_i = max(max(max(v) for v in d.values()) for d in ds.user_items.values())
    uit_path = os.path.join(file_dir, "user_item_tag.txt")
    uit = np.loadtxt(uit_path, dtype=np.int32).reshape(-1, 3) if os.path.exists(uit_path) else None
    return finish(ds, mx_u + 1, mx_i + 1, uit, int(uit[:, 2].max()) + 1 if uit is not None else 0)


def write_text(ds, file_dir):
    os.makedirs(file_dir, exist_ok=True)
    for part, d in ds.user_items.items():
        with open(os.path.join(file_dir, f"{part}.txt"), "w") as f:
            for u, its in d.items():
                f.write(" ".join(str(x) for x in [u] + list(its)) + "\n")
    if ds.uit_data is not None:
        np.savetxt(os.path.join(file_dir, "user_item_tag.txt"), ds.uit_data, fmt="%d")


def synth_bipartite(n_user, n_item, n_edge, seed=2020, zipf=0.9, sigma=1.0, test_frac=0.2, n_tag=0, tags_per_edge=1.5):
    """Host (numpy) generator for the small/medium shapes (C1-C4).  ``n_edge`` counts train+test interactions."""
    rng = np.random.RandomState(seed)
    act = rng.lognormal(0.0, sigma, n_user)
    pop = 1.0 / np.arange(1, n_item + 1) ** zipf
    cdf = np.cumsum(pop / pop.sum())
    perm = rng.permutation(n_item)                    # popularity is not correlated with the item id
    want = float(n_edge)
    for _ in range(6):                                # duplicates are dropped; re-draw until the count is close
        deg = np.maximum(1, np.round(act / act.sum() * want)).astype(np.int64)
        deg = np.minimum(deg, n_item // 2)
        users = np.repeat(np.arange(n_user), deg)
        items = perm[np.minimum(np.searchsorted(cdf, rng.rand(len(users))), n_item - 1)]
        key = np.unique(users.astype(np.int64) * n_item + items)
        if abs(len(key) - n_edge) <= 0.01 * n_edge:
            break
        want *= n_edge / len(key)
    users, items = key // n_item, key % n_item
    is_test = rng.rand(len(users)) < test_frac
    # every user keeps at least one train item
    first = np.r_[True, users[1:] != users[:-1]]
    is_test[first] = False
    ds = Dataset()
    order = rng.permutation(len(users))               # list order inside a user is arbitrary in the files
    users, items, is_test = users[order], items[order], is_test[order]
    srt = np.argsort(users, kind="stable")
    users, items, is_test = users[srt], items[srt], is_test[srt]
    for part, m in (("train", ~is_test), ("test", is_test)):
        u, i = users[m], items[m]
        cut = np.flatnonzero(np.r_[True, u[1:] != u[:-1]])
        ds.user_items[part] = {int(u[a]): i[a:b].tolist() for a, b in zip(cut, np.r_[cut[1:], len(u)])}
    uit = None
    if n_tag:
        tu, ti = users[~is_test], items[~is_test]
        k = rng.poisson(tags_per_edge, len(tu)).clip(0, 4)
        ru, ri = np.repeat(tu, k), np.repeat(ti, k)
        tp = 1.0 / np.arange(1, n_tag + 1) ** 0.8
        tags = np.minimum(np.searchsorted(np.cumsum(tp / tp.sum()), rng.rand(len(ru))), n_tag - 1)
        uit = np.stack([ru, ri, tags], 1)
    return finish(ds, n_user, n_item, uit, n_tag)


SHAPES = {
    # BASELINE.json configs: total interactions; ~80 % land in train
    "lastfm": dict(n_user=1892, n_item=17632, n_edge=92834),
    "delicious_tags": dict(n_user=1867, n_item=69223, n_edge=104799, n_tag=40897),
    "gowalla": dict(n_user=29858, n_item=40981, n_edge=1027370),
    "amazon_book": dict(n_user=52643, n_item=91599, n_edge=2984108),
}


def synth_named(name, seed=2020):
    return synth_bipartite(seed=seed, **SHAPES[name])


def synth_bipartite_device(n_user, n_item, n_edge, device, seed=2020, zipf=0.9, sigma=1.0):
    """Device (torch) generator for the 1 B-edge shape (C5): returns unique, (u, i)-sorted train pairs as int64
    device tensors.  Counter-style: one pass, no host arrays, no global np.unique (SURVEY §8 d)."""
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    act = torch.exp(torch.randn(n_user, generator=g, device=device, dtype=torch.float64) * sigma)
    deg = torch.clamp((act / act.sum() * n_edge).round().to(torch.int64), 1, n_item // 2)
    users = torch.repeat_interleave(torch.arange(n_user, device=device), deg)
    m = users.numel()
    # inverse CDF of a truncated power law p(k) ~ k^-s on [1, n_item]
    a = 1.0 - zipf
    r = torch.rand(m, generator=g, device=device, dtype=torch.float64)
    rank = torch.pow(r * (float(n_item) ** a - 1.0) + 1.0, 1.0 / a).to(torch.int64).clamp_(1, n_item) - 1
    del r
    # scatter popularity over the id space with an affine bijection (multiplier coprime with n_item)
    mult = 2654435761 % n_item
    while np.gcd(mult, n_item) != 1:
        mult += 1
    items = (rank * mult + 12345) % n_item
    del rank
    key = torch.unique(users * n_item + items)        # sorted, duplicates removed
    del users, items
    return torch.div(key, n_item, rounding_mode="floor"), key % n_item


def all_neighbor_sample(matrix, max_deg):
    """data/utils.py:87-106 — one padded neighbour table: per row ``max_deg`` neighbour ids (+1; 0 = padding, only in
    empty rows) sampled WITH replacement when the row is shorter than the table, a random permutation prefix otherwise,
    and the matching integer edge weights.  Same ``np.random.choice`` calls in the same order as the reference (so the
    same numpy state gives the same tables), but straight from the CSR arrays — the reference densifies every row with
    ``matrix[i].toarray()``."""
    m = matrix.tocsr()
    m.sum_duplicates()
    n = m.shape[0]
    data = np.zeros((n, max_deg), dtype=np.int64)
    weight = np.zeros((n, max_deg), dtype=np.int64)
    indptr, indices, vals = m.indptr, m.indices, m.data
    for i in range(n):
        lo, hi = indptr[i], indptr[i + 1]
        keep = vals[lo:hi] != 0
        ids = indices[lo:hi][keep]
        x = len(ids)
        if x == 0:
            continue
        if x < max_deg:
            sample = np.random.choice(ids, max_deg)
        else:
            sample = np.random.choice(ids, max_deg, replace=False)
        data[i] = sample + 1
        weight[i] = vals[lo:hi][keep][np.searchsorted(ids, sample)].astype(np.int64)
    return [data, weight]


def get_all_neighbor(ds, fork_semantics=True, width=None):
    """data/tgcn_load.py:41-53 TGCN_load.get_all_neighbor: the six tables (ui, ut, iu, it, tu, ti).  The reference
    builds them in a forked worker (cpu_core == 1: one worker, the six matrices in order), i.e. from a COPY of numpy's
    global generator; ``fork_semantics`` reproduces that by restoring the global state afterwards.  The reference's
    tables are as wide as the largest stored row (thousands of columns on tag graphs) although TGCN reads the first
    ``neighbor_k`` columns only (tgcn.py:199); ``width`` caps them (different random stream, same distribution)."""
    mats = [ds.ui_adj, ds.ut_adj, ds.ui_adj.transpose(), ds.it_adj, ds.ut_adj.transpose(), ds.it_adj.transpose()]
    # tgcn_load.py:44: getnnz(1) of the COO as stored — duplicate (u, t) entries count, so the table can be wider than
    # the number of DISTINCT neighbours (rows are then filled by sampling with replacement)
    max_deg = [int(max(a.getnnz(1))) for a in mats]
    if width is not None:      # NOT stream-identical to the reference: tables only as wide as the model reads (neighbor_k)
        max_deg = [min(d, int(width)) for d in max_deg]
    saved = np.random.get_state() if fork_semantics else None
    out = [all_neighbor_sample(a, d) for a, d in zip(mats, max_deg)]
    if saved is not None:
        np.random.set_state(saved)
    return out


def get_all_neighbor_device(ds, width, device, seed=2020):
    """The six neighbour tables (ui, ut, iu, it, tu, ti) of data/tgcn_load.py:41-53 built on the device
    (``tagrec_neighbor_table``): K0 builds the weighted tripartite CSR from the (u, i) / (u, i, t) lists, then one launch
    per relation samples ``width`` (= neighbor_k, the columns TGCN reads, tgcn.py:199) neighbours per row.  Returns a
    list of ``(ids, weights)`` int64 device tensors [n_rows, width] — same contract as ``get_all_neighbor`` (ids + 1,
    0 = padding in empty rows only), same distribution, a Philox stream instead of numpy's.  No per-row host loop:
    112 K rows x 6 relations take microseconds instead of the reference's minutes."""
    from . import adj
    from ._lib import check, lib, ptr, stream_ptr
    U, I, Tn = int(ds.num["user"]), int(ds.num["item"]), int(ds.num["tag"])
    g = adj.build_csr(U, I, (ds.ui_adj.row, ds.ui_adj.col), "plain", device, Tn,
                      (ds.ut_adj.row, ds.ut_adj.col), (ds.it_adj.row, ds.it_adj.col))
    off = [0, U, U + I, U + I + Tn]
    rel = [(0, 1), (0, 2), (1, 0), (1, 2), (2, 0), (2, 1)]                  # (row type, column type): ui ut iu it tu ti
    out = []
    for k, (a, b) in enumerate(rel):
        n = off[a + 1] - off[a]
        ids = torch.empty((n, width), dtype=torch.int64, device=device)
        wts = torch.empty((n, width), dtype=torch.int64, device=device)
        check(lib().tagrec_neighbor_table(ptr(g.rowptr), ptr(g.col), ptr(g.val), off[a], n, off[b], off[b + 1], int(width),
                                          int(seed), k, ptr(ids), ptr(wts), stream_ptr(torch.device(device))),
              "tagrec_neighbor_table")
        out.append((ids, wts))
    return out
