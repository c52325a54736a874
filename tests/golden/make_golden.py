#!/usr/bin/env python
"""Generate golden fixtures by RUNNING THE UNMODIFIED REFERENCE on CPU.

Runs only in the build container (needs /root/reference, read-only).  The GPU
box never executes this file; it only reads the .npz files written next to it.

Recipe = SURVEY.md Appendix C: four shims applied outside the reference tree
(stub tensorboardX, collections.Iterable, sys.argv before import, np.int), never
import com.py, compose objects as com.py:21-29 does.

    python tests/golden/make_golden.py            # rewrites tests/golden/*.npz

Every array saved here is an output of reference code (cited per block) or an
input that was fed to it.
"""
import collections
import collections.abc
import os
import sys
import types

import numpy as np

REF = os.environ.get("TAGREC_REFERENCE", "/root/reference")
OUT = os.path.dirname(os.path.abspath(__file__))


def _install_shims(model="lightgcn"):
    collections.Iterable = collections.abc.Iterable          # utility/utils.py:6
    np.int = int                                             # data/utils.py:73-74
    tb = types.ModuleType("tensorboardX")                    # utility/word.py:1
    tb.SummaryWriter = object
    sys.modules["tensorboardX"] = tb
    if REF not in sys.path:
        sys.path.insert(0, REF)
    sys.argv = ["golden", "--model", model, "--use_tag", "", "--cpu_core", "1",
                "--dim_layer_list", "[64,64,64]", "--topks", "[5,20]"]


_install_shims()
import scipy.sparse as sp  # noqa: E402
import torch  # noqa: E402

torch.set_num_threads(1)
from utility.word import CFG  # noqa: E402
from utility.utils import init_seed  # noqa: E402
from utility.config import dict_map  # noqa: E402
import model.help as H  # noqa: E402
from model.lightgcn import LightGCN  # noqa: E402
from model.ngcf import NGCF  # noqa: E402
from model.dgcf import DGCF  # noqa: E402
from model.disengcn import DisenGCN  # noqa: E402
from model.tgcn import TGCN  # noqa: E402
import train_data.utils as TU  # noqa: E402
from train_data.bpr_training_data import BPR_training_data  # noqa: E402
import training.basic_test as BT  # noqa: E402
import training.utils as TRU  # noqa: E402
import data.utils as DU  # noqa: E402


def set_cfg(model, **kw):
    """Mutate the reference's global CFG in place (classes read it at construction)."""
    base = dict(train_batch=64, test_batch=16, has_val=False, use_tag=False, topks=[5, 20], lr=0.01, reg=0.0,
                cor_reg=0, dim_latent=64, dim_layer_list=[64, 64, 64], message_drop_list=[0., 0., 0.], node_drop=0.,
                seed=2020, cpu_core=1, split_adj_k=1, device=torch.device("cpu"), model=model)
    CFG.update(base)
    CFG.update(dict_map[model])
    CFG.update(kw)


class Mock:
    """Plain stand-in for data.TGCN_load (no files needed)."""
    pass


class Args:
    pool = None
    writer = None
    out_dir = "/tmp"


def make_dataset(seed, U, I, T, n_edge, n_uit, p_test=0.2):
    """Small HetRec-shaped dataset.  Mirrors what data/cf_load.py + data/tgcn_load.py hand to the models:
    user_items dicts, edge_index, ui/ut/it COO float32 with duplicate (u,t)/(i,t) pairs kept (summed later
    by the reference's COO->LIL conversion, data/utils.py:50-53)."""
    rng = np.random.RandomState(seed)
    # zipf-ish item popularity, lognormal-ish user activity; last user and last two items stay isolated in train
    pu = rng.lognormal(0, 1, U - 1); pu /= pu.sum()
    pi = 1.0 / np.arange(1, I - 1) ** 0.9; pi /= pi.sum()
    pairs = set()
    while len(pairs) < n_edge:
        pairs.add((int(rng.choice(U - 1, p=pu)), int(rng.choice(I - 2, p=pi))))
    pairs = sorted(pairs)
    train, test = {}, {}
    for u, i in pairs:
        (test if rng.rand() < p_test else train).setdefault(u, []).append(i)
    # a user with test items but no train items (basic_test.py:37 -> no mask), and vice versa
    test.setdefault(U - 1, []).append(int(I - 1))
    for u in list(train):
        rng.shuffle(train[u])        # dict order is file order, not sorted (data/utils.py:34 uses set())
    d = Mock()
    d.user_items = {"train": train, "test": test}
    e = [(u, i) for u, its in train.items() for i in its]
    d.edge_index = {"train": np.array(e, dtype=np.int64)}
    d.num = {"user": U, "item": I, "tag": T}
    d.ui_adj = DU.to_sparse_adj(d.edge_index["train"][:, 0], d.edge_index["train"][:, 1], (U, I))
    # (u,i,t) assignments on train edges, unique triples (data/utils.py:11-13 np.unique)
    pt = 1.0 / np.arange(1, T) ** 0.7; pt /= pt.sum()
    uit = set()
    while len(uit) < n_uit:
        u, i = e[rng.randint(len(e))]
        uit.add((u, i, int(rng.choice(T - 1, p=pt))))
    d.uit_data = np.array(sorted(uit), dtype=np.int32)
    d.ut_adj = DU.to_sparse_adj(d.uit_data[:, 0], d.uit_data[:, 2], (U, T))
    d.it_adj = DU.to_sparse_adj(d.uit_data[:, 1], d.uit_data[:, 2], (I, T))
    d.num["weight"] = int(max(d.ui_adj.max(), d.ut_adj.tocsr().max(), d.it_adj.tocsr().max()))
    return d


def dict_to_csr(dic, n):
    ptr = np.zeros(n + 1, dtype=np.int64)
    flat = []
    for u in range(n):
        its = dic.get(u, [])
        ptr[u + 1] = ptr[u] + len(its)
        flat.extend(its)
    return ptr, np.array(flat, dtype=np.int64)


def dataset_arrays(d, prefix=""):
    out = {}
    for part in ("train", "test"):
        keys = np.array(list(d.user_items[part].keys()), dtype=np.int64)
        ptr, flat = dict_to_csr(d.user_items[part], d.num["user"])
        out[f"{prefix}{part}_keys"] = keys            # dict key order matters for eval batching
        out[f"{prefix}{part}_ptr"] = ptr
        out[f"{prefix}{part}_items"] = flat           # list order inside a user matters for the sampler
    out[f"{prefix}edge_index_train"] = d.edge_index["train"]
    out[f"{prefix}uit_data"] = d.uit_data
    out[f"{prefix}num"] = np.array([d.num["user"], d.num["item"], d.num["tag"], d.num["weight"]], dtype=np.int64)
    return out


def coo_of(t):
    return t._indices().numpy().astype(np.int64), t._values().numpy()


# ---------------------------------------------------------------------------------------------------------------
def golden_adjacency(d):
    """model/help/adj.py:38-46 creat_adj for every norm_type x use_tag; plus split_adj_k row folds."""
    out = {}
    for use_tag in (False, True):
        for nt in ("bi_norm", "si_norm", "si_norm_self", "ngcf", "plain"):
            with np.errstate(divide="ignore"):
                adj = H.creat_adj(d, use_tag, nt, 1, torch.device("cpu"))
            idx, val = coo_of(adj)
            tag = f"adj_{'uit' if use_tag else 'ui'}_{nt}"
            out[tag + "_row"], out[tag + "_col"], out[tag + "_val"] = idx[0], idx[1], val
    with np.errstate(divide="ignore"):
        folds = H.creat_adj(d, False, "bi_norm", 3, torch.device("cpu"))
    out["adj_fold3_rows"] = np.array([f.shape[0] for f in folds], dtype=np.int64)
    out["adj_fold3_nnz"] = np.array([f._nnz() for f in folds], dtype=np.int64)
    return out


def _batch(d, rng, b):
    """A (b,3) triple batch with repeated users/items, negatives not in train (shape of bpr_training_data.py:44)."""
    e = d.edge_index["train"]
    sel = rng.randint(0, len(e), b)
    neg = np.empty(b, dtype=np.int64)
    for k, s in enumerate(sel):
        u = e[s, 0]
        while True:
            j = rng.randint(0, d.num["item"])
            if j not in d.user_items["train"][u]:
                neg[k] = j
                break
    return np.stack([e[sel, 0], e[sel, 1], neg], 1).astype(np.int64)


def golden_model(d, name, cls, use_tag, reg, tag, extra_cfg=None):
    """<model>.forward() / .loss(batch) + autograd grads / .predict_rating(users) from the reference class itself."""
    set_cfg(name, use_tag=use_tag, reg=reg, **(extra_cfg or {}))
    init_seed(2020)
    with np.errstate(divide="ignore"):
        m = cls(d)
    out = {}
    for k, v in m.state_dict().items():
        out[f"{tag}_param_{k}"] = v.detach().numpy().copy()
    m.train()
    fw = m.forward()
    for k, t in enumerate(fw):
        out[f"{tag}_fwd_{k}"] = t.detach().numpy().copy()
    rng = np.random.RandomState(7)
    batch = _batch(d, rng, 48)
    out[f"{tag}_batch"] = batch
    bt = torch.tensor(batch, dtype=torch.long)
    if name in ("dgcf", "disengcn"):
        lossx = m.loss((bt, None))
    else:
        lossx = m.loss(bt)
    out[f"{tag}_loss"] = np.array([x.item() for x in lossx], dtype=np.float64)
    m.zero_grad()
    sum(lossx).backward()
    for k, p in m.named_parameters():
        g = p.grad if p.grad is not None else torch.zeros_like(p)
        out[f"{tag}_grad_{k}"] = g.detach().numpy().copy()
    m.eval()
    users = torch.tensor([0, 3, 5, d.num["user"] - 1, 3], dtype=torch.long)
    out[f"{tag}_pred_users"] = users.numpy()
    with torch.no_grad():
        out[f"{tag}_pred"] = m.predict_rating(users).numpy().copy()
    return out, m


def golden_training(d, tag):
    """3 Adam steps through training/basic_train.py:10-30 epoch_training body, fixed batches."""
    from training.basic_train import epoch_training
    set_cfg("lightgcn", use_tag=False, reg=1e-3, train_batch=64)
    init_seed(2020)
    with np.errstate(divide="ignore"):
        m = LightGCN(d)
    opt = torch.optim.Adam(m.parameters(), lr=CFG["lr"])
    rng = np.random.RandomState(11)
    triples = _batch(d, rng, 64 * 3 + 20)          # 212 rows -> batches 64,64,84 + tail 20 again (abstract.py:17-23)

    class Fixed:
        batch_size = 64

        def reset(self):
            self.all_train_data = torch.tensor(triples, dtype=torch.long)

        mini_batch = BPR_training_data.mini_batch

    m.train()
    losses = epoch_training(Fixed(), m.loss, opt)
    out = {f"{tag}_triples": triples, f"{tag}_losses": np.array(losses, dtype=np.float64)}
    for k, v in m.state_dict().items():
        out[f"{tag}_after_{k}"] = v.detach().numpy().copy()
    return out


def golden_sampler(d, tag):
    """train_data/bpr_training_data.py:12-45 with cpu_core=1 (the only reproducible setting)."""
    set_cfg("lightgcn", train_batch=64, cpu_core=1)
    init_seed(2020)
    s = BPR_training_data(d, Args())
    out = {f"{tag}_first": s.all_train_data.numpy().copy()}
    s.reset()
    out[f"{tag}_second"] = s.all_train_data.numpy().copy()
    out[f"{tag}_batch_sizes"] = np.array([len(b) for b in s.mini_batch()], dtype=np.int64)
    # primitive streams (train_data/utils.py:23,52-55) for pinning the MT19937 restatement
    np.random.seed(99)
    out[f"{tag}_randint_1000"] = np.array([np.random.randint(0, 1000) for _ in range(64)], dtype=np.int64)
    out[f"{tag}_randint_17632"] = np.array([np.random.randint(0, 17632) for _ in range(64)], dtype=np.int64)
    idx = np.arange(50)
    np.random.shuffle(idx)
    out[f"{tag}_shuffle_50"] = idx
    return out


def golden_eval(d, m, tag, topks):
    """training/basic_test.py:30-80 epoch_test (recall/precision/hr/ndcg/auc) on the reference model."""
    set_cfg("lightgcn", test_batch=16, topks=topks)
    m.eval()
    res = BT.epoch_test(m, d.user_items["train"], d.user_items["test"], Args())
    out = {f"{tag}_topks": np.array(topks, dtype=np.int64)}
    for k, v in res.items():
        out[f"{tag}_{k}"] = np.array(v, dtype=np.float64)
    # masked score rows for every test user, in dict key order (basic_test.py:37-47)
    users = list(d.user_items["test"].keys())
    with torch.no_grad():
        r = m.predict_rating(torch.tensor(users, dtype=torch.long)).clone()
    for row, u in enumerate(users):
        its = d.user_items["train"].get(u, [])
        r[row, its] = -(1 << 10)
    out[f"{tag}_masked_scores"] = r.numpy().copy()
    out[f"{tag}_users"] = np.array(users, dtype=np.int64)
    grp = TRU.user_group_split(d.user_items["test"], d.user_items["train"], 4)
    out[f"{tag}_group_keys"] = np.array(list(grp.keys()), dtype=np.int64)
    out[f"{tag}_group_sizes"] = np.array([len(v) for v in grp.values()], dtype=np.int64)
    return out


def golden_tgcn(d, tag):
    """model/tgcn.py forward/loss with the reference's own neighbour tables (data/tgcn_load.py:41-53)."""
    from data.tgcn_load import TGCN_load
    d.args = Args()
    d.cpu_core = 1
    d.get_all_neighbor = types.MethodType(TGCN_load.get_all_neighbor, d)
    set_cfg("tgcn", use_tag=True, reg=1e-3, dim_layer_list=[64, 64], neighbor_k=5)
    init_seed(2020)
    m = TGCN(d)
    out = {}
    names = ["ui", "ut", "iu", "it", "tu", "ti"]
    for n, (idx, w) in zip(names, m.all_sample):
        out[f"{tag}_nbr_{n}"] = np.asarray(idx, dtype=np.int64)
        out[f"{tag}_nbw_{n}"] = np.asarray(w, dtype=np.int64)
    for k, v in m.state_dict().items():
        out[f"{tag}_param_{k}"] = v.detach().numpy().copy()
    m.train()
    fw = m.forward()
    for k, t in enumerate(fw):
        out[f"{tag}_fwd_{k}"] = t.detach().numpy().copy()
    rng = np.random.RandomState(7)
    batch = _batch(d, rng, 48)
    out[f"{tag}_batch"] = batch
    lossx = m.loss(torch.tensor(batch, dtype=torch.long))
    out[f"{tag}_loss"] = np.array([x.item() for x in lossx], dtype=np.float64)
    m.zero_grad()
    sum(lossx).backward()
    for k, p in m.named_parameters():
        g = p.grad if p.grad is not None else torch.zeros_like(p)
        out[f"{tag}_grad_{k}"] = g.detach().numpy().copy()
    return out


def main():
    os.chdir("/tmp")
    # ---- tiny: every reference code path on the hot path, small enough for pure-python checks ----
    d = make_dataset(seed=1, U=40, I=60, T=25, n_edge=420, n_uit=500)
    out = dataset_arrays(d)
    out.update(golden_adjacency(d))
    g, m_l = golden_model(d, "lightgcn", LightGCN, False, 1e-3, "lgcn")
    out.update(g)
    out.update(golden_eval(d, m_l, "eval", [5, 20]))
    g, _ = golden_model(d, "lightgcn", LightGCN, True, 1e-3, "lgcn_tag")
    out.update(g)
    g, _ = golden_model(d, "lightgcn", LightGCN, False, 1e-3, "lgcn_logsig", {"mul_loss_func": "logsigmoid"})
    out.update(g)
    g, _ = golden_model(d, "ngcf", NGCF, False, 1e-3, "ngcf")
    out.update(g)
    g, _ = golden_model(d, "ngcf", NGCF, True, 1e-3, "ngcf_tag")
    out.update(g)
    g, _ = golden_model(d, "dgcf", DGCF, False, 1e-3, "dgcf")
    out.update(g)
    g, _ = golden_model(d, "disengcn", DisenGCN, True, 1e-3, "disengcn")
    out.update(g)
    out.update(golden_training(d, "train"))
    out.update(golden_sampler(d, "sampler"))
    np.savez_compressed(os.path.join(OUT, "tiny.npz"), **out)
    np.savez_compressed(os.path.join(OUT, "tiny_tgcn.npz"), **golden_tgcn(d, "tgcn"))

    # ---- medium: enough items/users for a meaningful top-20 and batch-boundary behaviour ----
    d2 = make_dataset(seed=2, U=150, I=700, T=40, n_edge=3000, n_uit=2000)
    out2 = dataset_arrays(d2)
    g, m2 = golden_model(d2, "lightgcn", LightGCN, False, 0.0, "lgcn")
    # keep only what the eval test needs (params + propagated tables) to stay small
    out2.update({k: v for k, v in g.items() if "_param_" in k or "_fwd_" in k})
    out2.update(golden_eval(d2, m2, "eval", [10, 20]))
    ms = out2.pop("eval_masked_scores")
    # store the (-score, id) top-40 of the reference's masked scores instead of the full matrix
    order = np.lexsort((np.arange(ms.shape[1])[None, :].repeat(ms.shape[0], 0), -ms), axis=1)[:, :40]
    out2["eval_top40_ids"] = order.astype(np.int64)
    out2["eval_top40_scores"] = np.take_along_axis(ms, order, 1)
    out2.update(golden_sampler(d2, "sampler"))
    np.savez_compressed(os.path.join(OUT, "medium.npz"), **out2)
    for f in ("tiny.npz", "tiny_tgcn.npz", "medium.npz"):
        print(f, os.path.getsize(os.path.join(OUT, f)) // 1024, "KiB")


if __name__ == "__main__":
    main()
