"""Oracle: the reference's LightGCN training step AS THE REFERENCE EXECUTES IT ON CPU (the timing baseline).

Restates, op for op, model/lightgcn.py:49-82 + model/help/adj.py:144-167 + model/help/loss.py:4-32 +
training/basic_train.py:14-27 with the same torch calls the reference makes: an UN-coalesced sparse COO adjacency
(adj.py:149), torch.sparse.mm per layer, F.normalize, stack+mean, advanced-index gathers, softplus BPR,
norm(2).pow(2) regulariser, autograd backward, torch.optim.Adam, and the three per-step ``.cpu().item()`` syncs.
Used by bench.py's cpu_baseline leg / --impl reference, and by tests as an autograd cross-check of the closed forms.
TEST / BASELINE INFRASTRUCTURE — see oracle/__init__.py.
"""
import numpy as np
import torch
import torch.nn.functional as F


def coo_adjacency(n, rowptr, col, val):
    """adj.py:144-150 sp2tensor: COO built from (row, col, val) triplets, never coalesced."""
    row = torch.from_numpy(np.repeat(np.arange(n, dtype=np.int64), np.diff(rowptr)))
    c = torch.from_numpy(np.asarray(col, dtype=np.int64))
    v = torch.from_numpy(np.asarray(val, dtype=np.float32))
    return torch.sparse_coo_tensor(torch.stack([row, c]), v, (n, n))


class LightGCNStep:
    def __init__(self, n_user, n_item, csr, dim=64, n_layer=3, reg=0.0, lr=0.01, seed=2020, loss_func="softplus"):
        n, rowptr, col, val = csr
        self.adj = coo_adjacency(n, rowptr, col, val)
        self.num_list = [n_user, n_item]
        self.n_layer, self.reg, self.loss_func = n_layer, reg, loss_func
        torch.manual_seed(seed)
        self.embed = [torch.nn.Parameter(torch.empty(k, dim)) for k in self.num_list]      # lightgcn.py:39-47
        for p in self.embed:
            torch.nn.init.xavier_uniform_(p)
        self.opt = torch.optim.Adam(self.embed, lr=lr)                                      # com.py:25

    def forward(self):
        all_embed = torch.cat(self.embed, dim=0)                                            # lightgcn.py:52
        layers = [all_embed]
        for _ in range(self.n_layer):
            all_embed = torch.sparse.mm(self.adj, all_embed)                                # adj.py:166
            all_embed = F.dropout(all_embed, p=0.0, training=True)                          # lightgcn.py:56
            layers.append(F.normalize(all_embed, p=2, dim=1))                               # lightgcn.py:57
        all_embed = torch.mean(torch.stack(layers, dim=1), dim=1)                           # lightgcn.py:60
        return torch.split(all_embed, self.num_list, dim=0)

    def loss(self, batch):
        users, pos, neg = batch.T                                                           # lightgcn.py:69
        all_users, all_items = self.forward()[:2]
        ue, pe, ne = all_users[users], all_items[pos], all_items[neg]
        ps, ns = (ue * pe).sum(1), (ue * ne).sum(1)                                         # loss.py:5-6
        if self.loss_func == "logsigmoid":
            loss = -F.logsigmoid(ps - ns).mean()
        else:
            loss = F.softplus(ns - ps).mean()
        eu, ei = self.embed
        reg = 0
        for e in (eu[users], ei[pos], ei[neg]):                                             # loss.py:27-32
            reg = reg + e.norm(2).pow(2)
        return loss, self.reg * (0.5 * reg / float(len(users)))

    def step(self, batch):
        """training/basic_train.py:15-27 for one mini-batch; returns the logged floats."""
        lossx = self.loss(batch)
        parts = [x.cpu().item() for x in lossx]
        loss = sum(lossx)
        self.opt.zero_grad()
        loss.backward()
        self.opt.step()
        return parts, loss.cpu().item()


# ----------------------------------------------------------------------------------------------------------------
# The other model families of BASELINE.json (configs[1..3]) as the reference executes them on CPU: same torch calls in
# the same order (sparse COO products, per-factor sparse tensor construction, [N, k, d] gathers, Conv2d), autograd
# backward, torch.optim.Adam over all parameters, the per-step ``.cpu().item()`` syncs.  bench.py's CPU arm only.
def _xavier(shape):
    p = torch.nn.Parameter(torch.empty(*shape))
    torch.nn.init.xavier_uniform_(p)
    return p


def _bpr(users_emb, pos_emb, neg_emb, kind):
    ps, ns = (users_emb * pos_emb).sum(1), (users_emb * neg_emb).sum(1)                   # loss.py:5-6
    return -F.logsigmoid(ps - ns).mean() if kind == "logsigmoid" else F.softplus(ns - ps).mean()


def _l2(*embs):
    reg = 0
    for e in embs:                                                                         # loss.py:27-32
        reg = reg + e.norm(2).pow(2)
    return 0.5 * reg / float(embs[0].shape[0])


class _PortStep:
    loss_func = "softplus"
    reg_on_final = False

    def parameters(self):
        raise NotImplementedError

    def forward(self):
        raise NotImplementedError

    def _finish_init(self, lr):
        self.opt = torch.optim.Adam(self.parameters(), lr=lr)                              # com.py:25

    def loss(self, batch):
        users, pos, neg = batch.T
        all_users, all_items = self.forward()[:2]
        ue, pe, ne = all_users[users], all_items[pos], all_items[neg]
        loss = _bpr(ue, pe, ne, self.loss_func)
        if self.reg_on_final:                                                              # ngcf.py:103, tgcn.py:247
            reg = _l2(ue, pe, ne)
        else:                                                                              # dgcf.py:126-130
            eu, ei = self.embed[:2]
            reg = _l2(eu[users], ei[pos], ei[neg])
        return loss, self.reg * reg

    def step(self, batch):
        lossx = self.loss(batch)
        parts = [x.cpu().item() for x in lossx]                                            # basic_train.py:16
        loss = sum(lossx)
        self.opt.zero_grad()
        loss.backward()
        self.opt.step()
        return parts, loss.cpu().item()                                                    # basic_train.py:27


class NGCFStep(_PortStep):
    """model/ngcf.py:62-105 (bi_agg, norm_type 'ngcf' = D^-1 A + I, logsigmoid BPR, L2 on the propagated rows)."""
    loss_func, reg_on_final = "logsigmoid", True

    def __init__(self, n_user, n_item, csr, dims=(64, 64, 64, 64), reg=0.0, lr=0.01, seed=2020):
        n, rowptr, col, val = csr
        self.adj = coo_adjacency(n, rowptr, col, val)
        self.num_list, self.reg, self.n_layer = [n_user, n_item], reg, len(dims) - 1
        torch.manual_seed(seed)
        self.embed = [_xavier((k, dims[0])) for k in self.num_list]
        self.mat = {}
        for k in range(self.n_layer):                                                      # ngcf.py:45-56
            for w in ("W1", "b1", "W2", "b2"):
                self.mat[f"{w}_{k}"] = _xavier((dims[k] if w[0] == "W" else 1, dims[k + 1]))
        self._finish_init(lr)

    def parameters(self):
        return list(self.embed) + list(self.mat.values())

    def forward(self):
        all_embed = torch.cat(self.embed, dim=0)
        outs = [all_embed]
        for k in range(self.n_layer):                                                      # ngcf.py:73-88
            nei = torch.sparse.mm(self.adj, all_embed)
            s = F.leaky_relu(torch.matmul(nei + all_embed, self.mat[f"W1_{k}"] + self.mat[f"b1_{k}"]), 0.2)
            b = F.leaky_relu(torch.matmul(torch.mul(nei, all_embed), self.mat[f"W2_{k}"] + self.mat[f"b2_{k}"]), 0.2)
            all_embed = F.dropout(s + b, p=0.0, training=True)
            outs.append(F.normalize(all_embed, p=2, dim=1))
        return torch.split(torch.cat(outs, dim=1), self.num_list, dim=0)


class DGCFStep(_PortStep):
    """model/dgcf.py:49-145 (4 intents, 2 routing iterations, 'plain' adjacency: only the indices are used)."""

    def __init__(self, n_user, n_item, csr, dim=64, n_layer=3, factor_k=4, iterate_k=2, reg=0.0, lr=0.01, seed=2020):
        n, rowptr, col, _ = csr
        row = torch.from_numpy(np.repeat(np.arange(n, dtype=np.int64), np.diff(rowptr)))
        self.indices = torch.stack([row, torch.from_numpy(np.asarray(col, dtype=np.int64))])
        self.shape = (n, n)
        self.num_list, self.reg = [n_user, n_item], reg
        self.n_layer, self.factor_k, self.iterate_k, self.dim_k = n_layer, factor_k, iterate_k, dim // factor_k
        torch.manual_seed(seed)
        self.embed = [_xavier((k, dim)) for k in self.num_list]
        self._finish_init(lr)

    def parameters(self):
        return list(self.embed)

    def _factor_update(self, a_factor, ego_split):                                         # dgcf.py:91-110
        adj = torch.sparse_coo_tensor(self.indices, a_factor.detach().cpu(), self.shape)
        col_sum = torch.sparse.sum(adj, dim=1)
        val = 1 / torch.sqrt(col_sum._values())
        val[torch.isinf(val)] = 0.0
        d = torch.sparse_coo_tensor(col_sum._indices()[0].unsqueeze(0).repeat(2, 1), val, self.shape)
        f = torch.sparse.mm(d, ego_split)
        f = torch.sparse.mm(adj, f)
        f = torch.sparse.mm(d, f)
        head, tail = self.indices
        h_emb = F.normalize(f[head], p=2, dim=1)
        t_emb = F.normalize(ego_split[tail], p=2, dim=1)
        return f, torch.sum(torch.mul(h_emb, torch.tanh(t_emb)), dim=1)

    def forward(self):
        a_values = torch.ones(self.factor_k, self.indices.shape[1])
        ego = torch.cat(self.embed, dim=0)
        layers = [ego]
        for _ in range(self.n_layer):                                                      # dgcf.py:68-89
            split = torch.split(ego, self.dim_k, dim=1)
            layer_emb = []
            for t in range(self.iterate_k):
                a_factor = torch.softmax(a_values, dim=0)
                scores = []
                for i in range(self.factor_k):
                    f, sc = self._factor_update(a_factor[i], split[i])
                    scores.append(sc)
                    if t == self.iterate_k - 1:
                        layer_emb.append(f)
                a_values = a_values + torch.stack(scores, dim=0)
            ego = torch.cat(list(F.normalize(torch.stack(layer_emb), p=2, dim=2)), dim=1)
            layers.append(ego)
        return torch.split(torch.mean(torch.stack(layers, dim=1), dim=1), self.num_list, dim=0)


class TGCNStep(_PortStep):
    """model/tgcn.py:140-262 through oracle/tgcn.py's restatement of Attention1 / BasicLayer (k = neighbor_k columns of
    the padded neighbour tables, type attention, bit-/vector-level Conv2d, 2096 -> 64 fusion), logsigmoid BPR."""
    loss_func, reg_on_final = "logsigmoid", True

    def __init__(self, nums, tables, n_layer=2, neighbor_k=25, dim=64, dim_weight=10, dim_atten=32, num_bit_conv=32,
                 num_vec_conv=8, reg=0.0, lr=0.01, seed=2020):
        n_user, n_item, n_tag, n_weight = nums
        self.tables, self.n_layer, self.k, self.reg = tables, n_layer, neighbor_k, reg
        torch.manual_seed(seed)
        P = {"embed.user": _xavier((n_user, dim)), "embed.item": _xavier((n_item, dim)),
             "embed.tag": _xavier((n_tag, dim)), "embed.weight": _xavier((n_weight + 1, dim_weight))}
        feat = num_bit_conv * dim + num_vec_conv * 6
        for k in range(n_layer):                                                           # tgcn.py:40-76, 11-19
            pre = f"layer.{k}."
            for kind in ("user", "item", "tag"):
                a = f"{pre}atten1.{kind}."
                P[a + "W_1"], P[a + "W_2"] = _xavier((dim + dim_weight, dim_atten)), _xavier((dim, dim_atten))
                P[a + "b"], P[a + "v"] = _xavier((1, dim_atten)), _xavier((1, dim_atten))
            P[pre + "U"], P[pre + "q"], P[pre + "p"] = _xavier((dim, dim_atten)), _xavier((1, dim_atten)), _xavier((1, dim_atten))
            P[pre + "conv.bit_level.weight"] = _xavier((num_bit_conv, 1, 3, 1))
            for j in (1, 2, 3):
                P[pre + f"conv.vec_level.conv_{j}.weight"] = _xavier((num_vec_conv, 1, j, dim))
            P[pre + "Wf"], P[pre + "bf"] = _xavier((feat, dim)), _xavier((1, dim))
        self.P = P
        self.embed = [P["embed.user"], P["embed.item"], P["embed.tag"]]
        self._finish_init(lr)

    def parameters(self):
        return list(self.P.values())

    def forward(self):
        from .tgcn import tgcn_forward
        return tgcn_forward(self.P, self.tables, self.n_layer, self.k)
