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
