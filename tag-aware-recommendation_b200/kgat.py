"""KGAT — drop-in for model/kgat.py (same constructor, parameters, state_dict keys, forward / loss / transe_loss /
get_embed / predict_rating) and for ``KGAT_training_data`` (train_data/transe_training_data.py:12-40).

Hot path: the relation-aware attention (kgat.py:63-78) gives one logit per stored edge; ``torch.sparse.softmax`` over
rows turns them into the propagation matrix (kgat.py:95-96), which then drives NGCF-style bi-interaction layers
(kgat.py:106-125).  Here the graph STRUCTURE is built once (CSR + the CSR of the transpose with the permutation between
them); per forward only the values change.  Propagation runs on K1 (``tagrec_spmm``) through an autograd node that
also returns the gradient of the edge values (the attention is NOT detached in the reference: kgat.py:95-99), the dense
half-layers on K6 when all widths are 64, the BPR loss on K2 and evaluation on K3 like every other model.

Reference behaviours reproduced on purpose (SURVEY A-table style):
* ``forward`` only propagates when ``agg_type == 'bi_inter'`` (kgat.py:99) and only then are ``W2_k`` / ``b2_k`` created
  (kgat.py:53); the stock overlay sets ``agg_type = 'bi_agg'`` (utility/config.py:58), under which the reference's KGAT
  returns the ego embeddings — so does this class (the unused attention is then not evaluated).
* edge arrays are indexed ``[:, 0]`` / ``[:, 1]`` (kgat.py:70-71) whatever their shape: ``TGCN_load.create_edge``
  (data/tgcn_load.py:55-70) returns [2, E] arrays, of which the reference therefore reads two pseudo-edges per relation;
  an [E, 2] edge list gives the intended model.  Same indexing here.
* the regulariser of ``transe_loss`` is weighted by ``cor_reg`` (kgat.py:158), not by ``transe_reg``.
* ``KGAT_training_data.mini_batch`` slides its window by ONE row per batch (``all_triplet[i:i + batch]``,
  transe_training_data.py:34-36) and ``reset()`` is a no-op.
"""
import time
from collections import defaultdict

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import config
from .adj import CsrGraph, spmm_raw
from .bpr_training_data import Abstract_training_data
from .eval_ops import EvalMixin
from .functional import BprLossFn, NgcfDenseFn


class _EdgeStructure:
    """CSR of the (row, col) pairs of all relations (duplicates merged, as ``torch.sparse.softmax`` coalesces), the CSR
    of its transpose, and the maps edge -> CSR slot and CSR slot -> transposed slot."""

    def __init__(self, n, row, col, num_list):
        dev = row.device
        key = row * n + col
        uniq, self.slot_of_edge = torch.unique(key, return_inverse=True)          # sorted: row-major, ascending columns
        self.n, self.nnz = n, int(uniq.numel())
        r, c = torch.div(uniq, n, rounding_mode="floor"), uniq % n
        self.row_ids = r
        rowptr = torch.zeros(n + 1, dtype=torch.int64, device=dev)
        rowptr[1:] = torch.cumsum(torch.bincount(r, minlength=n), 0)
        ones = torch.ones(self.nnz, dtype=torch.float32, device=dev)
        self.graph = CsrGraph(n, rowptr, c.to(torch.int32).contiguous(), ones, None, None, "plain", num_list)
        self.t_order = torch.argsort(c * n + r)                                   # transposed slot -> CSR slot
        rowptr_t = torch.zeros(n + 1, dtype=torch.int64, device=dev)
        rowptr_t[1:] = torch.cumsum(torch.bincount(c, minlength=n), 0)
        self.graph_t = CsrGraph(n, rowptr_t, r[self.t_order].to(torch.int32).contiguous(), ones.clone(), None, None,
                                "plain", num_list)


class _AttnSpmmFn(torch.autograd.Function):
    """y = A(val) x on K1 with gradients for BOTH operands: g_x = A^T g (K1 on the transposed structure),
    g_val[e] = <g[row_e], x[col_e]>."""

    @staticmethod
    def forward(ctx, struct, val, x):
        x = x.contiguous()
        struct.graph.val = struct.graph.val_t = val.detach().contiguous()
        y = spmm_raw(struct.graph, x)
        ctx.struct = struct
        ctx.save_for_backward(val, x)
        return y

    @staticmethod
    def backward(ctx, g):
        struct = ctx.struct
        val, x = ctx.saved_tensors
        g = g.contiguous()
        gx = gval = None
        if ctx.needs_input_grad[2]:
            struct.graph_t.val = struct.graph_t.val_t = val.detach()[struct.t_order].contiguous()
            gx = spmm_raw(struct.graph_t, g)
        if ctx.needs_input_grad[1]:
            gval = (g[struct.row_ids] * x[struct.graph.col.long()]).sum(1)
        return None, gval, gx


def _row_softmax(struct, logits):
    """torch.sparse.softmax(adj, dim=1) over the stored entries of each row (kgat.py:96)."""
    rid = struct.row_ids
    mx = torch.full((struct.n,), float("-inf"), dtype=logits.dtype, device=logits.device)
    mx = mx.scatter_reduce(0, rid, logits.detach(), reduce="amax")
    e = torch.exp(logits - mx[rid])
    s = torch.zeros(struct.n, dtype=logits.dtype, device=logits.device).index_add(0, rid, e)
    return e / s[rid]


class KGAT(EvalMixin, nn.Module):
    def __init__(self, data, args=None):
        super().__init__()
        self._config(config.current())
        self.num_user = data.num['user']
        self.num_entity = data.num['item'] + data.num['tag']
        self.num_relation = 6
        edge_index_dict = data.create_edge()                                   # kgat.py:19
        self.edge_index_dict = {k: torch.as_tensor(np.asarray(v)).to(self.device) for k, v in edge_index_dict.items()}
        self._struct = None
        self._cache = None
        self._init_weight()

    def _config(self, cfg):
        self.dim_latent = cfg['dim_latent']
        self.dim_relation = cfg['dim_relation']
        self.dim_layer_list = list(cfg['dim_layer_list'])
        self.num_layer = len(self.dim_layer_list)
        self.dim_layer_list = [self.dim_latent] + self.dim_layer_list
        self.agg_type = cfg['agg_type']
        self.device = cfg['device']
        self.message_drop_list = cfg['message_drop_list']
        self.split_adj_k = cfg["split_adj_k"]
        self.reg = cfg['reg']
        self.cor_reg = cfg['cor_reg']
        self.loss_func = cfg['mul_loss_func']

    def _init_weight(self):
        # kgat.py:38-61 — creation order defines how torch.manual_seed maps to the initial weights
        self.embed = nn.ParameterDict({
            "user": nn.Parameter(torch.empty(self.num_user, self.dim_latent)),
            "entity": nn.Parameter(torch.empty(self.num_entity, self.dim_latent)),
            "relation": nn.Parameter(torch.empty(self.num_relation, self.dim_relation)),
        })
        self.mat = nn.ParameterDict({
            "transE": nn.Parameter(torch.empty(self.num_relation, self.dim_latent, self.dim_relation)),
        })
        for k in range(self.num_layer):
            self.mat.update({
                f"W1_{k}": nn.Parameter(torch.empty(self.dim_layer_list[k], self.dim_layer_list[k + 1])),
                f"b1_{k}": nn.Parameter(torch.empty(1, self.dim_layer_list[k + 1])),
            })
            if self.agg_type == "bi_inter":
                self.mat.update({
                    f"W2_{k}": nn.Parameter(torch.empty(self.dim_layer_list[k], self.dim_layer_list[k + 1])),
                    f"b2_{k}": nn.Parameter(torch.empty(1, self.dim_layer_list[k + 1])),
                })
        for p in self.parameters():
            nn.init.xavier_uniform_(p)

    # ------------------------------------------------------------------------------------------------
    def _edges(self):
        """(row, col) per relation exactly as kgat.py:70-71 indexes them, and the merged structure (built once)."""
        rows = [self.edge_index_dict[k][:, 0].long() for k in self.edge_index_dict]
        cols = [self.edge_index_dict[k][:, 1].long() for k in self.edge_index_dict]
        dev = self.embed["user"].device
        if self._struct is None or self._struct.row_ids.device != dev:
            n = self.num_user + self.num_entity
            self._struct = _EdgeStructure(n, torch.cat(rows).to(dev), torch.cat(cols).to(dev),
                                          [self.num_user, self.num_entity])
        return [r.to(dev) for r in rows], [c.to(dev) for c in cols]

    def _attention(self, all_embed):
        """kgat.py:63-96: pai(h, r, t) = <e_t W_r, tanh(e_h W_r + e_r)> per stored edge, then a row softmax.  The node
        projections are formed once per relation (N x 64 x dim_relation) instead of once per edge."""
        rows, cols = self._edges()
        pai = []
        for k, (row, col) in zip(self.edge_index_dict.keys(), zip(rows, cols)):
            proj = torch.matmul(all_embed, self.mat['transE'][k])
            pai.append(torch.sum(proj[col] * torch.tanh(proj[row] + self.embed['relation'][k]), dim=1))
        val = torch.cat(pai)
        st = self._struct
        merged = torch.zeros(st.nnz, dtype=val.dtype, device=val.device).index_add(0, st.slot_of_edge, val)
        return _row_softmax(st, merged)

    def bi_inter_embed(self, att, all_embed):
        """kgat.py:106-125 (== ngcf.py:73-90 on the attention matrix)."""
        outs = [all_embed]
        fused = all(d == 64 for d in self.dim_layer_list)
        for k in range(self.num_layer):
            nei = _AttnSpmmFn.apply(self._struct, att, all_embed)
            p = self.message_drop_list[k] if self.training else 0.0
            if fused and p == 0.0:
                all_embed, norm = NgcfDenseFn.apply(nei, all_embed, self.mat[f'W1_{k}'], self.mat[f'b1_{k}'],
                                                    self.mat[f'W2_{k}'], self.mat[f'b2_{k}'])
            else:
                s = F.leaky_relu(torch.matmul(nei + all_embed, self.mat[f'W1_{k}'] + self.mat[f'b1_{k}']), 0.2)
                b = F.leaky_relu(torch.matmul(nei * all_embed, self.mat[f'W2_{k}'] + self.mat[f'b2_{k}']), 0.2)
                all_embed = F.dropout(s + b, p=self.message_drop_list[k], training=self.training)
                norm = F.normalize(all_embed, p=2, dim=1)
            outs.append(norm)
        return torch.cat(outs, dim=1)

    def _propagate(self):
        all_embed = torch.cat([self.embed['user'], self.embed['entity']], dim=0)
        if self.agg_type == "bi_inter":                                   # kgat.py:99 — otherwise the ego table is returned
            all_embed = self.bi_inter_embed(self._attention(all_embed), all_embed)
        return all_embed

    def forward(self):
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            table = self._propagate()
        else:
            if self._cache is None:
                with torch.no_grad():
                    self._cache = self._propagate()
            table = self._cache
        return torch.split(table, [self.num_user, self.num_entity], dim=0)

    def loss(self, batch_data):
        self._cache = None
        all_users, all_items = self.forward()[:2]
        final = torch.cat([all_users, all_items], dim=0)
        return BprLossFn.apply(batch_data, self.num_user, self.reg, self.loss_func, final, final)

    def get_embed(self, batch_data):
        """kgat.py:127-142."""
        head, rela, pos_tail, neg_tail = batch_data.T
        all_embed = torch.cat([self.embed['user'], self.embed['entity']], dim=0)
        r_e = self.embed['relation'][rela.long()]
        trans = self.mat['transE'][rela.long()]
        proj = lambda idx: torch.matmul(all_embed[idx.long()].unsqueeze(1), trans).squeeze()      # noqa: E731
        return proj(head), r_e, proj(pos_tail), proj(neg_tail)

    def transe_loss(self, batch_data):
        """kgat.py:155-162 — batch-sized, plain torch ops."""
        self._cache = None
        h_e, r_e, pos_t_e, neg_t_e = self.get_embed(batch_data)
        pos_score = torch.norm(h_e + r_e - pos_t_e, p=2, dim=1).pow(2)
        neg_score = torch.norm(h_e + r_e - neg_t_e, p=2, dim=1).pow(2)
        kg_loss = torch.mean(F.softplus(pos_score - neg_score))
        reg = 0
        for emb in (h_e, r_e, pos_t_e, neg_t_e):
            reg = reg + emb.norm(2).pow(2)
        return kg_loss, self.cor_reg * (0.5 * reg / float(h_e.shape[0]))

    def predict_rating(self, users):
        all_users, all_items = self.forward()[:2]
        return torch.sigmoid(torch.matmul(all_users[users], all_items.t()))


class KGAT_training_data(Abstract_training_data):
    """Drop-in for train_data/transe_training_data.py:12-40: (head, relation, tail) triples of all six relations, one
    negative tail per row that is not a tail of (head, relation) (train_data/utils.py:30-39), sampled batch by batch in
    the calling process with numpy's global generator — same stream as the reference for the same seed."""

    def __init__(self, data, args=None):
        super().__init__(args)
        cfg = config.current()
        self.batch_size = cfg['transe_batch']
        self.num = data.num['user'] + data.num['item'] + data.num['tag']
        edge_list = data.create_edge()
        kg = [np.vstack([edge_list[k], np.ones(edge_list[k].shape[1]) * k]) for k in edge_list.keys()]
        self.all_triplet = np.hstack(kg).transpose()[:, [0, 2, 1]]
        self.h_r_dict = defaultdict(dict)                                    # train_data/utils.py:42-48
        for h, r, t in self.all_triplet:
            self.h_r_dict[h].setdefault(r, []).append(t)
        self.tot_inter = self.all_triplet.shape[0] // self.batch_size
        start = time.time()
        print(f"TransE_training_data producer, tot_inter: {self.tot_inter},[mini_sample time:{time.time()-start}]")

    def reset(self):
        pass

    def mini_batch(self):
        for i in range(0, self.tot_inter):
            batch = self.all_triplet[i:i + self.batch_size]                  # (sic) the window moves by one row
            data = []
            for h, r, t in batch:
                while True:
                    idx = np.random.randint(0, self.num)
                    if idx not in self.h_r_dict[h][r]:
                        data.append([h, r, t, idx])
                        break
            yield torch.tensor(np.array(data), dtype=torch.long, device=self.device)
