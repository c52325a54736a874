"""NGCF — drop-in for model/ngcf.py (same constructor, parameter names/shapes/creation order, forward / loss /
predict_rating / get_ego_emb), on the kernels of libtagrec_b200.so.

Per layer (ngcf.py:73-90):  nei = A E   -> K1 ``tagrec_spmm`` (backward: the same kernel on the values of A^T);
the two 64x64 products with the bias added to the WEIGHT matrix (ngcf.py:78,82), LeakyReLU(0.2), the row
normalisation and the concat -> K6 ``tagrec_ngcf_dense_fwd`` / ``tagrec_ngcf_dense_bwd`` (one fused pass each, weight
gradients included: no library GEMM on the 64-d path).
Loss (ngcf.py:95-105): K2 on the 256-d propagated rows, logsigmoid form, L2 term on the PROPAGATED rows.
Evaluation: K3 (fp32 CUDA-core tiles for the 256-d concat table).
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import adj as utils
from . import config
from .eval_ops import EvalMixin
from .functional import BprLossFn, NgcfDenseFn


class NGCF(EvalMixin, nn.Module):
    def __init__(self, data, args=None):
        super().__init__()
        self._config(config.current())
        if self.use_tag:
            self.num_list = [data.num['user'], data.num['item'], data.num['tag']]
        else:
            self.num_list = [data.num['user'], data.num['item']]
        self.norm_adj = getattr(data, "prebuilt_adj", None) or \
            utils.creat_adj(data, self.use_tag, self.norm_type, self.split_adj_k, self.device)
        self._cache = None
        self._init_weight()

    def _config(self, cfg):
        self.dim_latent = cfg['dim_latent']
        self.dim_layer_list = cfg['dim_layer_list']
        self.num_layer = len(self.dim_layer_list)
        self.dim_layer_list = [self.dim_latent] + list(self.dim_layer_list)
        self.agg_type = cfg['agg_type']
        self.device = cfg['device']
        self.message_drop_list = cfg['message_drop_list']
        self.norm_type = cfg['norm_type']
        self.split_adj_k = cfg["split_adj_k"]
        self.reg = cfg['reg']
        self.loss_func = cfg['mul_loss_func']
        self.use_tag = cfg['use_tag']

    def _init_weight(self):
        # ngcf.py:39-60 — creation order defines how torch.manual_seed maps to initial weights
        self.embed = nn.ParameterList()
        for num in self.num_list:
            self.embed.append(nn.Parameter(torch.empty(num, self.dim_latent)))
        self.mat = nn.ParameterDict()
        for k in range(self.num_layer):
            self.mat.update({
                f"W1_{k}": nn.Parameter(torch.empty(self.dim_layer_list[k], self.dim_layer_list[k + 1])),
                f"b1_{k}": nn.Parameter(torch.empty(1, self.dim_layer_list[k + 1])),
            })
            if self.agg_type == "bi_agg":
                self.mat.update({
                    f"W2_{k}": nn.Parameter(torch.empty(self.dim_layer_list[k], self.dim_layer_list[k + 1])),
                    f"b2_{k}": nn.Parameter(torch.empty(1, self.dim_layer_list[k + 1])),
                })
        for p in self.parameters():
            nn.init.xavier_uniform_(p)

    # ------------------------------------------------------------------------------------------------
    def _fused_ok(self):
        return all(d == 64 for d in self.dim_layer_list)

    def _bi_inter_embed_sharded(self, all_embed):
        """The same layers on a node-range sharded graph (distributed.shard_graph; multi-GPU, no reference equivalent):
        rank p computes rows R_p of every layer — K1 on its row block of A, K6 on those rows — and the layer outputs are
        all-gathered; in backward the dense weight gradients (partial sums over R_p) are all-reduced, the SpMM backward
        all-gathers the upstream gradient rows, and the embedding gradient rows are all-gathered at the end, so every
        replica sees exactly the single-GPU gradients."""
        from . import distributed as D
        graph = self.norm_adj
        comm, n = graph.comm, all_embed.shape[0]
        lo, hi = comm.lo, comm.hi
        all_embed = D.RowOwnedParamFn.apply(comm, all_embed)
        all_embed_list = [all_embed]
        for k in range(self.num_layer):
            nei = D.ShardedSpMMFn.apply(graph, all_embed)                      # rows [lo, hi) filled
            nei_l, e_l = nei[lo:hi], all_embed[lo:hi]
            w = {name: D.AllReduceGradFn.apply(comm, self.mat[f'{name}_{k}']) for name in ("W1", "b1", "W2", "b2")}
            p = self.message_drop_list[k] if self.training else 0.0
            if self._fused_ok() and p == 0.0:
                out_l, nrm_l = NgcfDenseFn.apply(nei_l, e_l, w["W1"], w["b1"], w["W2"], w["b2"])
            else:
                if p != 0.0:
                    raise NotImplementedError("message dropout on a sharded graph (the ranks would have to share the mask)")
                sum_embed = F.leaky_relu(torch.matmul(nei_l + e_l, w["W1"] + w["b1"]), 0.2)
                bi_embed = F.leaky_relu(torch.matmul(nei_l * e_l, w["W2"] + w["b2"]), 0.2)
                out_l = sum_embed + bi_embed
                nrm_l = F.normalize(out_l, p=2, dim=1)
            if k + 1 < self.num_layer:                                         # the last raw layer feeds nothing
                all_embed = D.GatherRowsFn.apply(comm, n, out_l)
            all_embed_list += [D.GatherRowsFn.apply(comm, n, nrm_l)]
        return torch.cat(all_embed_list, dim=1)

    def bi_inter_embed(self, all_embed):
        """ngcf.py:73-90."""
        from .distributed import is_sharded
        if is_sharded(self.norm_adj):
            return self._bi_inter_embed_sharded(all_embed)
        all_embed_list = [all_embed]
        for k in range(self.num_layer):
            nei_embed = utils.split_mm(self.norm_adj, all_embed)
            p = self.message_drop_list[k] if self.training else 0.0
            if self._fused_ok() and p == 0.0:
                all_embed, norm_embed = NgcfDenseFn.apply(nei_embed, all_embed, self.mat[f'W1_{k}'], self.mat[f'b1_{k}'],
                                                          self.mat[f'W2_{k}'], self.mat[f'b2_{k}'])
            else:   # other layer widths / message dropout: the same maths from torch ops around K1
                sum_embed = F.leaky_relu(torch.matmul(nei_embed + all_embed, self.mat[f'W1_{k}'] + self.mat[f'b1_{k}']), 0.2)
                bi_embed = F.leaky_relu(torch.matmul(nei_embed * all_embed, self.mat[f'W2_{k}'] + self.mat[f'b2_{k}']), 0.2)
                all_embed = F.dropout(sum_embed + bi_embed, p=self.message_drop_list[k], training=self.training)
                norm_embed = F.normalize(all_embed, p=2, dim=1)
            all_embed_list += [norm_embed]
        return torch.cat(all_embed_list, dim=1)

    def _final_table(self):
        if self.agg_type != "bi_agg":
            raise NotImplementedError
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            return self.bi_inter_embed(torch.cat(list(self.embed), dim=0))
        # inference: propagate once per parameter version (the reference re-propagates per user batch, ngcf.py:108)
        key = tuple((p.data_ptr(), p._version) for p in self.parameters())
        if self._cache is None or self._cache[0] != key:
            with torch.no_grad():
                self._cache = (key, self.bi_inter_embed(torch.cat(list(self.embed), dim=0)))
        return self._cache[1]

    def forward(self):
        return torch.split(self._final_table(), self.num_list, dim=0)

    def get_ego_emb(self):
        return list(self.embed)

    def loss(self, batch_data):
        self._cache = None               # a training step follows: the cached inference table goes stale
        final = self._final_table()
        return BprLossFn.apply(batch_data, self.num_list[0], self.reg, self.loss_func, final, final)

    def predict_rating(self, users):
        all_users, all_items = self.forward()[:2]
        return torch.sigmoid(torch.matmul(all_users[users], all_items.t()))
