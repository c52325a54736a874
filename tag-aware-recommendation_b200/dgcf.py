"""DGCF — drop-in for model/dgcf.py (same constructor, ``embed`` ParameterList, forward / loss / predict_rating /
get_ego_embed), with the 4-intent routing propagation on the K5 kernels (csrc/routing.cu).

``loss`` takes the ``(data, cor)`` tuple DGCF_training_data yields (dgcf.py:116); the L2 term reads the EGO rows
(dgcf.py:126-130); the correlation loss is disabled in the reference (commented out, dgcf.py:131-145) and here.
"""
import torch
import torch.nn as nn

from . import adj as utils
from . import config
from .eval_ops import EvalMixin
from .functional import BprLossFn
from .routing import DgcfPropagateFn


class DGCF(EvalMixin, nn.Module):
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
        self.num_layer = len(cfg['dim_layer_list'])
        self.device = cfg['device']
        self.norm_type = cfg['norm_type']
        self.split_adj_k = cfg["split_adj_k"]
        self.factor_k = cfg['factor_k']
        self.iterate_k = cfg['iterate_k']
        self.dim_k = self.dim_latent // self.factor_k
        self.reg = cfg['reg']
        self.cor_reg = cfg['cor_reg']
        self.loss_func = cfg['mul_loss_func']
        self.use_tag = cfg['use_tag']
        if self.factor_k != 4 or self.dim_latent != 64:
            raise NotImplementedError("the routing kernels are built for factor_k == 4 and dim_latent == 64 "
                                      "(utility/config.py:15-22 defaults)")

    def _init_weight(self):
        self.embed = nn.ParameterList()
        for num in self.num_list:
            self.embed.append(nn.Parameter(torch.empty(num, self.dim_latent)))
        for p in self.parameters():
            nn.init.xavier_uniform_(p)

    def _final_table(self, ego=None):
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.embed):
            ego = torch.cat(list(self.embed), dim=0) if ego is None else ego
            return DgcfPropagateFn.apply(self.norm_adj, self.num_layer, self.iterate_k, ego)
        key = tuple((p.data_ptr(), p._version) for p in self.embed)
        if self._cache is None or self._cache[0] != key:
            with torch.no_grad():
                ego = torch.cat([p.detach() for p in self.embed], dim=0)
                self._cache = (key, DgcfPropagateFn.apply(self.norm_adj, self.num_layer, self.iterate_k, ego))
        return self._cache[1]

    def forward(self, out_A=False):
        if out_A:
            # dgcf.py:57,62-63,81: per layer, per factor, the sparse adjacency carrying that factor's routing weights
            # (softmax over the factors of the edge logits) of the layer's last iteration
            from .routing import dgcf_propagate
            with torch.no_grad():
                ego = torch.cat([p.detach() for p in self.embed], dim=0)
                _, _, _, weights = dgcf_propagate(self.norm_adj, self.num_layer, self.iterate_k, ego, True)
            idx = self.norm_adj._indices()
            return [[torch.sparse_coo_tensor(idx, w[:, i].contiguous(), self.norm_adj.shape) for i in range(self.factor_k)]
                    for w in weights]
        return torch.split(self._final_table(), self.num_list, dim=0)

    def get_ego_embed(self):
        return list(self.embed)

    def loss(self, batch_data):
        self._cache = None               # a training step follows: the cached inference table goes stale
        data, cor = batch_data
        ego = torch.cat(list(self.embed), dim=0)
        final = self._final_table(ego)
        return BprLossFn.apply(data, self.num_list[0], self.reg, self.loss_func, final, ego)

    def predict_rating(self, users):
        all_users, all_items = self.forward()[:2]
        return torch.sigmoid(torch.matmul(all_users[users], all_items.t()))
