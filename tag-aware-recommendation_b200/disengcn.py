"""DisenGCN — drop-in for model/disengcn.py (``embed`` ParameterList + ``layer`` ModuleList of ``Layer`` with
parameters ``W`` [4, 64, 16] and ``b`` [4, 1, 16]; forward / loss / predict_rating).

Per layer (disengcn.py:23-46): the per-factor projection ``all_emb @ (W + b)`` — bias added to the WEIGHT, SURVEY A3 —
is one [N,64] x [64,64] cuBLAS GEMM over the concatenated factor weights, LeakyReLU(0.2), then the per-factor
normalisation (R5/R6) and the neighbour routing (R4 + R3, ``DisenRouteFn``) run on the K5 kernels.  The output is the
LAST layer only (disengcn.py:97); the L2 term reads the propagated rows (disengcn.py:115).
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import adj as utils
from . import config
from .eval_ops import EvalMixin
from .functional import BprLossFn, skinny_mm
from .routing import ChunkNormFn, DisenRouteFn


class Layer(nn.Module):
    def __init__(self, fac_k, iter_k, in_dim, out_dim):
        super().__init__()
        self.fac_k, self.iter_k, self.in_dim, self.out_dim = fac_k, iter_k, in_dim, out_dim
        dim_k = int(out_dim / fac_k)
        self.W = nn.Parameter(torch.empty(fac_k, in_dim, dim_k))
        self.b = nn.Parameter(torch.empty(fac_k, 1, dim_k))

    def forward(self, adj, all_emb):
        # [K, in, dk] -> [in, K*dk]: column block k is factor k, so chunk k of a row == fac_emb[k] of the reference
        wc = (self.W + self.b).permute(1, 0, 2).reshape(self.in_dim, self.out_dim)
        fac = F.leaky_relu(skinny_mm(all_emb, wc), negative_slope=0.2)
        fac = ChunkNormFn.apply(fac)
        return DisenRouteFn.apply(adj, self.iter_k, fac)


class DisenGCN(EvalMixin, nn.Module):
    def __init__(self, data, args=None):
        super().__init__()
        self._config(config.current())
        self.num_list = [data.num['user'], data.num['item'], data.num['tag']]        # disengcn.py:52 (always 3)
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
        self.message_drop_list = cfg['message_drop_list']
        if self.factor_k != 4 or self.dim_latent != 64:
            raise NotImplementedError("the routing kernels are built for factor_k == 4 and dim_latent == 64")

    def _init_weight(self):
        self.embed = nn.ParameterList()
        for num in self.num_list:
            self.embed.append(nn.Parameter(torch.empty(num, self.dim_latent)))
        self.layer = nn.ModuleList()
        for i in range(self.num_layer):
            self.layer.append(Layer(self.factor_k, self.iterate_k, self.dim_latent, self.dim_latent))
        for p in self.parameters():
            nn.init.xavier_uniform_(p)

    def _propagate(self):
        all_emb = torch.cat(list(self.embed), dim=0)
        for i in range(self.num_layer):
            all_emb = self.layer[i].forward(self.norm_adj, all_emb)
            all_emb = F.dropout(all_emb, p=self.message_drop_list[i], training=self.training)
        return all_emb

    def _final_table(self):
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            return self._propagate()
        key = tuple((p.data_ptr(), p._version) for p in self.parameters())
        if self._cache is None or self._cache[0] != key:
            with torch.no_grad():
                self._cache = (key, self._propagate())
        return self._cache[1]

    def forward(self):
        return torch.split(self._final_table(), self.num_list, dim=0)

    def loss(self, batch_data):
        self._cache = None               # a training step follows: the cached inference table goes stale
        data, cor = batch_data
        final = self._final_table()
        return BprLossFn.apply(data, self.num_list[0], self.reg, self.loss_func, final, final)

    def predict_rating(self, users):
        all_users, all_items = self.forward()[:2]
        return torch.sigmoid(torch.matmul(all_users[users], all_items.t()))
