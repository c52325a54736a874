"""LightGCN — drop-in for model/lightgcn.py (same constructor, parameters, state_dict keys, forward / loss /
predict_rating / get_ego_embed), running on K1/K2/K3 of libtagrec_b200.so."""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import adj as utils
from . import config
from .eval_ops import EvalMixin
from .functional import LightGCNLossFn, LightGCNPropagateFn, lightgcn_forward_layers


class LightGCN(EvalMixin, nn.Module):
    def __init__(self, data, args=None):
        super().__init__()
        self._config(config.current())
        if self.use_tag:
            self.num_list = [data.num['user'], data.num['item'], data.num['tag']]
        else:
            self.num_list = [data.num['user'], data.num['item']]
        # lightgcn.py:20 — a plain attribute, not a buffer: not part of state_dict
        self.norm_adj = getattr(data, "prebuilt_adj", None) or \
            utils.creat_adj(data, self.use_tag, self.norm_type, self.split_adj_k, self.device)
        self._ws = {}
        self._cache = None
        self._init_weight()

    def _config(self, cfg):
        self.dim_latent = cfg['dim_latent']
        self.num_layer = len(cfg['dim_layer_list'])
        self.device = cfg['device']
        self.norm_type = cfg['norm_type']
        self.split_adj_k = cfg["split_adj_k"]
        self.reg = cfg['reg']
        self.loss_func = cfg['mul_loss_func']
        self.use_tag = cfg['use_tag']
        self.message_drop_list = cfg['message_drop_list']
        self.node_drop = cfg['node_drop']
        self.init_device = cfg.get('init_device', 'cpu')

    def _init_weight(self):
        # lightgcn.py:39-47: one Parameter per node type, Xavier-uniform in creation order (CPU RNG by default, so
        # torch.manual_seed reproduces the reference's initial weights)
        self.embed = nn.ParameterList()
        for num in self.num_list:
            self.embed.append(nn.Parameter(torch.empty(num, self.dim_latent, device=self.init_device)))
        for p in self.parameters():
            nn.init.xavier_uniform_(p)

    # ------------------------------------------------------------------------------------------------
    def _flat_params(self):
        """One [N, dim] table whose row blocks ARE the parameters (``embed[k].data`` are views of it), so the
        kernels read E0 in place instead of torch.cat-ing the ParameterList every step (lightgcn.py:52).  Rebuilt
        whenever the parameters were re-allocated behind our back (``.to()``, ``load_state_dict(assign=True)``)."""
        flat = self._ws.get("flat")
        off, ok = 0, flat is not None
        if ok:
            for p in self.embed:
                ok = ok and p.device == flat.device and p.data_ptr() == flat[off:off + p.shape[0]].data_ptr()
                off += p.shape[0]
        if not ok:
            flat = torch.cat([p.detach() for p in self.embed], dim=0)
            off = 0
            for p in self.embed:
                p.data = flat[off:off + p.shape[0]]
                off += p.shape[0]
            self._ws["flat"] = flat
        return flat

    def _dropout_active(self):
        return self.training and (self.node_drop > 0 or any(p > 0 for p in self.message_drop_list[:self.num_layer]))

    def _forward_unfused(self):
        """lightgcn.py:49-62 composed from the SpMM primitive — only used when dropout is switched on."""
        norm_adj = utils.node_drop(self.norm_adj, self.node_drop, self.training)
        all_embed = torch.cat(list(self.embed), dim=0)
        acc = all_embed
        for k in range(self.num_layer):
            all_embed = utils.split_mm(norm_adj, all_embed)
            all_embed = F.dropout(all_embed, p=self.message_drop_list[k], training=self.training)
            acc = acc + F.normalize(all_embed, p=2, dim=1)
        return acc / (self.num_layer + 1)

    def _final_table(self):
        if self._dropout_active():
            return self._forward_unfused()
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.embed):
            return LightGCNPropagateFn.apply(self, *self.embed)
        # inference: propagate once per parameter version (the reference re-propagates for every user batch,
        # lightgcn.py:85; in eval() the result is identical — SURVEY A11)
        key = tuple((p.data_ptr(), p._version) for p in self.embed)
        if self._cache is None or self._cache[0] != key:
            e0 = torch.cat([p.detach() for p in self.embed], dim=0)
            raw = [torch.empty_like(e0), torch.empty_like(e0)]
            final = torch.empty_like(e0)
            lightgcn_forward_layers(self.norm_adj, e0, self.num_layer, [raw[k % 2] for k in range(self.num_layer)], final)
            self._cache = (key, final)
        return self._cache[1]

    def forward(self):
        return torch.split(self._final_table(), self.num_list, dim=0)

    def get_ego_embed(self):
        return list(self.embed)

    def loss(self, batch_data):
        self._cache = None               # a training step follows: the cached inference table goes stale
        if self._dropout_active():
            from .functional import BprLossFn
            final = self._final_table()
            ego = torch.cat(list(self.embed), dim=0)
            return BprLossFn.apply(batch_data, self.num_list[0], self.reg, self.loss_func, final, ego)
        return LightGCNLossFn.apply(self, batch_data, *self.embed)

    def predict_rating(self, users):
        """lightgcn.py:84-89 — kept for callers that want the dense (B, n_item) matrix; the evaluation loop of this
        package uses :meth:`eval_topk` instead and never materialises it."""
        all_users, all_items = self.forward()[:2]
        return torch.sigmoid(torch.matmul(all_users[users], all_items.t()))
