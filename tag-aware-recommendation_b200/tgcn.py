"""TGCN — drop-in for model/tgcn.py: same classes (``Attention1``, ``BasicLayer``, ``TGCN``), parameter names, shapes
and creation order (state_dict keys ``embed.user|item|tag|weight``, ``layer.<k>.U|q|p|Wf|bf``,
``layer.<k>.atten1.<type>.W_1|W_2|b|v``, ``layer.<k>.conv.bit_level.weight``, ``layer.<k>.conv.vec_level.conv_<j>
.weight``), ``forward`` / ``loss`` / ``transtag_loss`` / ``predict_rating`` / ``get_ego_embed``.

Neighbour attention (tgcn.py:11-37) — the gather/scatter family that is 45 % of the reference's step — runs on K4
(csrc/nbr_attention.cu) through :class:`NbrAttentionFn`; its three dense projections, the type-level attention
(tgcn.py:78-84) and the 48 vector-level conv features (tgcn.py:92-98) run per node on K7a (csrc/tgcn_mix.cu,
:class:`TgcnMixFn`); the bit-level Conv2d, the concat and the 2096 -> 64 fusion layer (tgcn.py:86-106) run fused on K7
(csrc/tgcn_tail.cu, :class:`TgcnTailFn`).  ``BasicLayer._atten2`` / ``_vec_conv`` keep the torch formulation (tests).  The BPR loss runs on K2 over the 64*(L+1)-d concat
rows (L2 term on the propagated rows, tgcn.py:247), evaluation on K3.
"""
import time

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import config
from ._lib import check, lib, ptr, stream_ptr
from .eval_ops import EvalMixin
from .functional import BprLossFn, skinny_mm


class NbrAttentionFn(torch.autograd.Function):
    """out[v] = sum_k softmax_k(relu(pv[v] + ww[w_vk] + pj[j_vk]) . vvec) * ej[j_vk]   (index 0 = padding)."""

    @staticmethod
    def forward(ctx, pv, ww, pj, ej, vvec, nbr, nbw, k):
        pv, ww, pj, ej, vvec = (t.detach().contiguous() for t in (pv, ww, pj, ej, vvec))
        n = pv.shape[0]
        out = torch.empty((n, ej.shape[1]), dtype=torch.float32, device=pv.device)
        att = torch.empty((n, k), dtype=torch.float32, device=pv.device)
        check(lib().tagrec_nbr_attention_fwd(ptr(pv), ptr(ww), ptr(pj), ptr(ej), ptr(vvec), ptr(nbr), ptr(nbw), n, k,
                                             nbr.stride(0), ej.shape[1], pv.shape[1], ptr(out), ptr(att),
                                             stream_ptr(pv.device)), "tagrec_nbr_attention_fwd")
        ctx.save_for_backward(pv, ww, pj, ej, vvec, nbr, nbw, att)
        ctx.k = k
        return out

    @staticmethod
    def backward(ctx, g_out):
        pv, ww, pj, ej, vvec, nbr, nbw, att = ctx.saved_tensors
        g_out = g_out.contiguous()
        g_pv = torch.empty_like(pv)
        g_ww, g_pj, g_ej, g_v = (torch.zeros_like(t) for t in (ww, pj, ej, vvec))
        check(lib().tagrec_nbr_attention_bwd(ptr(g_out), ptr(att), ptr(pv), ptr(ww), ptr(pj), ptr(ej), ptr(vvec), ptr(nbr),
                                             ptr(nbw), pv.shape[0], ctx.k, nbr.stride(0), ww.shape[0], ej.shape[1],
                                             pv.shape[1], ptr(g_pv), ptr(g_ww), ptr(g_pj), ptr(g_ej), ptr(g_v),
                                             stream_ptr(pv.device)), "tagrec_nbr_attention_bwd")
        return g_pv, g_ww, g_pj, g_ej, g_v, None, None, None


class TgcnTailFn(torch.autograd.Function):
    """out = relu([relu(bit_conv(z)) | xf] Wf + bf) on K7 (csrc/tgcn_tail.cu); the [N, 2096] feature matrix of
    tgcn.py:86-106 is generated and consumed on chip, forward and backward.  ``path``: forward on the tensor cores
    (3xTF32 tcgen05 MMAs, csrc/tgcn_tail_tc.cu; "auto" / "tf32") or on the fp32 FMA kernel ("fp32")."""
    path = "auto"

    @staticmethod
    def forward(ctx, z, wb, xf, wf, bf):
        z, wb, xf, wf, bf = (t.detach().contiguous() for t in (z, wb, xf, wf, bf))
        n, c, e = z.shape[0], wb.shape[0], xf.shape[1]
        out = torch.empty((n, z.shape[2]), dtype=torch.float32, device=z.device)
        nbytes = int(lib().tagrec_tgcn_tail_fwd_workspace_bytes(c))
        ws = torch.empty(nbytes, dtype=torch.uint8, device=z.device)
        check(lib().tagrec_tgcn_tail_fwd_ex(ptr(z), ptr(wb), ptr(xf), ptr(wf), ptr(bf), n, z.shape[2], c, e, ptr(out),
                                            ptr(ws), nbytes, {"auto": 0, "fp32": 1, "tf32": 2}[TgcnTailFn.path],
                                            stream_ptr(z.device)), "tagrec_tgcn_tail_fwd_ex")
        ctx.save_for_backward(z, wb, xf, wf, out)
        return out

    @staticmethod
    def backward(ctx, g_out):
        z, wb, xf, wf, out = ctx.saved_tensors
        n, c, e = z.shape[0], wb.shape[0], xf.shape[1]
        g_out = g_out.contiguous()
        g_z, g_wb, g_xf, g_wf = (torch.empty_like(t) for t in (z, wb, xf, wf))
        g_bf = torch.empty(z.shape[2], dtype=torch.float32, device=z.device)
        nbytes = int(lib().tagrec_tgcn_tail_workspace_bytes(n, c))
        ws = torch.empty(nbytes, dtype=torch.uint8, device=z.device)
        check(lib().tagrec_tgcn_tail_bwd_ex(ptr(g_out), ptr(out), ptr(z), ptr(wb), ptr(xf), ptr(wf), n, z.shape[2], c, e,
                                            ptr(ws), nbytes, ptr(g_z), ptr(g_wb), ptr(g_xf), ptr(g_wf), ptr(g_bf),
                                            {"auto": 0, "fp32": 1, "tf32": 2}[TgcnTailFn.path], stream_ptr(z.device)),
              "tagrec_tgcn_tail_bwd_ex")
        return g_z, g_wb, g_xf, g_wf, g_bf


class TgcnMixFn(torch.autograd.Function):
    """(z, xf) = type-level attention over the (user, item, tag) slots of one node type + rectified vector-level conv
    features, on K7a (csrc/tgcn_mix.cu) — tgcn.py:78-84 and 92-98."""

    @staticmethod
    def forward(ctx, x0, x1, x2, U, q, p, w1, w2, w3):
        x0, x1, x2, U, q, p, w1, w2, w3 = (t.detach().contiguous() for t in (x0, x1, x2, U, q, p, w1, w2, w3))
        n, v = x0.shape[0], w1.shape[0]
        z = torch.empty((n, 3, x0.shape[1]), dtype=torch.float32, device=x0.device)
        xf = torch.empty((n, 6 * v), dtype=torch.float32, device=x0.device)
        check(lib().tagrec_tgcn_mix_fwd(ptr(x0), ptr(x1), ptr(x2), ptr(U), ptr(q), ptr(p), ptr(w1), ptr(w2), ptr(w3), n,
                                        x0.shape[1], U.shape[1], v, ptr(z), ptr(xf), stream_ptr(x0.device)),
              "tagrec_tgcn_mix_fwd")
        ctx.save_for_backward(x0, x1, x2, U, q, p, w1, w2, w3, z, xf)
        return z, xf

    @staticmethod
    def backward(ctx, g_z, g_xf):
        x0, x1, x2, U, q, p, w1, w2, w3, z, xf = ctx.saved_tensors
        n, v = x0.shape[0], w1.shape[0]
        g_z = torch.zeros((n, 3, x0.shape[1]), dtype=torch.float32, device=x0.device) if g_z is None else g_z.contiguous()
        g_xf = torch.zeros_like(xf) if g_xf is None else g_xf.contiguous()
        gx = [torch.empty_like(x0) for _ in range(3)]
        nbytes = int(lib().tagrec_tgcn_mix_workspace_bytes(n))
        ws = torch.empty(nbytes, dtype=torch.uint8, device=x0.device)
        flat = torch.zeros(sum(t.numel() for t in (U, q, p, w1, w2, w3)), dtype=torch.float32, device=x0.device)
        gp = [c.view_as(t) for c, t in zip(flat.split([t.numel() for t in (U, q, p, w1, w2, w3)]), (U, q, p, w1, w2, w3))]
        check(lib().tagrec_tgcn_mix_bwd(ptr(x0), ptr(x1), ptr(x2), ptr(U), ptr(q), ptr(p), ptr(w1), ptr(w2), ptr(w3), n,
                                        x0.shape[1], U.shape[1], v, ptr(z), ptr(g_z), ptr(g_xf), ptr(xf), ptr(gx[0]), ptr(gx[1]),
                                        ptr(gx[2]), *(ptr(t) for t in gp), ptr(ws), nbytes, stream_ptr(x0.device)),
              "tagrec_tgcn_mix_bwd")
        return (*gx, *gp)


class Attention1(nn.Module):
    def __init__(self, in_features: int, atten_dim: int, dim_w: int):
        super().__init__()
        self.in_features = in_features
        self.W_1 = nn.Parameter(torch.empty(in_features + dim_w, atten_dim))
        self.W_2 = nn.Parameter(torch.empty(in_features, atten_dim))
        self.b = nn.Parameter(torch.empty(1, atten_dim))
        self.v = nn.Parameter(torch.empty(1, atten_dim))

    def forward_torch(self, ev, ej, ew, v_jw):
        """tgcn.py:20-37 composed from torch ops — layer widths K4 is not built for (anything but 64-d rows with a
        32-d attention space): index 0 of a table entry is the zero padding row, padding slots take part in the softmax."""
        v_j, v_w = v_jw
        ej0 = torch.cat([ej.new_zeros((1, ej.shape[1])), ej])
        ew0 = torch.cat([ew.new_zeros((1, ew.shape[1])), ew])
        e_nj, e_nw = ej0[v_j], ew0[v_w]
        d = self.in_features
        av = (torch.matmul(ev, self.W_1[:d]) + self.b).unsqueeze(1) + torch.matmul(e_nw, self.W_1[d:]) \
            + torch.matmul(e_nj, self.W_2)
        a = torch.softmax(torch.matmul(F.relu(av), self.v.T), dim=1)
        return torch.sum(a * e_nj, dim=1)

    def forward(self, ev, ej, ew, v_jw, pj=None):
        """tgcn.py:20-37.  ``pj`` = ej @ W_2 may be passed in when two calls share the neighbour type."""
        v_j, v_w = v_jw
        d = self.in_features
        pv = skinny_mm(ev, self.W_1[:d]) + self.b          # [e_v | e_w] W1 + b, split by rows of W1
        ww = torch.matmul(ew, self.W_1[d:])
        if pj is None:
            pj = skinny_mm(ej, self.W_2)
        return NbrAttentionFn.apply(pv, ww, pj, ej, self.v.reshape(-1), v_j, v_w, v_j.shape[1])


class BasicLayer(nn.Module):
    def __init__(self, in_features, out_features, atten_dim, weight_dim, num_bit_conv, num_vector_conv):
        super().__init__()
        self._in_dim = in_features
        self._out_dim = out_features
        self._atten_dim = atten_dim
        self._num_vector_conv = num_vector_conv
        self._num_bit_conv = num_bit_conv
        self.atten1 = nn.ModuleDict()
        for name in ("user", "item", "tag"):
            self.atten1.update({name: Attention1(in_features, atten_dim, weight_dim)})
        self.U = nn.Parameter(torch.empty(in_features, atten_dim))
        self.q = nn.Parameter(torch.empty(1, atten_dim))
        self.p = nn.Parameter(torch.empty(1, atten_dim))
        self.conv = self._conv_layer()
        in_k = self._num_bit_conv * in_features + self._num_vector_conv * (3 + 2 + 1)
        self.Wf = nn.Parameter(torch.empty(in_k, out_features))
        self.bf = nn.Parameter(torch.empty(1, out_features))

    def _conv_layer(self):
        vector_dict = nn.ModuleDict()
        for j in range(1, 4):
            vector_dict.update({f"conv_{j}": nn.Conv2d(1, self._num_vector_conv, kernel_size=(j, self._in_dim), bias=False)})
        return nn.ModuleDict({
            "bit_level": nn.Conv2d(1, self._num_bit_conv, kernel_size=(3, 1), bias=False),
            "vec_level": vector_dict,
        })

    def _atten2(self, u, i, t):
        uit = torch.stack([u, i, t], dim=1)
        x = torch.matmul(uit, self.U) + self.q
        x = torch.matmul(F.relu(x), self.p.T)
        return torch.softmax(x, dim=1) * uit

    def _vec_conv(self, eN):
        """Vector-level branch of tgcn.py:92-98: Conv2d(1 -> 8, (j, 64)), j = 1..3, over a [N, 3, 64] stack — tiny
        contractions, evaluated as fp32 matmuls on the modules' own weights (same values as F.conv2d; cuDNN is free
        to run convolutions, forward AND backward, in TF32 under torch's defaults, which would break fp32 parity of
        every upstream gradient).  Returns the rectified [N, 6 * 8] features, channel-major per conv."""
        n = eN.shape[0]
        vec_e = []
        for j, model in enumerate(self.conv["vec_level"].values(), start=1):
            w = model.weight.reshape(model.weight.shape[0], -1)              # [8, j*64]
            pos = [torch.matmul(eN[:, p:p + j, :].reshape(n, -1), w.t()) for p in range(4 - j)]
            vec_e.append(F.relu(torch.stack(pos, dim=2)).reshape(n, -1))     # [N, 8*(4-j)], channel-major
        return torch.cat(vec_e, dim=-1)

    def _conv_fusion(self, eN):
        """tgcn.py:86-106 (_conv + _fusion) on K7: bit-level conv, concat and the 2096 -> 64 fusion layer fused."""
        wb = self.conv["bit_level"].weight[:, 0, :, 0]                       # [32, 3]
        return TgcnTailFn.apply(eN, wb, self._vec_conv(eN), self.Wf, self.bf.reshape(-1))

    def _cat_tables(self, a, b):
        """Row-wise concatenation of two (neighbour, weight-id) table pairs, cached by storage (the tables are built
        once per model, tgcn.py:194-202)."""
        key = (a[0].data_ptr(), b[0].data_ptr(), a[0].shape, b[0].shape)
        cache = self.__dict__.setdefault("_table_cache", {})
        if key not in cache:
            cache[key] = (torch.cat([a[0], b[0]], 0).contiguous(), torch.cat([a[1], b[1]], 0).contiguous())
        return cache[key]

    def _tail_torch(self, z, xf):
        """tgcn.py:86-91,100-106 from torch ops (fusion widths K7 is not built for): rectified bit-level features
        relu(wb[c, :] . z[v, :, d]) channel-major, the vector-level features appended, then the fusion layer."""
        wb = self.conv["bit_level"].weight[:, 0, :, 0]                       # [C, 3]
        bit = F.relu(torch.einsum("cs,nsd->ncd", wb, z)).reshape(z.shape[0], -1)
        return F.relu(torch.matmul(torch.cat([bit, xf], dim=1), self.Wf) + self.bf)

    def _forward_torch(self, eu, ei, et, ew, u_iw, u_tw, i_uw, i_tw, t_uw, t_iw):
        """The whole layer in the reference's own formulation (tgcn.py:107-137) for input widths other than 64."""
        a_u, a_i, a_t = self.atten1["user"], self.atten1["item"], self.atten1["tag"]
        eu_iN, eu_tN = a_i.forward_torch(eu, ei, ew, u_iw), a_t.forward_torch(eu, et, ew, u_tw)
        ei_uN, ei_tN = a_u.forward_torch(ei, eu, ew, i_uw), a_t.forward_torch(ei, et, ew, i_tw)
        et_uN, et_iN = a_u.forward_torch(et, eu, ew, t_uw), a_i.forward_torch(et, ei, ew, t_iw)
        outs = []
        for trip in ((eu, eu_iN, eu_tN), (ei_uN, ei, ei_tN), (et_uN, et_iN, et)):
            z = self._atten2(*trip)
            outs.append(self._tail_torch(z, self._vec_conv(z)))
        return tuple(outs)

    def forward(self, eu, ei, et, ew, u_iw, u_tw, i_uw, i_tw, t_uw, t_iw):
        if self._in_dim != 64 or self._atten_dim != 32 or u_iw[0].shape[1] > 32 or self._num_vector_conv not in (4, 8) \
                or self._num_bit_conv > 256:
            return self._forward_torch(eu, ei, et, ew, u_iw, u_tw, i_uw, i_tw, t_uw, t_iw)
        a_u, a_i, a_t = self.atten1["user"], self.atten1["item"], self.atten1["tag"]
        pj_u, pj_i, pj_t = skinny_mm(eu, a_u.W_2), skinny_mm(ei, a_i.W_2), skinny_mm(et, a_t.W_2)
        # Each Attention1 module serves two node types (e.g. "item" neighbours of users and of tags): one K4 pass over
        # both (node rows and neighbour tables concatenated; the tables are static, their concatenation is cached).
        nu, ni, nt = eu.shape[0], ei.shape[0], et.shape[0]
        eu_iN, et_iN = torch.split(a_i.forward(torch.cat([eu, et], 0), ei, ew, self._cat_tables(u_iw, t_iw), pj_i), [nu, nt])
        eu_tN, ei_tN = torch.split(a_t.forward(torch.cat([eu, ei], 0), et, ew, self._cat_tables(u_tw, i_tw), pj_t), [nu, ni])
        ei_uN, et_uN = torch.split(a_u.forward(torch.cat([ei, et], 0), eu, ew, self._cat_tables(i_uw, t_uw), pj_u), [ni, nt])
        # The three node types share U/q/p, the convolutions and the fusion layer: one K7a and one K7 pass over their
        # concatenation (the reference's own commented-out variant, tgcn.py:131-137).
        par = (self.U, self.q.reshape(-1), self.p.reshape(-1)) + tuple(
            m.weight.reshape(m.weight.shape[0], -1) for m in self.conv["vec_level"].values())
        z, xf = TgcnMixFn.apply(torch.cat([eu, ei_uN, et_uN], 0), torch.cat([eu_iN, ei, et_iN], 0),
                                torch.cat([eu_tN, ei_tN, et], 0), *par)
        if self._out_dim != 64:          # K7 fuses a 64-wide fusion layer; other output widths: same maths from torch ops
            out = self._tail_torch(z, xf)
        else:
            wb = self.conv["bit_level"].weight[:, 0, :, 0]                   # [32, 3]
            out = TgcnTailFn.apply(z, wb, xf, self.Wf, self.bf.reshape(-1))
        return torch.split(out, [eu.shape[0], ei.shape[0], et.shape[0]], dim=0)


class TGCN(EvalMixin, nn.Module):
    def __init__(self, data):
        super().__init__()
        self._config(config.current())
        self.num_user = data.num['user']
        self.num_item = data.num['item']
        self.num_tag = data.num['tag']
        self.num_weight = data.num['weight']
        self._init_weight()
        self.data = data
        start = time.time()
        self.all_sample = data.get_all_neighbor()
        self._nbr_dev = None
        self._cache = None
        print(f"TGCN got ready! [neighbor sample time {time.time()-start}]")

    def _config(self, cfg):
        self.dim_latent = cfg['dim_latent']
        self.dim_weight = cfg['dim_weight']
        self.num_layer = len(cfg['dim_layer_list'])
        self.dim_layer_list = [self.dim_latent] + list(cfg['dim_layer_list'])
        self.dim_atten = cfg['dim_atten']
        self.num_bit_conv = cfg['num_bit_conv']
        self.num_vec_conv = cfg['num_vec_conv']
        self.message_drop_list = cfg['message_drop_list']
        self.device = cfg['device']
        self.neighbor_k = cfg['neighbor_k']
        self.reg = cfg['reg']
        self.transtag_reg = cfg['transtag_reg']
        self.loss_func = cfg['mul_loss_func']
        self.margin = cfg['margin']
        # The kernels (K4 / K7a / K7) are built for the reference's tgcn overlay (utility/config.py:41-51: 64-d rows,
        # dim_atten 32, neighbor_k 25, 32 bit-level and 8 vector-level channels) with 64-wide layers; every other
        # configuration — e.g. the argparse default dim_layer_list [64, 32, 16] (utility/utils.py:39) — runs the same
        # maths layer by layer from torch ops on the device (BasicLayer._forward_torch / _tail_torch).

    def _init_weight(self):
        self.embed = nn.ParameterDict({
            "user": nn.Parameter(torch.empty(self.num_user, self.dim_latent)),
            "item": nn.Parameter(torch.empty(self.num_item, self.dim_latent)),
            "tag": nn.Parameter(torch.empty(self.num_tag, self.dim_latent)),
            "weight": nn.Parameter(torch.empty(self.num_weight, self.dim_weight)),
        })
        self.layer = nn.ModuleDict()
        for k in range(self.num_layer):
            self.layer.update({f'{k}': BasicLayer(self.dim_layer_list[k], self.dim_layer_list[k + 1], self.dim_atten,
                                                  self.dim_weight, self.num_bit_conv, self.num_vec_conv)})
        # tgcn.py:187-192: Xavier-uniform on every parameter (incl. biases and the Conv2d weights), creation order
        for param in self.parameters():
            nn.init.xavier_uniform_(param)

    def sample(self):
        """tgcn.py:194-202.  The reference shuffles an index vector it never uses (dead code) and takes the FIRST
        ``neighbor_k`` columns; the shuffle still advances numpy's global generator once per relation, which the
        parity-mode samplers read — so the draw is reproduced, the tables are uploaded once."""
        for adj_w in self.all_sample:
            np.random.shuffle(np.arange(int((adj_w[0] if torch.is_tensor(adj_w[0]) else np.asarray(adj_w[0])).shape[1])))
        if self._nbr_dev is None or self._nbr_dev[0][0].device != self.embed["user"].device:
            dev = self.embed["user"].device

            def up(x):       # host tables (numpy, the reference's contract) or device tables (get_all_neighbor_device)
                if torch.is_tensor(x):
                    return x[:, :self.neighbor_k].to(device=dev, dtype=torch.long).contiguous()
                return torch.as_tensor(np.ascontiguousarray(np.asarray(x)[:, :self.neighbor_k]), dtype=torch.long, device=dev)
            self._nbr_dev = [tuple(up(x) for x in adj_w) for adj_w in self.all_sample]
        return self._nbr_dev

    def _propagate(self):
        eu, ei = self.embed['user'], self.embed['item']
        et, ew = self.embed['tag'], self.embed['weight']
        embs_u, embs_i, embs_t = [eu], [ei], [et]
        for i, layer in enumerate(self.layer.values()):
            u_iw, u_tw, i_uw, i_tw, t_uw, t_iw = self.sample()
            eu, ei, et = layer(eu, ei, et, ew, u_iw, u_tw, i_uw, i_tw, t_uw, t_iw)
            eu = F.dropout(eu, p=self.message_drop_list[i], training=self.training)
            ei = F.dropout(ei, p=self.message_drop_list[i], training=self.training)
            et = F.dropout(et, p=self.message_drop_list[i], training=self.training)
            embs_u.append(F.normalize(eu, p=2, dim=1))
            embs_i.append(F.normalize(ei, p=2, dim=1))
            embs_t.append(F.normalize(et, p=2, dim=1))
        return torch.cat(embs_u, dim=1), torch.cat(embs_i, dim=1), torch.cat(embs_t, dim=1)

    def forward(self):
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            return self._propagate()
        key = tuple((p.data_ptr(), p._version) for p in self.parameters())
        if self._cache is None or self._cache[0] != key:
            with torch.no_grad():
                self._cache = (key, self._propagate())
        return self._cache[1]

    def get_ego_embed(self):
        return self.embed['user'], self.embed['item'], self.embed['tag']

    def loss(self, batch_data):
        self._cache = None               # a training step follows: the cached inference table goes stale
        all_users, all_items = self.forward()[:2]
        final = torch.cat([all_users, all_items], dim=0)
        return BprLossFn.apply(batch_data, self.num_user, self.reg, self.loss_func, final, final)

    def transtag_loss(self, batch_data):
        """tgcn.py:251-261 + model/help/loss.py:35-41 — ego rows only, tiny; plain torch ops."""
        user, tag, pos_item, neg_item = batch_data.T
        all_users, all_items, all_tags = self.get_ego_embed()
        tag_emb, user_emb = all_tags[tag.long()], all_users[user.long()]
        pos_i_emb, neg_i_emb = all_items[pos_item.long()], all_items[neg_item.long()]
        pos_score = torch.norm(user_emb + tag_emb - pos_i_emb, p=2, dim=1)
        neg_score = torch.norm(user_emb + tag_emb - neg_i_emb, p=2, dim=1)
        loss = torch.mean(torch.relu(self.margin + pos_score - neg_score))
        reg = 0
        for emb in (user_emb, tag_emb, pos_i_emb, neg_i_emb):
            reg = reg + emb.norm(2).pow(2)
        reg_loss = 0.5 * reg / float(user_emb.shape[0])
        return loss, self.transtag_reg * reg_loss

    def predict_rating(self, users):
        all_users, all_items = self.forward()[:2]
        return torch.sigmoid(torch.matmul(all_users[users], all_items.t()))
