"""Oracle: DGCF intent-aware routing propagation and DisenGCN neighbour routing (torch-CPU, autograd for gradients).

Restates model/dgcf.py:49-110 (forward / iterate_update / factor_update) and model/disengcn.py:23-46 (Layer.forward),
model/disengcn.py:86-99 (DisenGCN.forward) on an explicit edge list instead of per-factor torch sparse tensors.
No gradient flows through the routing weights (dgcf.py:92 / disengcn.py:36 `.detach()`), which is reproduced with
``.detach()`` here.  dtype chosen by the caller (float32 = reference arithmetic, float64 = yardstick).
TEST INFRASTRUCTURE — see oracle/__init__.py.
"""
import torch
import torch.nn.functional as F


def _seg_sum(values, index, n):
    return torch.zeros((n,) + values.shape[1:], dtype=values.dtype).index_add_(0, index, values)


def dgcf_forward(head, tail, n, e0, n_layer, iterate_k, factor_k=4):
    """dgcf.py:49-65.  head/tail: int64 [nnz] endpoints of every stored entry of the ('plain') adjacency
    (norm_adj._indices(), dgcf.py:103); e0: [n, 64] = cat(embed).  Returns the mean table [n, 64]."""
    dk = e0.shape[1] // factor_k
    a_values = torch.ones(factor_k, head.numel(), dtype=e0.dtype)            # dgcf.py:50
    ego, layers = e0, [e0]
    for _ in range(n_layer):
        split = torch.split(ego, dk, dim=1)
        layer_emb = []
        for t in range(iterate_k):                                           # dgcf.py:71
            a_factor = torch.softmax(a_values, dim=0)                        # dgcf.py:74
            scores = []
            for i in range(factor_k):
                w = a_factor[i].detach()                                     # dgcf.py:92
                rowsum = _seg_sum(w, head, n)                                # dgcf.py:94 sparse.sum(adj, dim=1)
                d = torch.where(rowsum > 0, 1.0 / torch.sqrt(rowsum), torch.zeros_like(rowsum))   # dgcf.py:95-97
                f = d[:, None] * _seg_sum(w[:, None] * (d[:, None] * split[i])[tail], head, n)    # dgcf.py:99-101
                h_emb = F.normalize(f[head], p=2, dim=1)                     # dgcf.py:104,106
                t_emb = F.normalize(split[i][tail], p=2, dim=1)              # dgcf.py:105,107
                scores.append((h_emb * torch.tanh(t_emb)).sum(1))            # dgcf.py:108-109
                if t == iterate_k - 1:
                    layer_emb.append(f)
            a_values = a_values + torch.stack(scores, dim=0)                 # dgcf.py:82-83
        ego = torch.cat([F.normalize(x, p=2, dim=1) for x in layer_emb], dim=1)   # dgcf.py:85-87
        layers.append(ego)
    return torch.stack(layers, dim=1).mean(dim=1)                            # dgcf.py:59-60


def disengcn_layer(head, tail, n, all_emb, W, b, iterate_k):
    """disengcn.py:23-46.  W [K, in, dk], b [K, 1, dk] — bias added to the weight (disengcn.py:24)."""
    fac = torch.matmul(all_emb, W + b)                                       # [K, n, dk]
    fac = F.normalize(F.leaky_relu(fac, 0.2), p=2, dim=2)                    # disengcn.py:25-26
    new = fac
    for _ in range(iterate_k):
        p = (new[:, head] * fac[:, tail]).sum(2)                             # disengcn.py:31-33
        p = torch.softmax(p, dim=0).detach()                                 # disengcn.py:34,36
        out = []
        for i in range(W.shape[0]):
            emb = fac[i] + _seg_sum(p[i][:, None] * fac[i][tail], head, n)   # disengcn.py:39-40
            out.append(F.normalize(emb, p=2, dim=1))                         # disengcn.py:41
        new = torch.stack(out)
    return torch.cat(list(new), dim=1)                                       # disengcn.py:45


def disengcn_forward(head, tail, n, e0, weights, iterate_k):
    """disengcn.py:86-99: the LAST layer's output only.  weights: list of (W, b) per layer."""
    x = e0
    for W, b in weights:
        x = disengcn_layer(head, tail, n, x, W, b, iterate_k)
    return x
