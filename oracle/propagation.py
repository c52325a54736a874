"""Oracle: SpMM propagation, LightGCN / NGCF forward + closed-form backward, BPR loss.

Restates model/help/adj.py:158-167 (split_mm), model/lightgcn.py:49-82, model/ngcf.py:62-105,
model/help/loss.py:4-32.  torch-CPU, dtype chosen by the caller (float32 = reference arithmetic,
float64 = "exact" yardstick both the reference and the CUDA path are measured against).
TEST INFRASTRUCTURE — see oracle/__init__.py.
"""
import numpy as np
import torch
import torch.nn.functional as F

EPS = 1e-12            # F.normalize default eps (lightgcn.py:57, ngcf.py:86)


def spmm(rowptr, col, val, x):
    """adj.py:158-167 split_mm == torch.sparse.mm(A, X): y[r] = sum_j val[j] * x[col[j]], nnz order."""
    n = len(rowptr) - 1
    crow = torch.as_tensor(np.asarray(rowptr), dtype=torch.int64)
    ccol = torch.as_tensor(np.asarray(col), dtype=torch.int64)
    v = torch.as_tensor(np.asarray(val)).to(x.dtype)
    a = torch.sparse_csr_tensor(crow, ccol, v, size=(n, n))
    return a @ x


def spmm_t(rowptr, col, val, g):
    """Transposed product A^T g (SparseAddmmBackward0 of adj.py:166)."""
    n = len(rowptr) - 1
    r = torch.as_tensor(np.repeat(np.arange(n), np.diff(np.asarray(rowptr))), dtype=torch.int64)
    c = torch.as_tensor(np.asarray(col), dtype=torch.int64)
    v = torch.as_tensor(np.asarray(val)).to(g.dtype)
    out = torch.zeros_like(g)
    out.index_add_(0, c, g[r] * v[:, None])
    return out


def row_normalise(e):
    """F.normalize(e, p=2, dim=1): e / max(||e||_2, 1e-12)."""
    nrm = e.norm(dim=1, keepdim=True).clamp_min(EPS)
    return e / nrm, nrm


# ------------------------------------------------------------------------------------------------ LightGCN
def lightgcn_forward(csr, e0, n_layer):
    """lightgcn.py:52-60.  Returns final table F = mean(E0, Y1..YL) and the saved raw layers [E1..EL].
    The UN-normalised E^k propagates; the normalised copy enters the mean."""
    rowptr, col, val = csr
    e, acc, raw = e0, e0.clone(), []
    for _ in range(n_layer):
        e = spmm(rowptr, col, val, e)
        raw.append(e)
        acc = acc + row_normalise(e)[0]
    return acc / (n_layer + 1), raw


def normalise_backward(g, e):
    """d/de of e / max(||e||, eps) applied to g (autograd of F.normalize: clamp passes no grad below eps)."""
    nrm = e.norm(dim=1, keepdim=True)
    big = nrm >= EPS
    n = nrm.clamp_min(EPS)
    y = e / n
    proj = (y * g).sum(1, keepdim=True)
    return torch.where(big, (g - y * proj) / n, g / n)


def lightgcn_backward(csr, raw, g_final, n_layer):
    """Closed form of autograd through lightgcn.py:52-60 (verified vs autograd):
    gY = gF/(L+1);  G_L = nb(gY, E^L);  G_k = nb(gY, E^k) + A^T G_{k+1};  gE0 = gY + A^T G_1."""
    rowptr, col, val = csr
    gy = g_final / (n_layer + 1)
    g = normalise_backward(gy, raw[n_layer - 1])
    for k in range(n_layer - 1, 0, -1):
        g = normalise_backward(gy, raw[k - 1]) + spmm_t(rowptr, col, val, g)
    return gy + spmm_t(rowptr, col, val, g)


# ------------------------------------------------------------------------------------------------ BPR
def bpr_loss(fu, fp, fn, kind):
    """loss.py:4-12 mul_loss."""
    pos = (fu * fp).sum(1)
    neg = (fu * fn).sum(1)
    if kind == "logsigmoid":
        return -F.logsigmoid(pos - neg).mean()
    return F.softplus(neg - pos).mean()


def l2reg(*embs):
    """loss.py:27-32 l2reg_loss = 1/2 * sum ||block||_F^2 / B  (norm(2).pow(2): sqrt then square)."""
    tot = 0
    for e in embs:
        tot = tot + e.norm(2).pow(2)
    return 0.5 * tot / float(embs[0].shape[0])


def bpr_forward_backward(final, ego, batch, n_user, reg, kind, reg_on_final=False):
    """Closed form of loss + gradients of lightgcn.py:68-82 / ngcf.py:95-105 w.r.t. the final table and the
    table the L2 term reads (ego for LightGCN/DGCF, propagated for NGCF/TGCN/DisenGCN).
    batch: (B,3) [u, i+, i-]; item rows live at offset n_user in the concatenated table.
    Returns loss, reg_term (already multiplied by reg), g_final, g_ego (dense N x D)."""
    u = torch.as_tensor(batch[:, 0], dtype=torch.int64)
    p = torch.as_tensor(batch[:, 1], dtype=torch.int64) + n_user
    q = torch.as_tensor(batch[:, 2], dtype=torch.int64) + n_user
    b = len(u)
    fu, fp, fn = final[u], final[p], final[q]
    x = (fu * fn).sum(1) - (fu * fp).sum(1)
    loss = F.softplus(x).mean() if kind != "logsigmoid" else -F.logsigmoid(-x).mean()
    s = (torch.sigmoid(x) / b)[:, None]
    gf = torch.zeros_like(final)
    gf.index_add_(0, u, s * (fn - fp))
    gf.index_add_(0, p, -s * fu)
    gf.index_add_(0, q, s * fu)
    src = final if reg_on_final else ego
    ru, rp, rq = src[u], src[p], src[q]
    reg_term = reg * 0.5 * ((ru * ru).sum() + (rp * rp).sum() + (rq * rq).sum()) / b
    ge = torch.zeros_like(src)
    for idx, r in ((u, ru), (p, rp), (q, rq)):
        ge.index_add_(0, idx, r * (reg / b))
    return loss, reg_term, gf, ge


def lightgcn_loss_and_grad(csr, e0, batch, n_user, n_layer, reg, kind="softplus"):
    """End-to-end LightGCN.loss + backward via the closed forms above.  Returns (loss, reg_term, gE0, F)."""
    final, raw = lightgcn_forward(csr, e0, n_layer)
    loss, reg_term, gf, ge = bpr_forward_backward(final, e0, batch, n_user, reg, kind)
    g0 = lightgcn_backward(csr, raw, gf, n_layer) + ge
    return loss, reg_term, g0, final


# ------------------------------------------------------------------------------------------------ NGCF
def ngcf_forward(csr, e0, mats, n_layer):
    """ngcf.py:73-90 bi_inter_embed.  mats: dict W1_k,b1_k,W2_k,b2_k.  The bias is added to the WEIGHT matrix
    (ngcf.py:78,82).  Output = cat([E0, normalize(E'_1), ...], dim=1)."""
    rowptr, col, val = csr
    e, outs = e0, [e0]
    for k in range(n_layer):
        nb = spmm(rowptr, col, val, e)
        s = F.leaky_relu((nb + e) @ (mats[f"W1_{k}"] + mats[f"b1_{k}"]), 0.2)
        t = F.leaky_relu((nb * e) @ (mats[f"W2_{k}"] + mats[f"b2_{k}"]), 0.2)
        e = s + t
        outs.append(row_normalise(e)[0])
    return torch.cat(outs, dim=1)


def predict_rating(final_user, final_item, users):
    """lightgcn.py:84-89: sigmoid(F_u[users] @ F_i^T)."""
    return torch.sigmoid(final_user[torch.as_tensor(users, dtype=torch.int64)] @ final_item.t())
