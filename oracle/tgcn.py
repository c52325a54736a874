"""Oracle: TGCN forward (torch-CPU, autograd for gradients) in the reference's own formulation.

Restates model/tgcn.py:11-37 (Attention1), :40-137 (BasicLayer: type-level attention, bit-/vector-level Conv2d,
fusion) and :204-233 (TGCN.forward: first ``neighbor_k`` columns of the padded tables, message dropout off,
row-normalised layer outputs concatenated).  Parameters are passed as a flat dict keyed like the reference's
state_dict.  TEST INFRASTRUCTURE — see oracle/__init__.py.
"""
import torch
import torch.nn.functional as F


def attention1(P, prefix, ev, ej, ew, v_j, v_w):
    """tgcn.py:20-37: index 0 of a table entry = padding (zero row), padding slots take part in the softmax."""
    ej = torch.cat([torch.zeros(1, ej.shape[1], dtype=ej.dtype), ej])
    ew = torch.cat([torch.zeros(1, ew.shape[1], dtype=ew.dtype), ew])
    k = v_j.shape[1]
    e_nj, e_nw = ej[v_j], ew[v_w]
    e_nv = ev.unsqueeze(1).repeat(1, k, 1)
    av = torch.cat([e_nv, e_nw], dim=-1) @ P[prefix + "W_1"] + e_nj @ P[prefix + "W_2"] + P[prefix + "b"]
    a = torch.softmax(torch.relu(av) @ P[prefix + "v"].T, dim=1)
    return (a * e_nj).sum(1)


def basic_layer(P, pre, eu, ei, et, ew, tables):
    """tgcn.py:107-137.  tables = (u_iw, u_tw, i_uw, i_tw, t_uw, t_iw), each (ids, weights) int64 [n, k]."""
    u_iw, u_tw, i_uw, i_tw, t_uw, t_iw = tables
    att = lambda kind, ev, ej, tb: attention1(P, f"{pre}atten1.{kind}.", ev, ej, ew, tb[0], tb[1])   # noqa: E731
    eu_iN, eu_tN = att("item", eu, ei, u_iw), att("tag", eu, et, u_tw)
    ei_uN, ei_tN = att("user", ei, eu, i_uw), att("tag", ei, et, i_tw)
    et_uN, et_iN = att("user", et, eu, t_uw), att("item", et, ei, t_iw)

    def atten2(u, i, t):                                                      # tgcn.py:78-84
        uit = torch.stack([u, i, t], dim=1)
        x = torch.relu(uit @ P[pre + "U"] + P[pre + "q"]) @ P[pre + "p"].T
        return torch.softmax(x, dim=1) * uit

    def conv(eN):                                                             # tgcn.py:86-101
        x = eN.unsqueeze(1)
        bit = torch.relu(F.conv2d(x, P[pre + "conv.bit_level.weight"]))
        bit = bit.reshape(bit.shape[0], -1)
        vec = []
        for j in (1, 2, 3):
            y = torch.relu(F.conv2d(x, P[pre + f"conv.vec_level.conv_{j}.weight"])).squeeze(-1)
            vec.append(y.reshape(y.shape[0], -1))
        return torch.cat([bit, torch.cat(vec, dim=-1)], dim=1)

    fusion = lambda x: torch.relu(x @ P[pre + "Wf"] + P[pre + "bf"])          # noqa: E731  tgcn.py:103-106
    return (fusion(conv(atten2(eu, eu_iN, eu_tN))), fusion(conv(atten2(ei_uN, ei, ei_tN))),
            fusion(conv(atten2(et_uN, et_iN, et))))


def tgcn_forward(P, tables, n_layer, neighbor_k):
    """tgcn.py:204-233.  tables: the 6 (ids, weights) pairs of data.get_all_neighbor()."""
    eu, ei, et, ew = P["embed.user"], P["embed.item"], P["embed.tag"], P["embed.weight"]
    tb = tuple((torch.as_tensor(a)[:, :neighbor_k].long(), torch.as_tensor(w)[:, :neighbor_k].long()) for a, w in tables)
    outs = [[eu], [ei], [et]]
    for k in range(n_layer):
        eu, ei, et = basic_layer(P, f"layer.{k}.", eu, ei, et, ew, tb)
        for lst, e in zip(outs, (eu, ei, et)):
            lst.append(F.normalize(e, p=2, dim=1))
    return tuple(torch.cat(x, dim=1) for x in outs)
