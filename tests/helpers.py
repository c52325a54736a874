"""Shared helpers for the parity tests (dataset views over the golden fixtures)."""
import numpy as np


def blocks(g):
    """(ui, ut, it) index pairs as the reference's loaders hand them to creat_adj."""
    e = g["edge_index_train"]
    uit = g["uit_data"].astype(np.int64)
    return (e[:, 0], e[:, 1]), (uit[:, 0], uit[:, 2]), (uit[:, 1], uit[:, 2])


def nums(g):
    u, i, t, w = (int(x) for x in g["num"])
    return u, i, t, w


def user_lists(g, part):
    ptr, items = g[f"{part}_ptr"], g[f"{part}_items"]
    return {int(u): [int(x) for x in items[ptr[u]:ptr[u + 1]]] for u in g[f"{part}_keys"]}


def relerr(a, b):
    """Norm-wise relative error max|a-b| / max|b| (the tolerance north_star states is relative to tensor scale)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    den = np.abs(b).max()
    return float(np.abs(a - b).max() / (den if den > 0 else 1.0))


def coalesced(n, row, col, val):
    """(row, col)-sorted triplets == what torch's COO coalesce() hands to the CPU/CUDA SpMM."""
    order = np.argsort(row.astype(np.int64) * n + col, kind="stable")
    return row[order], col[order], val[order]
