"""Oracle: block adjacency + normalisation -> CSR  (restates model/help/adj.py:7-150).

numpy only (no scipy): integer structure + float32 values, bit-exact target.
TEST INFRASTRUCTURE — see oracle/__init__.py.
"""
import numpy as np


def directed_edges(n_user, n_item, ui, n_tag=0, ut=None, it=None):
    """Directed (row, col, weight) list of the block matrix.

    adj.py:7-16  create_ui_adj : [[0, R], [R^T, 0]]
    adj.py:19-35 create_uit_adj: 3x3 blocks  U-I, U-T, I-T and their transposes.
    ``ui``/``ut``/``it`` are (rows, cols) index pairs; one unit of weight per listed pair
    (data/utils.py:50-53 to_sparse_adj: val=ones; the COO->LIL conversion sums duplicates).
    """
    rows, cols = [], []

    def block(r, c, roff, coff):
        r = np.asarray(r, dtype=np.int64) + roff
        c = np.asarray(c, dtype=np.int64) + coff
        rows.extend([r, c])
        cols.extend([c, r])

    block(ui[0], ui[1], 0, n_user)
    if ut is not None:
        block(ut[0], ut[1], 0, n_user + n_item)
        block(it[0], it[1], n_user, n_user + n_item)
    n = n_user + n_item + (n_tag if ut is not None else 0)
    return n, np.concatenate(rows), np.concatenate(cols)


def coo_to_csr_sum(n, row, col):
    """Row-major, ascending columns, duplicates summed (what lil_matrix slice-assign + tocsr() yields,
    adj.py:12-15; verified row-major/sorted for every norm_type)."""
    key = row * n + col
    uk, cnt = np.unique(key, return_counts=True)
    r = uk // n
    c = uk % n
    a = cnt.astype(np.float32)
    rowptr = np.zeros(n + 1, dtype=np.int64)
    np.add.at(rowptr, r + 1, 1)
    rowptr = np.cumsum(rowptr)
    return rowptr, c.astype(np.int64), a


def row_ids(rowptr):
    return np.repeat(np.arange(len(rowptr) - 1, dtype=np.int64), np.diff(rowptr))


def _insert_diagonal(n, rowptr, col, a, diag_val=1.0):
    """A + I with sorted columns (adj.py:81,83: sp.eye added to a matrix without self loops)."""
    r = row_ids(rowptr)
    key = np.concatenate([r * n + col, np.arange(n, dtype=np.int64) * (n + 1)])
    v = np.concatenate([a, np.full(n, diag_val, dtype=np.float32)])
    order = np.argsort(key, kind="stable")
    key, v = key[order], v[order]
    assert len(np.unique(key)) == len(key), "adjacency already has self loops"
    r2, c2 = key // n, key % n
    ptr = np.zeros(n + 1, dtype=np.int64)
    np.add.at(ptr, r2 + 1, 1)
    return np.cumsum(ptr), c2, v.astype(np.float32)


def rowsum_f32(n, rowptr, a):
    """adj.py:92,103  np.array(adj.sum(1)) on float32 CSR — integer-valued, exact in fp32 below 2^24."""
    s = np.zeros(n, dtype=np.float32)
    np.add.at(s, row_ids(rowptr), a)
    return s


def inv_pow(rowsum, p):
    """adj.py:93-94 / 105-106: np.power(float32 rowsum, p) with inf -> 0.  numpy's float32 pow is what
    the reference uses, so the oracle (and the product's host side) must call the very same function."""
    with np.errstate(divide="ignore"):
        d = np.power(rowsum.astype(np.float32), p).astype(np.float32).flatten()
    d[np.isinf(d)] = 0.0
    return d


def normalise(n, rowptr, col, a, norm_type):
    """adj.py:75-110 get_norm_adj.  Returns (rowptr int64, col int64, val float32)."""
    if norm_type == "bi_norm":                                   # adj.py:90-98  (D^-1/2 A) D^-1/2
        d = inv_pow(rowsum_f32(n, rowptr, a), -0.5)
        r = row_ids(rowptr)
        val = ((d[r] * a).astype(np.float32) * d[col]).astype(np.float32)
        return rowptr, col, val
    if norm_type == "si_norm":                                   # adj.py:101-110  D^-1 A
        d = inv_pow(rowsum_f32(n, rowptr, a), -1)
        return rowptr, col, (d[row_ids(rowptr)] * a).astype(np.float32)
    if norm_type == "si_norm_self":                              # adj.py:80-81  si_norm(A + I)
        rowptr, col, a = _insert_diagonal(n, rowptr, col, a)
        d = inv_pow(rowsum_f32(n, rowptr, a), -1)
        return rowptr, col, (d[row_ids(rowptr)] * a).astype(np.float32)
    if norm_type == "ngcf":                                      # adj.py:82-83  si_norm(A) + I
        d = inv_pow(rowsum_f32(n, rowptr, a), -1)
        val = (d[row_ids(rowptr)] * a).astype(np.float32)
        return _insert_diagonal(n, rowptr, col, val)
    return rowptr, col, a                                        # adj.py:84-85  "plain"


def creat_adj(n_user, n_item, ui, norm_type, n_tag=0, ut=None, it=None):
    """adj.py:38-46 creat_adj -> CSR triplet of the normalised N x N matrix."""
    n, row, col = directed_edges(n_user, n_item, ui, n_tag, ut, it)
    rowptr, c, a = coo_to_csr_sum(n, row, col)
    return (n,) + normalise(n, rowptr, c, a, norm_type)


def fold_rows(n, k):
    """adj.py:114-130 split_sp_mat: k row slabs of n//k rows, last takes the remainder."""
    if k < 2:
        return [(0, n)]
    f = n // k
    return [(i * f, n if i == k - 1 else (i + 1) * f) for i in range(k)]
