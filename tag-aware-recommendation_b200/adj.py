"""Adjacency build -> device CSR + the SpMM primitive (drop-in for model/help/adj.py).

``creat_adj(data, use_tag, norm_type, split_adj_k, device)`` keeps the reference's signature (adj.py:38-46) but
returns a :class:`CsrGraph` (rowptr int64 / col int32 / val fp32 resident in HBM, built by K0 on the device)
instead of an un-coalesced torch COO that is re-sorted on every multiply.  ``split_mm(graph, E)`` (adj.py:158-167)
is the K1 SpMM kernel.  The only host arithmetic is ``np.power(degree, p)`` — numpy's float32 pow, the very call
the reference makes (adj.py:93,105), needed for bit-exact values.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import CsrDesc, check, lib, ptr, stream_ptr

_NORM = {"bi_norm": (0, 0, -0.5), "si_norm": (1, 0, -1), "si_norm_self": (1, 1, -1), "ngcf": (1, 2, -1)}


class CsrGraph:
    """Normalised N x N adjacency in CSR on one device, plus the long-row plan K1 needs."""

    def __init__(self, n, rowptr, col, val, val_t, weight, norm_type, num_list, row_offset=0, comm=None):
        self.n = int(n)                                    # rows of the tables it multiplies (global node count)
        self.n_rows = int(rowptr.numel()) - 1              # rows stored here (== n unless this is a rank's row block)
        self.row_offset = int(row_offset)
        self.comm = comm                                   # distributed.RowComm when the rows are sharded
        self.rowptr, self.col, self.val = rowptr, col, val
        self.val_t = val_t if val_t is not None else val       # values of A^T (same tensor when A is symmetric)
        self.weight = weight
        self.norm_type = norm_type
        self.num_list = list(num_list)
        self.shape = (self.n, self.n)
        self.device = rowptr.device
        self._build_plan()
        # the masked backward launch (source rows mostly all-zero: their gathers are skipped) is bound by instruction
        # issue, not by DRAM — it keeps the PLAIN plan; a second, plain plan is built only when the main one is blocked
        self._plain_plan = None
        if self.col_block is not None:
            main = self._plan_state()
            self._build_plan(allow_block=False)
            self._plain_plan = self._plan_state()
            self._set_plan_state(main)

    # --- torch-sparse-like accessors the reference's models use (dgcf.py:50, disengcn.py:27) ---
    def _nnz(self):
        return int(self.col.numel())

    def row_ids(self):
        deg = self.rowptr[1:] - self.rowptr[:-1]
        return torch.repeat_interleave(torch.arange(self.n_rows, device=self.device) + self.row_offset, deg)

    def _indices(self):
        return torch.stack([self.row_ids(), self.col.long()])

    def _values(self):
        return self.val

    # ---- the launch plan of K1 -------------------------------------------------------------------------------
    COLBLOCK_MIN_TABLE_BYTES = 1 << 30      # block only rows that gather from a table far larger than the 126 MB L2

    def _column_block_spec(self):
        """(first local row, min degree, window rows) of the column-blocked region, or None.

        Rows of the non-user node types (items, tags) gather rows of the USER table; when that table is far larger than
        the L2 (1 B-edge graph: 2.56 GB) every such gather is a DRAM access — the item-row half of a layer then runs
        at the HBM roofline of a gather formulation (profiles/r1_spmm_l2_probe.md).  The plan cuts those rows at
        boundaries of L2-sized column windows and orders the pieces window-major, so that the pieces in flight at any
        time gather from ONE window of the table and each of its rows is fetched from DRAM once per launch instead of
        once per use.  TAGREC_COLBLOCK=0 switches it off; _MB / _MIN_DEG tune window and the shortest blocked row."""
        import os
        if os.environ.get("TAGREC_COLBLOCK", "1") == "0" or len(self.num_list) < 2:
            return None
        n_user = int(self.num_list[0])
        force = os.environ.get("TAGREC_COLBLOCK_FORCE") == "1"           # tests: block graphs of any size
        if n_user * 256 < self.COLBLOCK_MIN_TABLE_BYTES and not force:
            return None
        begin = min(max(n_user - self.row_offset, 0), self.n_rows)
        if begin >= self.n_rows:
            return None
        window = int(float(os.environ.get("TAGREC_COLBLOCK_MB", "96")) * (1 << 20)) // 256
        min_deg = int(os.environ.get("TAGREC_COLBLOCK_MIN_DEG", "384"))
        return begin, min_deg, max(window, 64)

    _PLAN_KEYS = ("long_row", "long_chunk", "col_block", "blocked_row_begin", "blocked_min_deg", "chunk_lanes",
                  "long_nchunks", "_scratch", "n_long", "long_rows", "item_slot", "item_begin", "item_end", "n_items")

    def _plan_state(self):
        return {k: getattr(self, k) for k in self._PLAN_KEYS}

    def _set_plan_state(self, st):
        for k, v in st.items():
            setattr(self, k, v)

    def _build_plan(self, allow_block=True):
        """Rows above ``long_row`` nnz are cut into ``long_chunk`` pieces (see csrc/spmm.cu).  The thresholds are the
        tuned 4096 / 2048 on big graphs; a small graph is only a few waves of rows, where one sub-warp walking a
        2000-entry hub row IS the launch time, so it is planned with 256 / 256.  Rows of the column-blocked region
        (``_column_block_spec``) are cut at column-window boundaries instead and listed window-major."""
        small = self._nnz() < _lib.SMALL_GRAPH_NNZ
        self.long_row = _lib.SMALL_LONG_ROW if small else _lib.LONG_ROW
        self.long_chunk = _lib.SMALL_LONG_CHUNK if small else _lib.LONG_CHUNK
        LONG_ROW, LONG_CHUNK = self.long_row, self.long_chunk
        dev = self.device
        deg = self.rowptr[1:] - self.rowptr[:-1]
        import os
        spec = None if (not allow_block or (small and os.environ.get("TAGREC_COLBLOCK_FORCE") != "1")) \
            else self._column_block_spec()
        self.col_block = None
        self.blocked_row_begin, self.blocked_min_deg, self.chunk_lanes = 0, 0, 0
        self.long_nchunks = None
        self._scratch = {}
        is_long = deg > LONG_ROW
        if spec is not None:
            begin_row, min_deg, window = spec
            blocked = torch.zeros_like(is_long)
            blocked[begin_row:] = deg[begin_row:] > min_deg
            is_long = is_long | blocked
        long_rows = torch.nonzero(is_long).flatten()
        self.n_long = int(long_rows.numel())
        if self.n_long == 0:
            self.long_rows = self.item_slot = self.item_begin = self.item_end = None
            self.n_items = 0
            return
        slot_of = torch.arange(self.n_long, device=dev)
        slots, begins, ends = [], [], []
        counts = torch.zeros(self.n_long, dtype=torch.int64, device=dev)
        if spec is not None:
            sel = blocked[long_rows]                                  # long rows that are column-blocked
            b_slots = slot_of[sel]
            b_rows = long_rows[sel].to(torch.int32).contiguous()
            nb = int(b_rows.numel())
            if nb:
                n_win = (self.n + window - 1) // window
                bounds = torch.empty((n_win + 1, nb), dtype=torch.int64, device=dev)
                check(lib().tagrec_csr_window_bounds(ptr(self.rowptr), ptr(self.col), ptr(b_rows), nb, window, n_win,
                                                     ptr(bounds), stream_ptr(dev)), "tagrec_csr_window_bounds")
                seg_b, seg_e = bounds[:-1].reshape(-1), bounds[1:].reshape(-1)       # window-major [n_win * nb]
                pieces = (seg_e - seg_b + LONG_CHUNK - 1) // LONG_CHUNK
                counts[b_slots] = pieces.reshape(n_win, nb).sum(0)
                keep = torch.nonzero(pieces > 0).flatten()
                seg_b, seg_e, pieces = seg_b[keep], seg_e[keep], pieces[keep]
                seg_slot = b_slots[keep % nb]
                del keep, bounds
                if int(pieces.max()) == 1:
                    slots.append(seg_slot); begins.append(seg_b); ends.append(seg_e)
                else:
                    rep = torch.repeat_interleave(torch.arange(pieces.numel(), device=dev), pieces)
                    first = torch.cumsum(pieces, 0) - pieces
                    k = torch.arange(rep.numel(), device=dev) - first[rep]
                    pb = seg_b[rep] + k * LONG_CHUNK
                    slots.append(seg_slot[rep]); begins.append(pb); ends.append(torch.minimum(pb + LONG_CHUNK, seg_e[rep]))
                    del rep, first, k, pb
                self.col_block = {"rows": nb, "window_rows": window, "windows": int(n_win), "min_deg": min_deg,
                                  "pieces": int(sum(x.numel() for x in slots))}
                self.blocked_row_begin, self.blocked_min_deg, self.chunk_lanes = int(begin_row), int(min_deg), 1
            plain = ~sel
        else:
            plain = torch.ones(self.n_long, dtype=torch.bool, device=dev)
        p_slots = slot_of[plain]
        if p_slots.numel():                                           # plain long rows: equal-count pieces
            rows_p = long_rows[plain]
            nchunks = (deg[rows_p] + LONG_CHUNK - 1) // LONG_CHUNK
            counts[p_slots] = nchunks
            rep = torch.repeat_interleave(torch.arange(p_slots.numel(), device=dev), nchunks)
            first = torch.cumsum(nchunks, 0) - nchunks
            k = torch.arange(rep.numel(), device=dev) - first[rep]
            pb = self.rowptr[rows_p][rep] + k * LONG_CHUNK
            slots.append(p_slots[rep]); begins.append(pb)
            ends.append(torch.minimum(pb + LONG_CHUNK, self.rowptr[rows_p + 1][rep]))
        self.long_rows = long_rows.to(torch.int32)
        self.item_slot = torch.cat(slots).to(torch.int32).contiguous()
        self.item_begin, self.item_end = torch.cat(begins).contiguous(), torch.cat(ends).contiguous()
        self.n_items = int(self.item_slot.numel())
        self.long_nchunks = counts.to(torch.int32).contiguous()

    def desc(self, dim, transposed=False, plain=False):
        """tagrec_csr_t for a launch at feature width ``dim`` (scratch rows are dim floats wide).  ``plain``: the plan
        without column blocking (used by the masked backward launch)."""
        if plain and self._plain_plan is not None:
            main = self._plan_state()
            self._set_plan_state(self._plain_plan)
            try:
                return self.desc(dim, transposed)
            finally:
                self._plain_plan = self._plan_state()        # keeps the scratch tables created on first use
                self._set_plan_state(main)
        d = CsrDesc()
        d.rowptr, d.col = ptr(self.rowptr), ptr(self.col)
        d.val = ptr(self.val_t if transposed else self.val)
        d.n_rows = self.n_rows
        d.row_offset = self.row_offset
        d.n_long, d.n_items = self.n_long, self.n_items
        d.long_row, d.long_chunk = self.long_row, self.long_chunk
        d.blocked_row_begin, d.blocked_min_deg, d.chunk_lanes = self.blocked_row_begin, self.blocked_min_deg, self.chunk_lanes
        if self.n_long:
            d.long_nchunks = ptr(self.long_nchunks)
            if dim not in self._scratch:
                self._scratch[dim] = (torch.zeros(self.n_long, dim, dtype=torch.float32, device=self.device),
                                      torch.zeros(self.n_long, dtype=torch.int32, device=self.device))
            scr, cnt = self._scratch[dim]
            d.long_rows, d.item_slot = ptr(self.long_rows), ptr(self.item_slot)
            d.item_begin, d.item_end = ptr(self.item_begin), ptr(self.item_end)
            d.long_scratch, d.long_counter = ptr(scr), ptr(cnt)
        return d

    def subset_desc(self, dim, rows):
        """(tagrec_csr_t, undo) for a launch that produces the listed rows only.  ``rows`` = int32 local row ids, no
        duplicates, negative entries = blanks (skipped).
        Short rows run from the row list; the chunk list of the PLAIN plan is passed whole and the chunks of long rows
        that are not listed exit on a byte map (``row_sel``), so nothing of data-dependent size is built per call.
        ``undo()`` clears the byte map again (call it after the launch has been enqueued)."""
        d = self.desc(dim, plain=True)
        sel = self.__dict__.get("_row_sel")
        if sel is None:
            sel = self._row_sel = torch.zeros(self.n_rows + 1, dtype=torch.uint8, device=self.device)
        rows64 = rows.to(torch.int64)
        rows64 = torch.where(rows64 < 0, torch.full_like(rows64, self.n_rows), rows64)    # blanked entries -> a dummy slot
        sel.index_fill_(0, rows64, 1)
        d.row_list = ptr(rows)
        d.n_rows = int(rows.numel())
        d.row_sel = ptr(sel)
        return d, (lambda: sel.index_fill_(0, rows64, 0))

    def halves(self):
        """(user-row block, item-row block) of an unsharded bipartite graph as CsrGraph views with their own launch
        plans (user rows come first in the CSR: the user block is made of views, the item block of a re-based rowptr).
        Used by the first backward launch of a BPR step, whose two halves see very different sources (see
        functional.lightgcn_backward_layers).  None for sharded / tripartite graphs."""
        if self.comm is not None or self.row_offset != 0 or len(self.num_list) != 2 or self.n_rows != self.n:
            return None
        h = self.__dict__.get("_halves")
        if h is None:
            U = int(self.num_list[0])
            cut = int(self.rowptr[U])
            sym = self.val_t is self.val
            ub = CsrGraph(self.n, self.rowptr[:U + 1], self.col[:cut], self.val[:cut], None if sym else self.val_t[:cut],
                          None, self.norm_type, self.num_list, row_offset=0)
            ib = CsrGraph(self.n, (self.rowptr[U:] - cut).contiguous(), self.col[cut:], self.val[cut:],
                          None if sym else self.val_t[cut:], None, self.norm_type, self.num_list, row_offset=U)
            h = self._halves = (ub, ib)
        return h

    def row_slabs(self, k):
        """adj.py:114-130 split_sp_mat row folds (kept for API parity; K1 always runs on the whole CSR)."""
        f = self.n // k
        return [(i * f, self.n if i == k - 1 else (i + 1) * f) for i in range(k)]


def _as_dev_i64(x, device):
    if isinstance(x, torch.Tensor):
        return x.to(device=device, dtype=torch.int64).contiguous()
    return torch.as_tensor(np.ascontiguousarray(x, dtype=np.int64), device=device)


def build_csr(n_user, n_item, ui, norm_type, device, n_tag=0, ut=None, it=None):
    """K0.  ``ui``/``ut``/``it`` = (row_idx, col_idx) pairs (numpy or torch), block-local ids, one entry per listed
    pair (duplicates are summed like the reference's COO->LIL conversion, data/utils.py:50-53)."""
    device = torch.device(device)
    if device.type != "cuda":
        raise _lib.TagrecError("tagrec_b200 builds and multiplies the adjacency on a CUDA device only (no CPU fallback)")
    L = lib()
    with torch.cuda.device(device):
        st = stream_ptr(device)
        mode, self_loops, power = _NORM.get(norm_type, (3, 0, None))
        blocks = [(_as_dev_i64(ui[0], device), _as_dev_i64(ui[1], device))]
        use_tag = ut is not None
        if use_tag:
            blocks += [(_as_dev_i64(ut[0], device), _as_dev_i64(ut[1], device)),
                       (_as_dev_i64(it[0], device), _as_dev_i64(it[1], device))]
        else:
            blocks += [(None, None), (None, None)]
        counts = [0 if r is None else int(r.numel()) for r, _ in blocks]
        n = n_user + n_item + (n_tag if use_tag else 0)
        m = 2 * sum(counts) + (n if self_loops else 0)
        ws = torch.empty(int(L.tagrec_csr_workspace_bytes(m)), dtype=torch.uint8, device=device)
        rowptr = torch.empty(n + 1, dtype=torch.int64, device=device)
        col = torch.empty(max(m, 1), dtype=torch.int32, device=device)
        weight = torch.empty(max(m, 1), dtype=torch.float32, device=device)
        degree = torch.empty(n, dtype=torch.float32, device=device)
        nnz = C.c_int64(0)
        check(L.tagrec_csr_build_structure(ptr(blocks[0][0]), ptr(blocks[0][1]), counts[0],
                                           ptr(blocks[1][0]), ptr(blocks[1][1]), counts[1],
                                           ptr(blocks[2][0]), ptr(blocks[2][1]), counts[2],
                                           n_user, n_item, n_tag if use_tag else 0, self_loops,
                                           ptr(ws), ws.numel(), ptr(rowptr), ptr(col), ptr(weight), col.numel(),
                                           ptr(degree), C.byref(nnz), st), "tagrec_csr_build_structure")
        del ws
        nnz = nnz.value
        col = col[:nnz].clone() if nnz < col.numel() else col
        weight = weight[:nnz].clone() if nnz < weight.numel() else weight
        val_t = None
        if mode == 3:
            val = weight
        else:
            # adj.py:93-94 / 105-106 — numpy's float32 pow on the host, inf -> 0
            with np.errstate(divide="ignore"):
                dpow = np.power(degree.cpu().numpy(), power).astype(np.float32)
            dpow[np.isinf(dpow)] = 0.0
            dpow = torch.from_numpy(dpow).to(device)
            val = torch.empty(nnz, dtype=torch.float32, device=device)
            check(L.tagrec_csr_normalise(ptr(rowptr), ptr(col), ptr(weight), ptr(dpow), n, mode, self_loops,
                                         ptr(val), st), "tagrec_csr_normalise")
            if mode == 1:   # D^-1 A is not symmetric: backward needs the values of A^T = A D^-1
                val_t = torch.empty(nnz, dtype=torch.float32, device=device)
                check(L.tagrec_csr_normalise(ptr(rowptr), ptr(col), ptr(weight), ptr(dpow), n, 2, self_loops,
                                             ptr(val_t), st), "tagrec_csr_normalise")
        nums = [n_user, n_item] + ([n_tag] if use_tag else [])
        return CsrGraph(n, rowptr, col, val, val_t, weight if mode != 3 else None, norm_type, nums)


def creat_adj(data, use_tag, norm_type, split_adj_k, device):
    """adj.py:38-46.  ``data.ui_adj`` / ``ut_adj`` / ``it_adj`` are scipy COO matrices (data/cf_load.py:24,
    data/tgcn_load.py:21-22).  ``split_adj_k`` is accepted and ignored (one CSR)."""
    ui = (data.ui_adj.row, data.ui_adj.col)
    n_user, n_item = data.ui_adj.shape
    if use_tag:
        return build_csr(n_user, n_item, ui, norm_type, device, data.ut_adj.shape[1],
                         (data.ut_adj.row, data.ut_adj.col), (data.it_adj.row, data.it_adj.col))
    return build_csr(n_user, n_item, ui, norm_type, device)


def spmm_raw(graph, x, out=None, transposed=False, beta=0.0):
    """y = A x (or A^T x) through K1, no autograd."""
    assert x.is_cuda and x.dtype == torch.float32 and x.is_contiguous()
    if out is None:
        out = torch.empty_like(x)
    d = graph.desc(x.shape[1], transposed)
    check(lib().tagrec_spmm(C.byref(d), ptr(x), ptr(out), x.shape[1], float(beta), stream_ptr(x.device)),
          "tagrec_spmm")
    return out


class _SpMM(torch.autograd.Function):
    @staticmethod
    def forward(ctx, graph, x):
        ctx.graph = graph
        return spmm_raw(graph, x.contiguous())

    @staticmethod
    def backward(ctx, g):
        return None, spmm_raw(ctx.graph, g.contiguous(), transposed=True)


def split_mm(norm_adj, all_embed):
    """adj.py:158-167 — differentiable A @ E."""
    return _SpMM.apply(norm_adj, all_embed)


def node_drop(graph, keep_prob, training=False):
    """adj.py:170-191 (the argument is the DROP probability p, as in the reference): every stored entry is kept with
    probability 1 - p, independently, and the kept values are divided by 1 - p.  Identity at p == 0 or in eval.
    The draw comes from torch's CUDA generator (the reference draws on the CPU), so parity with p > 0 is statistical
    (SURVEY A19).  Entries are dropped independently of their mirror entry, so the result is not symmetric: the values
    of its transpose (for the backward SpMM) go through the reverse-edge permutation."""
    assert 0 <= keep_prob < 1
    if keep_prob == 0 or not training:
        return graph
    import copy
    from .routing import reverse_perm
    k = 1.0 - keep_prob
    keep = (torch.rand(graph._nnz(), device=graph.device) >= keep_prob).to(torch.float32) / k
    g = copy.copy(graph)
    g.val = graph.val * keep
    g.val_t = graph.val_t * keep[reverse_perm(graph).long()]
    return g
