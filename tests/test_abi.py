"""CPU-only: the C-ABI library loads and exports every symbol include/tagrec_b200.h declares (no GPU calls)."""
import ctypes
import os
import re

import numpy as np

import tagrec_b200 as T
from helpers import nums, user_lists
from tagrec_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "tagrec_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(tagrec_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported_and_bound():
    names = declared_symbols()
    assert len(names) >= 15
    handle = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(handle, n), f"{n} declared in tagrec_b200.h but not exported"
        assert n in _lib.PROTOTYPES, f"{n} has no ctypes prototype in _lib.py"
    assert sorted(_lib.PROTOTYPES) == names


def test_struct_layouts_match_the_compiled_library():
    """tagrec_sizeof_struct() is what the .so was compiled with; the ctypes mirrors (and the struct INTEGRATION.md
    shows a maintainer) must agree field for field."""
    L = _lib.lib()
    for which, cls in enumerate((_lib.CsrDesc, _lib.MirrorDesc, _lib.RoutePlan, _lib.AdamDesc)):
        assert L.tagrec_sizeof_struct(which) == ctypes.sizeof(cls), cls.__name__
    assert L.tagrec_sizeof_struct(99) == 0
    # INTEGRATION.md's reference-side stub declares the same fields, in the same order, as _lib.CsrDesc
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    block = doc[doc.index("class Csr(C.Structure)"):]
    block = block[:block.index("]\n") + 1]
    fields = re.findall(r'\("([a-z_]+)",', block)
    assert fields == [f[0] for f in _lib.CsrDesc._fields_], fields


def test_version_and_error_text():
    L = _lib.lib()
    assert L.tagrec_version() >= 100
    # argument validation needs no GPU: a null output pointer is rejected with text
    rc = L.tagrec_spmm(None, None, None, 64, 0.0, None)
    assert rc == -1 and b"null" in L.tagrec_last_error()


def test_argument_validation_of_the_dense_kernels_needs_no_gpu():
    """Every entry point validates its arguments before it touches the device: wrong widths / paths / null pointers
    come back as TAGREC_EINVAL (-1) with the reason in tagrec_last_error() — the error behaviour the Python layer
    turns into TagrecError."""
    L = _lib.lib()
    one = ctypes.c_void_p(16)          # a non-null dummy pointer; never dereferenced by the checks
    cases = [
        (lambda: L.tagrec_xty(one, one, 10, 5, 32, one, None), b"multiples of 4"),
        (lambda: L.tagrec_xty(one, one, 10, 64, 68, one, None), b"multiples of 4"),
        (lambda: L.tagrec_xty(None, one, 10, 64, 32, one, None), b"null"),
        (lambda: L.tagrec_tgcn_tail_fwd(one, one, one, one, one, 10, 32, 32, 48, one, None), b"64-d"),
        (lambda: L.tagrec_tgcn_tail_fwd(one, one, one, one, one, 10, 64, 32, 50, one, None), b"multiple of 4"),
        (lambda: L.tagrec_tgcn_tail_fwd(one, None, one, one, one, 10, 64, 32, 48, one, None), b"null"),
        (lambda: L.tagrec_tgcn_mix_fwd(one, one, one, one, one, one, one, one, one, 10, 64, 32, 5, one, one, None),
         b"num_vec_conv"),
        (lambda: L.tagrec_tgcn_mix_fwd(one, one, one, one, one, one, one, one, one, 10, 64, 16, 8, one, one, None),
         b"dim_atten"),
        (lambda: L.tagrec_eval_auc_ex(one, 4, one, one, 100, 64, one, one, one, one, 8, one, 1 << 20, one, 7, None),
         b"bad path"),
        (lambda: L.tagrec_eval_auc_ex(one, 4, one, one, 100, 96, one, one, one, one, 8, one, 1 << 20, one, 2, None),
         b"dim 64"),
        (lambda: L.tagrec_nbr_attention_fwd(one, one, one, one, one, one, one, 10, 40, 40, 64, 32, one, one, None),
         b"neighbor_k"),
    ]
    # the optimizer epilogue of K1's last backward launch (round 2)
    csr = _lib.CsrDesc()
    csr.rowptr, csr.col, csr.val, csr.n_rows = 16, 16, 16, 4
    ad = _lib.AdamDesc()
    ad.param, ad.exp_avg, ad.exp_avg_sq, ad.step = 16, 16, 16, 0
    cases.append((lambda: L.tagrec_lightgcn_bwd_layer_adam(ctypes.byref(csr), one, None, one, None, None, 0.25, None, 64,
                                                           ctypes.byref(ad), None), b"step counts from 1"))
    ad2 = _lib.AdamDesc()
    ad2.param, ad2.exp_avg, ad2.exp_avg_sq, ad2.step = 16, 0, 16, 1
    cases.append((lambda: L.tagrec_lightgcn_bwd_layer_adam(ctypes.byref(csr), one, None, one, None, None, 0.25, None, 64,
                                                           ctypes.byref(ad2), None), b"null table"))
    ad3 = _lib.AdamDesc()
    ad3.param, ad3.exp_avg, ad3.exp_avg_sq, ad3.step = 20, 16, 16, 1
    cases.append((lambda: L.tagrec_lightgcn_bwd_layer_adam(ctypes.byref(csr), one, None, one, None, None, 0.25, None, 64,
                                                           ctypes.byref(ad3), None), b"16-byte aligned"))
    cases.append((lambda: L.tagrec_lightgcn_bwd_layer_adam(ctypes.byref(csr), one, None, one, None, None, 0.25, None, 64,
                                                           None, None), b"adam is null"))
    cases.append((lambda: L.tagrec_eval_plan(10, 10, 64, 5, None), b"plan is null"))
    for call, text in cases:
        assert call() == -1
        assert text in L.tagrec_last_error(), (text, L.tagrec_last_error())
    # n == 0 is a no-op that succeeds (no launch)
    assert L.tagrec_tgcn_tail_fwd(one, one, one, one, one, 0, 64, 32, 48, one, None) == 0
    assert L.tagrec_tgcn_mix_fwd(one, one, one, one, one, one, one, one, one, 0, 64, 32, 8, one, one, None) == 0


def test_missing_library_fails_loudly(monkeypatch):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libtagrec_b200.so")
    try:
        _lib.lib()
        raise AssertionError("expected TagrecError")
    except _lib.TagrecError as e:
        assert "no CPU fallback" in str(e)


def test_host_sampler_bit_exact_vs_reference(tiny, medium):
    """tagrec_sample_bpr_host (C++, host) == the reference's BPR_training_data with cpu_core=1, two epochs."""
    L = _lib.lib()
    for g in (tiny, medium):
        U, I, _, _ = nums(g)
        ptr_, items = T.bpr_training_data.user_items_to_csr(user_lists(g, "train"), U)
        state = np.empty(625, dtype=np.uint32)
        L.tagrec_mt19937_seed(2020, _lib.ptr(state))
        edges = np.ascontiguousarray(g["edge_index_train"], dtype=np.int64)
        for want in (g["sampler_first"], g["sampler_second"]):
            out = np.empty((len(edges), 3), dtype=np.int64)
            rc = L.tagrec_sample_bpr_host(_lib.ptr(state), _lib.ptr(edges), len(edges), _lib.ptr(ptr_),
                                          _lib.ptr(items), I, _lib.ptr(out))
            assert rc == 0
            assert np.array_equal(out, want)


def test_sampler_class_mt19937_mode_uses_numpy_global_state(tiny):
    """Drop-in class in parity mode: seeds come from np.random like the reference (init_seed -> np.random.seed)."""
    import torch
    U, I, _, _ = nums(tiny)
    T.set_config("lightgcn", train_batch=64, cpu_core=1, sampler="mt19937", device=torch.device("cpu"))

    class D:
        pass
    d = D()
    d.num = {"user": U, "item": I}
    d.user_items = {"train": user_lists(tiny, "train")}
    d.edge_index = {"train": tiny["edge_index_train"]}
    np.random.seed(2020)
    s = T.BPR_training_data(d, None)
    assert np.array_equal(s.all_train_data.numpy(), tiny["sampler_first"])
    s.reset()
    assert np.array_equal(s.all_train_data.numpy(), tiny["sampler_second"])
    assert [len(b) for b in s.mini_batch()] == list(tiny["sampler_batch_sizes"])


def test_dgcf_sampler_host_mode_bit_exact_vs_reference(tiny):
    """T.DGCF_training_data in parity mode == the reference's class (train_data/bpr_training_data.py:47-84) after the
    same random.seed / np.random.seed: sampled users, positives, rejection-sampled negatives and the cor indices."""
    import random
    import torch
    gold = dict(np.load(os.path.join(ROOT, "tests", "golden", "dgcf_sampler.npz")))
    U, I, Tg, _ = nums(tiny)

    class D:
        pass
    d = D()
    d.num = {"user": U, "item": I, "tag": Tg}
    d.user_items = {"train": user_lists(tiny, "train")}
    d.edge_index = {"train": tiny["edge_index_train"]}
    for use_tag in (False, True):
        tag = "tag" if use_tag else "notag"
        T.set_config("dgcf", train_batch=16, use_tag=use_tag, cor_batch=10, sampler="mt19937", device=torch.device("cpu"))
        random.seed(5)
        np.random.seed(5)
        s = T.DGCF_training_data(d, None)
        s.reset()
        batches = list(s.mini_batch())
        assert len(batches) == len(tiny["edge_index_train"]) // 16 + 1
        assert np.array_equal(np.stack([b[0].numpy() for b in batches]), gold[f"{tag}_data"])
        assert np.array_equal(np.stack([b[1].numpy() for b in batches]), gold[f"{tag}_cor"])
        T.set_config("dgcf", train_batch=64, use_tag=use_tag, cor_batch=10, sampler="mt19937", device=torch.device("cpu"))
        random.seed(6)
        np.random.seed(6)
        s = T.DGCF_training_data(d, None)
        assert np.array_equal(np.stack([b[0].numpy() for b in s.mini_batch()]), gold[f"{tag}_small_data"])


def test_transtag_sampler_host_mode_bit_exact_vs_reference(tiny):
    """T.TransTag_training_data (C++ numpy-legacy stream) == train_data/transe_training_data.py:42-70, two epochs; the
    global generator is left untouched, like the reference's forked worker leaves the parent's."""
    import torch
    gold = dict(np.load(os.path.join(ROOT, "tests", "golden", "transtag_sampler.npz")))
    U, I, Tg, _ = nums(tiny)

    class D:
        pass
    d = D()
    d.num = {"user": U, "item": I, "tag": Tg}
    d.uit_data = tiny["uit_data"]
    T.set_config("tgcn", transtag_batch=64, sampler="mt19937", device=torch.device("cpu"), cpu_core=1)
    np.random.seed(2020)
    before = np.random.get_state()[1].copy()
    s = T.TransTag_training_data(d, None)
    assert np.array_equal(s.all_train_data.numpy(), gold["first"])
    s.reset()
    assert np.array_equal(s.all_train_data.numpy(), gold["second"])
    assert np.array_equal(before, np.random.get_state()[1])
    sizes = [b.shape[0] for b in s.mini_batch()]
    assert sum(sizes[:-1]) + sizes[-1] >= len(gold["first"]) and sizes[0] == 64


def test_neighbour_tables_bit_exact_vs_reference(tiny, tiny_tgcn):
    """T.data.get_all_neighbor == TGCN_load.get_all_neighbor (data/tgcn_load.py:41-53, data/utils.py:87-106): the six
    padded neighbour / weight tables the reference built after init_seed(2020), incl. table widths that count
    duplicate COO entries."""
    import scipy.sparse as sp
    from helpers import blocks
    U, I, Tg, _ = nums(tiny)
    ui, ut, it = blocks(tiny)

    class D:
        pass
    d = D()
    coo = lambda rc, shape: sp.coo_matrix((np.ones(len(rc[0])), rc), dtype=np.float32, shape=shape)  # noqa: E731
    d.ui_adj, d.ut_adj, d.it_adj = coo(ui, (U, I)), coo(ut, (U, Tg)), coo(it, (I, Tg))
    np.random.seed(2020)
    before = np.random.get_state()[1].copy()
    tabs = T.data.get_all_neighbor(d)
    for name, (ids, w) in zip(["ui", "ut", "iu", "it", "tu", "ti"], tabs):
        assert np.array_equal(ids, tiny_tgcn[f"tgcn_nbr_{name}"]), name
        assert np.array_equal(w, tiny_tgcn[f"tgcn_nbw_{name}"]), name
    assert np.array_equal(before, np.random.get_state()[1])


def test_eval_plan_picks_the_pair_kernel(monkeypatch):
    """tagrec_eval_plan (host only): the benchmark shape of the evaluation leg runs on CTA pairs (eval_tc2_kernel,
    tcgen05.mma.cta_group::2 M256 x N256), small user batches and wide tables on the single-CTA kernels, K-lists too
    large for the pair kernel's shared memory fall back; TAGREC_EVAL_CG2=0 switches the pair kernel off."""
    from tagrec_b200.eval_ops import eval_plan
    monkeypatch.delenv("TAGREC_EVAL_CG2", raising=False)
    p = eval_plan(16384, 2_000_000, 64, 20)
    assert p["tensor_core"] == 1 and p["cta_pairs"] == 1 and p["lists"] == 2 * p["splits"] and p["stages"] >= 3
    assert eval_plan(512, 2_000_000, 64, 20)["cta_pairs"] == 0           # few users: one 128-user half per CTA, many splits
    assert eval_plan(16384, 500_000, 256, 20)["cta_pairs"] == 0          # wide tables: eval_tc_wide_kernel
    assert eval_plan(16384, 2_000_000, 64, 100)["cta_pairs"] == 0        # 2 x 100 x 256 list entries do not fit beside the stages
    assert eval_plan(16384, 2_000_000, 48, 20)["tensor_core"] == 0       # fp32 tiles
    monkeypatch.setenv("TAGREC_EVAL_CG2", "0")
    assert eval_plan(16384, 2_000_000, 64, 20)["cta_pairs"] == 0
    monkeypatch.setenv("TAGREC_EVAL_CG2", "force")
    assert eval_plan(130, 5000, 64, 20)["cta_pairs"] == 1
