"""ctypes binding of libtagrec_b200.so (the C ABI declared in include/tagrec_b200.h).

There is NO fallback: if the shared library is missing or a call fails, this raises.  The product never routes
through the CPU oracle or through torch ops for the kernels it declares.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("TAGREC_LIB") or os.path.join(_HERE, "libtagrec_b200.so")   # TAGREC_LIB: tuning builds

LONG_ROW = 4096       # TAGREC_LONG_ROW   (defaults of the long-row plan, tuned on the 1.9e9-nnz graph)
LONG_CHUNK = 2048     # TAGREC_LONG_CHUNK
SMALL_GRAPH_NNZ = 1 << 25     # below this many entries a graph is a few waves of rows: plan with 256 / 256 instead
SMALL_LONG_ROW = 256
SMALL_LONG_CHUNK = 256

_p = C.c_void_p
_i64 = C.c_int64
_i32 = C.c_int
_f32 = C.c_float
_sz = C.c_size_t
_u64 = C.c_uint64
_u32 = C.c_uint32


class CsrDesc(C.Structure):
    """tagrec_csr_t"""
    _fields_ = [("rowptr", _p), ("col", _p), ("val", _p), ("n_rows", _i64), ("row_offset", _i64), ("long_rows", _p), ("item_slot", _p),
                ("item_begin", _p), ("item_end", _p), ("n_long", _i64), ("n_items", _i64), ("long_scratch", _p),
                ("long_counter", _p), ("long_row", C.c_int32), ("long_chunk", C.c_int32), ("long_nchunks", _p),
                ("blocked_row_begin", _i64), ("blocked_min_deg", C.c_int32), ("chunk_lanes", C.c_int32), ("row_list", _p), ("row_sel", _p)]


class RoutePlan(C.Structure):
    """tagrec_route_plan_t"""
    _fields_ = [("long_rows", _p), ("n_long", _i64), ("piece_slot", _p), ("piece_begin", _p), ("piece_end", _p),
                ("n_pieces", _i64), ("scratch", _p)]


class MirrorDesc(C.Structure):
    """tagrec_mirror_t"""
    _fields_ = [("n", C.c_int32), ("self", C.c_int32), ("base", _p * 8)]


class AdamDesc(C.Structure):
    """tagrec_adam_t"""
    _fields_ = [("param", _p), ("exp_avg", _p), ("exp_avg_sq", _p), ("lr", C.c_float), ("beta1", C.c_float),
                ("beta2", C.c_float), ("eps", C.c_float), ("weight_decay", C.c_float), ("reserved", C.c_int32),
                ("step", C.c_int64), ("param_mirror", MirrorDesc)]


# name -> (restype, argtypes); must list every symbol of include/tagrec_b200.h (tests/test_abi.py checks that)
PROTOTYPES = {
    "tagrec_version": (_i32, []),
    "tagrec_last_error": (C.c_char_p, []),
    "tagrec_launch_count": (_u64, []),
    "tagrec_sizeof_struct": (_sz, [_i32]),
    "tagrec_csr_workspace_bytes": (_sz, [_i64]),
    "tagrec_csr_build_structure": (_i32, [_p, _p, _i64, _p, _p, _i64, _p, _p, _i64, _i64, _i64, _i64, _i32, _p, _sz,
                                          _p, _p, _p, _i64, _p, C.POINTER(_i64), _p]),
    "tagrec_csr_normalise": (_i32, [_p, _p, _p, _p, _i64, _i32, _i32, _p, _p]),
    "tagrec_lightgcn_bwd_first_sparse": (_i32, [_p, _i64, _i64, _i64, _p, _p, _p, _f32, _p, _p, _i32, _p]),
    "tagrec_rows_zero": (_i32, [_p, _i64, _p, _p, _i32, _p]),
    "tagrec_csr_window_bounds": (_i32, [_p, _p, _p, _i64, _i64, _i32, _p, _p]),
    "tagrec_spmm": (_i32, [C.POINTER(CsrDesc), _p, _p, _i32, _f32, _p]),
    "tagrec_lightgcn_fwd_layer": (_i32, [C.POINTER(CsrDesc), _p, _p, _p, _i32, _i32, _i32, _f32, _p]),
    "tagrec_lightgcn_bwd_layer": (_i32, [C.POINTER(CsrDesc), _p, _p, _p, _p, _p, _f32, _p, _i32, _p]),
    "tagrec_lightgcn_fwd_layer_p2p": (_i32, [C.POINTER(CsrDesc), _p, _p, _p, _i32, _i32, _i32, _f32,
                                             C.POINTER(MirrorDesc), C.POINTER(MirrorDesc), _p]),
    "tagrec_lightgcn_bwd_layer_p2p": (_i32, [C.POINTER(CsrDesc), _p, _p, _p, _p, _p, _f32, _p, _i32,
                                             C.POINTER(MirrorDesc), _p]),
    "tagrec_lightgcn_bwd_layer_ex": (_i32, [C.POINTER(CsrDesc), _p, _p, _p, _p, _p, _p, _f32, _p, _i32,
                                            C.POINTER(MirrorDesc), _p]),
    "tagrec_lightgcn_bwd_layer_adam": (_i32, [_p, _p, _p, _p, _p, _p, _f32, _p, _i32, _p, _p]),
    "tagrec_spmm_push_rows": (_i32, [_p, _p, _p, _p, _p, _i64, _p, _p, _i32, _p]),
    "tagrec_lightgcn_bwd_layer_acc": (_i32, [_p, _p, _p, _p, _p, _f32, _p, _i32, _p, _p]),
    "tagrec_row_nonzero": (_i32, [_p, _i64, _i32, _p, _p]),
    "tagrec_bpr_fwd_bwd": (_i32, [_p, _i64, _i64, _p, _p, _i32, _f32, _i32, _p, _p, _p, _p]),
    "tagrec_eval_topk": (_i32, [_p, _i64, _p, _p, _i64, _i32, _p, _p, _i32, _p, _p, _p, _sz, _p]),
    "tagrec_eval_workspace_bytes": (_sz, [_i64, _i64, _i32]),
    "tagrec_eval_plan": (_i32, [_i64, _i64, _i32, _i32, _p]),
    "tagrec_eval_topk_ex": (_i32, [_p, _i64, _p, _p, _i64, _i32, _p, _p, _i32, _p, _p, _p, _sz, _i32, _p]),
    "tagrec_eval_auc_workspace_bytes": (_sz, [_i64, _i64]),
    "tagrec_eval_auc": (_i32, [_p, _i64, _p, _p, _i64, _i32, _p, _p, _p, _p, _i64, _p, _sz, _p, _p]),
    "tagrec_eval_auc_ex": (_i32, [_p, _i64, _p, _p, _i64, _i32, _p, _p, _p, _p, _i64, _p, _sz, _p, _i32, _p]),
    "tagrec_eval_metrics": (_i32, [_p, _i64, _p, _i32, _p, _p, _p, _i32, _p, _p]),
    "tagrec_ngcf_dense_fwd": (_i32, [_p, _p, _p, _p, _p, _p, _i64, _i32, _p, _p, _p, _p, _p]),
    "tagrec_ngcf_dense_bwd": (_i32, [_p, _p, _i64, _p, _p, _p, _p, _p, _p, _p, _p, _p, _i64, _i32, _p, _p, _p, _p, _p, _p,
                                     _p]),
    "tagrec_edge_softmax_rowsum": (_i32, [_p, _i64, _i64, _p, _p, _p, _p]),
    "tagrec_edge_scale": (_i32, [_p, _p, _i64, _p, _p, _p, _p]),
    "tagrec_spmm4": (_i32, [_p, _p, _i64, C.POINTER(RoutePlan), _p, _p, _p, _p, _p, _p, _p, _p, _i32, _i32, _f32, _p]),
    "tagrec_spmm4_long_threshold": (_i32, []),
    "tagrec_spmm4_piece": (_i32, []),
    "tagrec_edge_dot4": (_i32, [_p, _p, _i64, _p, _p, _p, _i32, _p]),
    "tagrec_chunk_normalize": (_i32, [_p, _i64, _i32, _p, _p]),
    "tagrec_chunk_normalize_bwd": (_i32, [_p, _p, _i64, _p, _p]),
    "tagrec_csr_reverse_perm": (_i32, [_p, _p, _i64, _p, _p, _p]),
    "tagrec_nbr_attention_fwd": (_i32, [_p, _p, _p, _p, _p, _p, _p, _i64, _i32, _i64, _i32, _i32, _p, _p, _p]),
    "tagrec_nbr_attention_bwd": (_i32, [_p, _p, _p, _p, _p, _p, _p, _p, _p, _i64, _i32, _i64, _i32, _i32, _i32,
                                        _p, _p, _p, _p, _p, _p]),
    "tagrec_tgcn_tail_workspace_bytes": (_sz, [_i64, _i32]),
    "tagrec_tgcn_tail_fwd": (_i32, [_p, _p, _p, _p, _p, _i64, _i32, _i32, _i32, _p, _p]),
    "tagrec_tgcn_tail_fwd_workspace_bytes": (_sz, [_i32]),
    "tagrec_tgcn_tail_fwd_ex": (_i32, [_p, _p, _p, _p, _p, _i64, _i32, _i32, _i32, _p, _p, _sz, _i32, _p]),
    "tagrec_tgcn_tail_bwd": (_i32, [_p, _p, _p, _p, _p, _p, _i64, _i32, _i32, _i32, _p, _sz, _p, _p, _p, _p, _p, _p]),
    "tagrec_tgcn_tail_bwd_ex": (_i32, [_p, _p, _p, _p, _p, _p, _i64, _i32, _i32, _i32, _p, _sz, _p, _p, _p, _p, _p, _i32,
                                       _p]),
    "tagrec_tgcn_mix_fwd": (_i32, [_p] * 9 + [_i64, _i32, _i32, _i32, _p, _p, _p]),
    "tagrec_tgcn_mix_bwd": (_i32, [_p] * 9 + [_i64, _i32, _i32, _i32] + [_p] * 13 + [_p, _sz, _p]),
    "tagrec_tgcn_mix_workspace_bytes": (_sz, [_i64]),
    "tagrec_xty_acc": (_i32, [_p, _p, _i64, _i32, _i32, _p, _i32, _p]),
    "tagrec_xty": (_i32, [_p, _p, _i64, _i32, _i32, _p, _p]),
    "tagrec_mt19937_seed": (None, [_u32, _p]),
    "tagrec_sample_bpr_host": (_i32, [_p, _p, _i64, _p, _p, _i64, _p]),
    "tagrec_sample_neg_tail_host": (_i32, [_p, _p, _i64, _p, _p, _i64, _p]),
    "tagrec_sample_bpr_device": (_i32, [_p, _i64, _p, _p, _i64, _u64, _u64, _p, _p]),
    "tagrec_neighbor_table": (_i32, [_p, _p, _p, _i64, _i64, _i32, _i32, _i32, _u64, _u32, _p, _p, _p]),
    "tagrec_adam_step": (_i32, [_p, _p, _p, _p, _i64, _f32, _f32, _f32, _f32, _f32, _i64, _p]),
    "tagrec_adam_step_mirror": (_i32, [_p, _p, _p, _p, _i64, _f32, _f32, _f32, _f32, _f32, _i64, C.POINTER(MirrorDesc), _p]),
    "tagrec_adam_advance": (_i32, [_p, _f32, _f32, _f32, _p, _p]),
    "tagrec_adam_step_dev": (_i32, [_p, _p, _p, _p, _i64, _f32, _f32, _f32, _f32, _p, _p]),
}

_lib = None


class TagrecError(RuntimeError):
    pass


def lib():
    """The loaded library.  Raises (never falls back) when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise TagrecError(f"{LIB_PATH} is missing — run `python -c 'import __graft_entry__ as g; g.build()'` "
                              f"(nvcc -gencode arch=compute_100a,code=sm_100a); there is no CPU fallback")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        for which, cls in enumerate((CsrDesc, MirrorDesc, RoutePlan, AdamDesc)):
            want, got = int(handle.tagrec_sizeof_struct(which)), C.sizeof(cls)
            if want != got:
                raise TagrecError(f"ctypes layout of {cls.__name__} is {got} bytes but {LIB_PATH} was compiled with "
                                  f"{want}: _lib.py and include/tagrec_b200.h have drifted (or the .so is stale)")
        _lib = handle
    return _lib


def check(code, what):
    if code != 0:
        msg = lib().tagrec_last_error()
        raise TagrecError(f"{what} failed with code {code}: {msg.decode() if msg else '?'}")


def ptr(t):
    """Device (or host) address of a torch tensor / numpy array, None -> NULL."""
    if t is None:
        return None
    if hasattr(t, "data_ptr"):
        return t.data_ptr()
    return t.ctypes.data


def stream_ptr(device=None):
    import torch
    return torch.cuda.current_stream(device).cuda_stream


def launch_count():
    return int(lib().tagrec_launch_count())
