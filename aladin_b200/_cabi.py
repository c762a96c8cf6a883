"""ctypes binding of include/alad_b200.h.  This is the only place the package touches
the native library; there is NO fallback: a missing library or a failing call raises."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# ALAD_B200_LIB: another build of the same C ABI (A/B runs of two kernel versions on one box)
LIB_PATH = os.environ.get("ALAD_B200_LIB") or os.path.join(_HERE, "libalad_b200.so")

TILE_M, TILE_N, TILE_K, MAX_SEG = 128, 240, 64, 32
NTILE_WORDS = 20


class PackArgs(C.Structure):
    _fields_ = [
        ("src", C.c_void_p), ("stride_b", C.c_int64), ("stride_s", C.c_int64),
        ("B", C.c_int32), ("S", C.c_int32), ("d", C.c_int32), ("slot0", C.c_int32),
        ("count", C.c_void_p), ("row_off", C.c_void_p), ("dst", C.c_void_p),
        ("Kp", C.c_int32), ("mode", C.c_int32), ("normalize", C.c_int32), ("eps", C.c_float),
        ("row_item", C.c_void_p), ("item_base", C.c_int32),
    ]


class MrswFwdArgs(C.Structure):
    _fields_ = [
        ("words", C.c_void_p), ("n_word_rows", C.c_int64),
        ("regions", C.c_void_p), ("n_region_rows", C.c_int64),
        ("Kp", C.c_int32), ("row_cap", C.c_void_p), ("ntiles", C.c_void_p), ("n_ntiles", C.c_int32),
        ("S", C.c_void_p), ("ldS", C.c_int64), ("Ni", C.c_int32), ("Nc", C.c_int32),
        ("epilogue", C.c_int32), ("num_ctas", C.c_int32), ("cta_group", C.c_int32),
        ("transpose_out", C.c_int32), ("accumulate", C.c_int32), ("operand_format", C.c_int32),
    ]


class ScoresFusedArgs(C.Structure):
    _fields_ = [
        ("max_x", C.c_void_p), ("max_stride_b", C.c_int64), ("max_stride_s", C.c_int64),
        ("sum_x", C.c_void_p), ("sum_stride_b", C.c_int64), ("sum_stride_s", C.c_int64),
        ("n_max", C.c_int32), ("S_max", C.c_int32), ("slot0_max", C.c_int32),
        ("n_sum", C.c_int32), ("S_sum", C.c_int32), ("slot0_sum", C.c_int32),
        ("d", C.c_int32),
        ("max_count", C.c_void_p), ("sum_count", C.c_void_p), ("max_clamp", C.c_void_p),
        ("precision", C.c_int32), ("epilogue", C.c_int32), ("normalize", C.c_int32), ("eps", C.c_float),
        ("S", C.c_void_p), ("ldS", C.c_int64), ("transpose_out", C.c_int32),
        ("workspace", C.c_void_p), ("workspace_bytes", C.c_int64),
    ]


class MrswRetrievalArgs(C.Structure):
    """struct alad_mrsw_retrieval_args (include/alad_b200.h)."""
    _fields_ = [
        ("words", C.c_void_p), ("n_word_rows", C.c_int64), ("row_cap", C.c_void_p),
        ("regions", C.c_void_p), ("n_region_rows", C.c_int64), ("Kp", C.c_int32), ("operand_format", C.c_int32),
        ("nr", C.c_void_p), ("clamp", C.c_void_p), ("cap_row", C.c_void_p),
        ("Ni", C.c_int32), ("Nc", C.c_int32), ("group", C.c_int32), ("k", C.c_int32), ("block_images", C.c_int32),
        ("S", C.c_void_p), ("ldS", C.c_int64),
        ("rank_i2t", C.c_void_p), ("top1", C.c_void_p), ("rank_t2i", C.c_void_p), ("topk_score", C.c_void_p),
        ("topk_idx", C.c_void_p), ("workspace", C.c_void_p), ("workspace_bytes", C.c_int64),
    ]


class TrainLossesArgs(C.Structure):
    """struct alad_train_losses_args (include/alad_b200.h)."""
    _fields_ = [
        ("im_cls", C.c_void_p), ("ld_im_cls", C.c_int64), ("s_cls", C.c_void_p), ("ld_s_cls", C.c_int64),
        ("im_set", C.c_void_p), ("im_stride_b", C.c_int64), ("im_stride_s", C.c_int64),
        ("s_seq", C.c_void_p), ("s_stride_b", C.c_int64), ("s_stride_s", C.c_int64),
        ("B", C.c_int32), ("S_im", C.c_int32), ("S_s", C.c_int32), ("d", C.c_int32),
        ("nr", C.c_void_p), ("nw", C.c_void_p), ("clamp", C.c_void_p),
        ("precision", C.c_int32), ("precision_m", C.c_int32),
        ("margin_m", C.c_float), ("max_violation_m", C.c_int32), ("margin_a", C.c_float), ("max_violation_a", C.c_int32),
        ("with_distill", C.c_int32), ("temperature", C.c_float), ("listnet_eps", C.c_float), ("want_grad", C.c_int32),
        ("losses", C.c_void_p), ("M", C.c_void_p), ("S", C.c_void_p),
        ("G_m", C.c_void_p), ("G_a", C.c_void_p), ("dM", C.c_void_p),
        ("g", C.c_void_p), ("has_g_m", C.c_int32), ("has_g_a", C.c_int32), ("has_g_d", C.c_int32),
        ("d_im_cls", C.c_void_p), ("d_s_cls", C.c_void_p),
        ("d_im_set", C.c_void_p), ("d_im_stride_b", C.c_int64), ("d_im_stride_s", C.c_int64),
        ("d_s_seq", C.c_void_p), ("d_s_stride_b", C.c_int64), ("d_s_stride_s", C.c_int64),
        ("workspace", C.c_void_p), ("workspace_bytes", C.c_int64),
    ]


class MrswBwdArgs(C.Structure):
    _fields_ = [
        ("im", C.c_void_p), ("im_stride_b", C.c_int64), ("im_stride_s", C.c_int64),
        ("s", C.c_void_p), ("s_stride_b", C.c_int64), ("s_stride_s", C.c_int64),
        ("Bi", C.c_int32), ("S_im", C.c_int32), ("Bc", C.c_int32), ("S_s", C.c_int32), ("d", C.c_int32),
        ("nr", C.c_void_p), ("nw", C.c_void_p),
        ("G0", C.c_void_p), ("ldG0", C.c_int64), ("g0_scale", C.c_void_p), ("G1", C.c_void_p), ("ldG1", C.c_int64),
        ("d_im", C.c_void_p), ("d_s", C.c_void_p), ("eps", C.c_float), ("region_extent", C.c_int32),
        ("max_pairs", C.c_int64),
        ("workspace", C.c_void_p), ("workspace_bytes", C.c_int64),
        ("d_im_stride_b", C.c_int64), ("d_im_stride_s", C.c_int64),
        ("d_s_stride_b", C.c_int64), ("d_s_stride_s", C.c_int64),
    ]


PTILE_SLOTS, PTILE_WORDS = 8, 24
MAX_PEERS, PEER_HANDLE_BYTES = 32, 64


class PairtileArgs(C.Structure):
    """struct alad_pairtile_args (include/alad_b200.h)."""
    _fields_ = [
        ("n_groups", C.c_int32), ("group_row0", C.c_void_p), ("group_cap_lo", C.c_void_p), ("cap_group", C.c_void_p),
        ("Nc", C.c_int32), ("lists_t2i", C.c_void_p), ("k_t2i", C.c_int32), ("lists_i2t", C.c_void_p), ("k_i2t", C.c_int32),
        ("img_off", C.c_int32), ("n_loc", C.c_int32), ("region_row", C.c_void_p), ("nr", C.c_void_p), ("clamp", C.c_void_p),
        ("slot_rows", C.c_int32), ("block_images", C.c_int32), ("ptiles", C.c_void_p), ("capacity", C.c_int32),
        ("n_ptiles", C.c_void_p),
        ("workspace", C.c_void_p), ("workspace_bytes", C.c_int64),
    ]


class MrswPairsArgs(C.Structure):
    """struct alad_mrsw_pairs_args (include/alad_b200.h)."""
    _fields_ = [
        ("words", C.c_void_p), ("n_word_rows", C.c_int64), ("regions", C.c_void_p), ("n_region_rows", C.c_int64),
        ("Kp", C.c_int32), ("row_cap", C.c_void_p), ("ptiles", C.c_void_p), ("n_ptiles", C.c_void_p),
        ("max_ptiles", C.c_int32), ("slot_rows", C.c_int32), ("S", C.c_void_p), ("ldS", C.c_int64),
        ("Ni", C.c_int32), ("Nc", C.c_int32), ("transpose_out", C.c_int32), ("num_ctas", C.c_int32),
        ("word_box_rows", C.c_int32),
    ]


# name -> (restype, argtypes); mirrors include/alad_b200.h one to one
_I32, _I64, _P = C.c_int32, C.c_int64, C.c_void_p
PROTOTYPES = {
    "alad_abi_version": (C.c_int, []),
    "alad_last_error": (C.c_char_p, []),
    "alad_h2d_2d": (C.c_int, [_P, _I64, _P, _I64, _I64, _I64, _P]),
    "alad_h2d_2d_staged": (C.c_int, [_P, _I64, _P, _I64, _I64, _I64, _I32, _P]),
    "alad_region_tiles": (C.c_int, [_P, _P, _I32, _P, _I32, _P]),
    "alad_pack_tokens": (C.c_int, [C.POINTER(PackArgs), _P]),
    "alad_mrsw_scores_fwd": (C.c_int, [C.POINTER(MrswFwdArgs), _P]),
    "alad_scores_fused_workspace_bytes": (C.c_int64, [_I32, _I32, _I32, _I32, _I32, _I32, _I32, _I32]),
    "alad_scores_fused": (C.c_int, [C.POINTER(ScoresFusedArgs), _P]),
    "alad_pool_tokens": (C.c_int, [_P, _I64, _I64, _I32, _I32, _I32, _I32, _P, C.c_float, _P, _P]),
    "alad_scale_scores": (C.c_int, [_P, _I64, _I32, _I32, _P, C.c_float, _P]),
    "alad_mrsw_bwd_workspace_bytes": (C.c_int64, [_I32, _I32, _I32, _I32, _I64]),
    "alad_mrsw_scores_bwd": (C.c_int, [C.POINTER(MrswBwdArgs), _P]),
    "alad_loss_workspace_bytes": (C.c_int64, [_I32]),
    "alad_triplet_fwd_bwd": (C.c_int, [_P, _I64, _I32, C.c_float, _I32, _P, _P, _I64, _P, _P, _P, _P]),
    "alad_listnet_fwd_bwd": (C.c_int, [_P, _I64, _P, _I64, _I32, C.c_float, C.c_float, _P, _P, _I64, _P, _P]),
    "alad_train_losses_workspace_bytes": (C.c_int64, [_I32, _I32, _I32, _I32, _I32, _I32]),
    "alad_train_losses_fwd": (C.c_int, [C.POINTER(TrainLossesArgs), _P]),
    "alad_train_losses_bwd": (C.c_int, [C.POINTER(TrainLossesArgs), _P]),
    "alad_distill_workspace_bytes": (C.c_int64, [_I32, _I32]),
    "alad_distill_mse_fwd_bwd": (C.c_int, [_P, _I64, _P, _I64, _I32, _P, _P, _P, _I64, _P, _P, _P]),
    "alad_distill_contrastive_fwd_bwd": (C.c_int, [_P, _I64, _P, _I64, _I32, C.c_float, _I32, _P, _P, _I64, _P, _P]),
    "alad_distill_ordinal_fwd_bwd": (C.c_int, [_P, _I64, _P, _I64, _I32, C.c_float, C.c_float, _I32, _P, _P, _I64, _P, _P]),
    "alad_order_scores": (C.c_int, [_P, _I64, _P, _I64, _I32, _I32, _I32, _P, _I64, _P]),
    "alad_order_scores_bwd": (C.c_int, [_P, _I64, _P, _I64, _I32, _I32, _I32, _P, _I64, _P, _I64, _P, _P, _P]),
    "alad_normalize_bwd": (C.c_int, [_P, _I64, _I64, _I32, C.c_float, _P, _I64, _P]),
    "alad_pool_tokens_bwd": (C.c_int, [_P, _I64, _I64, _I32, _I32, _I32, _I32, _P, C.c_float, _P, _P, _P]),
    "alad_scan_gram": (C.c_int, [_P, _I32, _I32, _I32, _P, _P, _P]),
    "alad_scan_gram_bwd": (C.c_int, [_P, _I32, _I32, _I32, _P, _P, _P, _P]),
    "alad_scan_pool_fwd": (C.c_int, [_P, _I64, _I32, _I32, _I32, _I32, _P, _P, _I32, _I32, _P, _P, _I64, _P]),
    "alad_scan_pool_bwd": (C.c_int, [_P, _I64, _I32, _I32, _I32, _I32, _P, _P, _I32, _I32, _P, _P, _I64, _P, _I64, _P, _P]),
    "alad_scan_apply_pairs": (C.c_int, [_P, _I64, _P, _P, _P, _I32, _I32, _I32, _I32, _I32, _I32, _P, _P, _I32, _I32, _P, _P, _P]),
    "alad_rank_rows": (C.c_int, [_P, _I64, _I32, _I32, _I32, _I32, _P, _P, _P]),
    "alad_col_gt": (C.c_int, [_P, _I64, _I32, _I32, _I32, _I32, _P, _P]),
    "alad_col_count": (C.c_int, [_P, _I64, _I32, _I32, _I32, _I32, _P, _P, _P]),
    "alad_col_topk": (C.c_int, [_P, _I64, _I32, _I32, _I32, _I32, _I32, _P, _P, _P]),
    "alad_col_topk_select_workspace_bytes": (C.c_int64, [_I32, _I32, _I32]),
    "alad_col_topk_select": (C.c_int, [_P, _I64, _I32, _I32, _I32, _I32, _P, _P, _P, _P]),
    "alad_topk_merge": (C.c_int, [_P, _P, _I32, _I32, _I32, _P, _P, _P]),
    "alad_mrsw_retrieval_workspace_bytes": (C.c_int64, [_I32, _I32, _I32, _I32, _I32]),
    "alad_mrsw_retrieval": (C.c_int, [C.POINTER(MrswRetrievalArgs), _P]),
    "alad_rank_fused_workspace_bytes": (C.c_int64, [_I32, _I32, _I32, _I32]),
    "alad_rank_fused": (C.c_int, [_P, _I64, _I32, _I32, _I32, _I32, _I32, _I32, _I32, _P, _P, _P, _P, _P, _P, _P, _P]),
    "alad_shortlist_scatter": (C.c_int, [_P, _I64, _P, _I64, _I32, _I32, _P, _I32, _I32, _I32, _I32, _P]),
    "alad_caption_groups": (C.c_int, [_P, _I32, _P, _P, _P]),
    "alad_pairtile_workspace_bytes": (C.c_int64, [_I32, _I32, _I32]),
    "alad_pairtile_build": (C.c_int, [C.POINTER(PairtileArgs), _P]),
    "alad_mrsw_scores_pairs": (C.c_int, [C.POINTER(MrswPairsArgs), _P]),
    "alad_gather_list_scores": (C.c_int, [_P, _I64, _I32, _I32, _P, _I32, _I32, _I32, _I32, _P, _P, _P, _P]),
    "alad_list_rerank": (C.c_int, [_P, _P, _I32, _I32, _I32, _I32, _I32, _I32, _P, _P, _P, _P]),
    "alad_peer_alloc": (C.c_int, [C.POINTER(C.c_void_p), _I64]),
    "alad_peer_free": (C.c_int, [_P]),
    "alad_peer_export": (C.c_int, [_P, _P]),
    "alad_peer_open": (C.c_int, [_P, C.POINTER(C.c_void_p)]),
    "alad_peer_close": (C.c_int, [_P]),
    "alad_peer_copy": (C.c_int, [_P, _P, _I64, _P]),
    "alad_peer_signal": (C.c_int, [_P, _I32, _I32, _P]),
    "alad_peer_wait": (C.c_int, [_P, _I32, _I32, _I32, _I64, _P, _P]),
    "alad_host_atomic_add": (C.c_int64, [_P, _I64]),
    "alad_host_atomic_cas": (C.c_int32, [_P, _I64, _I64]),
    "alad_host_atomic_load": (C.c_int64, [_P]),
    "alad_host_atomic_store": (None, [_P, _I64]),
}

_lib = None


class AladError(RuntimeError):
    pass


def lib():
    """Load libalad_b200.so (built in-tree by ``__graft_entry__.build()``)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise AladError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(needs nvcc). aladin_b200 has no CPU or PyTorch fallback.")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(handle, name)      # AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


# kernels launched per successful entry-point call (bench.py reports the total as gpu_launches)
KERNELS_PER_CALL = {
    "alad_h2d_2d": 0, "alad_pack_tokens": 1, "alad_pool_tokens": 1, "alad_scale_scores": 1, "alad_mrsw_scores_fwd": 1, "alad_scores_fused": 3, "alad_mrsw_scores_bwd": 4,
    "alad_triplet_fwd_bwd": 2, "alad_listnet_fwd_bwd": 3, "alad_rank_rows": 1, "alad_col_gt": 1,
    "alad_col_count": 1, "alad_col_topk": 1, "alad_col_topk_select": 5, "alad_rank_fused": 6, "alad_mrsw_retrieval": 0, "alad_topk_merge": 1, "alad_shortlist_scatter": 2,
    "alad_distill_mse_fwd_bwd": 1, "alad_distill_contrastive_fwd_bwd": 2, "alad_distill_ordinal_fwd_bwd": 2,
    "alad_train_losses_fwd": 13, "alad_train_losses_bwd": 16,
    "alad_order_scores": 1, "alad_order_scores_bwd": 1, "alad_normalize_bwd": 1, "alad_pool_tokens_bwd": 1,
    "alad_pairtile_build": 5, "alad_mrsw_scores_pairs": 1, "alad_gather_list_scores": 1, "alad_list_rerank": 1,
    "alad_peer_signal": 1, "alad_peer_wait": 1,
    "alad_scan_gram": 1, "alad_scan_gram_bwd": 1, "alad_scan_pool_fwd": 1, "alad_scan_pool_bwd": 1, "alad_scan_apply_pairs": 1,
}
launch_count = {"kernels": 0}


def check(rc, what):
    launch_count["kernels"] += KERNELS_PER_CALL.get(what, 0)
    if rc != 0:
        msg = lib().alad_last_error()
        raise AladError(f"{what} failed (status {rc}): {msg.decode() if msg else ''}")


def stream_ptr():
    """Raw cudaStream_t of torch's current stream on the current device (one C call: this runs before every
    entry point, and torch.cuda.current_stream() costs several microseconds of Python)."""
    import torch
    return torch._C._cuda_getCurrentRawStream(torch._C._cuda_getDevice()) or None
