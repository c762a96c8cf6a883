"""ctypes binding of include/alad_b200.h.  This is the only place the package touches
the native library; there is NO fallback: a missing library or a failing call raises."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libalad_b200.so")

TILE_M, TILE_N, TILE_K, MAX_SEG = 128, 240, 64, 32
NTILE_WORDS = 20


class PackArgs(C.Structure):
    _fields_ = [
        ("src", C.c_void_p), ("stride_b", C.c_int64), ("stride_s", C.c_int64),
        ("B", C.c_int32), ("S", C.c_int32), ("d", C.c_int32), ("slot0", C.c_int32),
        ("count", C.c_void_p), ("row_off", C.c_void_p), ("dst", C.c_void_p),
        ("Kp", C.c_int32), ("mode", C.c_int32), ("normalize", C.c_int32), ("eps", C.c_float),
        ("row_item", C.c_void_p),
    ]


class MrswFwdArgs(C.Structure):
    _fields_ = [
        ("words", C.c_void_p), ("n_word_rows", C.c_int64),
        ("regions", C.c_void_p), ("n_region_rows", C.c_int64),
        ("Kp", C.c_int32), ("row_cap", C.c_void_p), ("ntiles", C.c_void_p), ("n_ntiles", C.c_int32),
        ("S", C.c_void_p), ("ldS", C.c_int64), ("Ni", C.c_int32), ("Nc", C.c_int32),
        ("epilogue", C.c_int32), ("num_ctas", C.c_int32),
    ]


# name -> (restype, argtypes); mirrors include/alad_b200.h one to one
_I32, _I64, _P = C.c_int32, C.c_int64, C.c_void_p
PROTOTYPES = {
    "alad_abi_version": (C.c_int, []),
    "alad_last_error": (C.c_char_p, []),
    "alad_pack_tokens": (C.c_int, [C.POINTER(PackArgs), _P]),
    "alad_mrsw_scores_fwd": (C.c_int, [C.POINTER(MrswFwdArgs), _P]),
    "alad_rank_rows": (C.c_int, [_P, _I64, _I32, _I32, _I32, _I32, _P, _P, _P]),
    "alad_col_gt": (C.c_int, [_P, _I64, _I32, _I32, _I32, _I32, _P, _P]),
    "alad_col_count": (C.c_int, [_P, _I64, _I32, _I32, _I32, _I32, _P, _P, _P]),
    "alad_col_topk": (C.c_int, [_P, _I64, _I32, _I32, _I32, _I32, _I32, _P, _P, _P]),
    "alad_topk_merge": (C.c_int, [_P, _P, _I32, _I32, _I32, _P, _P, _P]),
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


def check(rc, what):
    if rc != 0:
        msg = lib().alad_last_error()
        raise AladError(f"{what} failed (status {rc}): {msg.decode() if msg else ''}")


def stream_ptr():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)
