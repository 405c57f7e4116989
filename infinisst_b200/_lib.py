"""ctypes binding of libinfinisst_b200.so (the C-ABI declared in include/infinisst_b200.h).

There is no CPU fallback: if the library is missing it is built with nvcc when a toolchain is
present, otherwise importing the product path raises."""
from __future__ import annotations

import ctypes as C
import os
import shutil

from . import build as _build

MAX_CONV = 8
DTYPE_F32, DTYPE_BF16 = 0, 1


class IsstConfig(C.Structure):
    _fields_ = [
        ("n_conv", C.c_int), ("conv_dim", C.c_int * MAX_CONV), ("conv_k", C.c_int * MAX_CONV),
        ("conv_s", C.c_int * MAX_CONV),
        ("enc_dim", C.c_int), ("enc_ffn", C.c_int), ("enc_heads", C.c_int), ("enc_layers", C.c_int),
        ("block_size", C.c_int), ("max_cache_size", C.c_int),
        ("n_adapter", C.c_int), ("adapter_dim", C.c_int * MAX_CONV), ("adapter_k", C.c_int * MAX_CONV),
        ("adapter_s", C.c_int * MAX_CONV),
        ("hidden", C.c_int), ("layers", C.c_int), ("heads", C.c_int), ("kv_heads", C.c_int),
        ("head_dim", C.c_int), ("ffn", C.c_int), ("vocab", C.c_int), ("rms_eps", C.c_float),
        ("max_streams", C.c_int), ("max_batch", C.c_int), ("max_multiplier", C.c_int), ("kv_pages", C.c_int),
        ("max_kv_len", C.c_int), ("max_prompt", C.c_int), ("max_new_tokens", C.c_int),
        ("enc_xpos", C.c_int), ("enc_no_rope", C.c_int),
    ]


class IsstGenParams(C.Structure):
    _fields_ = [
        ("max_new_tokens", C.c_int), ("no_repeat_ngram_size", C.c_int), ("repetition_penalty", C.c_float),
        ("n_eos", C.c_int), ("eos_token_ids", C.c_int * 8), ("n_suppress", C.c_int),
        ("suppress_tokens", C.POINTER(C.c_int32)), ("pin_prefix", C.c_int),
    ]


class IsstBeamFollow(C.Structure):
    _fields_ = [("closed", C.POINTER(C.c_int32)), ("next", C.POINTER(C.c_int32)), ("steps", C.POINTER(C.c_int32)),
                ("done", C.POINTER(C.c_int32))]


class IsstBeamTrace(C.Structure):
    _fields_ = [("cand_scores", C.POINTER(C.c_float)), ("cand_index", C.POINTER(C.c_int32)),
                ("next", C.POINTER(C.c_int32)), ("next_scores", C.POINTER(C.c_float)), ("steps", C.POINTER(C.c_int32))]


# every symbol include/infinisst_b200.h declares: name -> (restype, argtypes)
_P = C.c_void_p
_I = C.c_int
_IP = C.POINTER(C.c_int)
_I32P = C.POINTER(C.c_int32)
SYMBOLS = {
    "isst_last_error": (C.c_char_p, []),
    "isst_create": (_I, [C.POINTER(IsstConfig), _I, C.POINTER(_P)]),
    "isst_destroy": (None, [_P]),
    "isst_load_weight": (_I, [_P, C.c_char_p, _P, C.POINTER(C.c_int64), _I, _I]),
    "isst_finalize_weights": (_I, [_P]),
    "isst_stream_open": (_I, [_P, _IP]),
    "isst_stream_close": (_I, [_P, _I]),
    "isst_encode_chunk": (_I, [_P, _I, _IP, _P, _I, _I, _P, _P]),
    "isst_generate": (_I, [_P, _I, _IP, _I32P, _IP, _I32P, _I32P, _IP, C.POINTER(IsstGenParams), _I32P, _I32P,
                           _IP, _P]),
    "isst_generate_beam": (_I, [_P, _I, _IP, _I32P, _IP, _I32P, _I32P, _IP, C.POINTER(IsstGenParams), _I, C.c_float,
                                C.POINTER(IsstBeamFollow), _I32P, _IP, C.POINTER(C.c_float), C.POINTER(IsstBeamTrace),
                                _P]),
    "isst_forward": (_I, [_P, _I, _IP, _I32P, _IP, _I32P, _P, _I, _P, _P]),
    "isst_forward_all": (_I, [_P, _I, _IP, _I32P, _IP, _I32P, _P, _I, _P, _P]),
    "isst_kv_len": (_I, [_P, _I, _IP]),
    "isst_kv_evict": (_I, [_P, _I, _I, _I]),
    "isst_enc_steps": (_I, [_P, _I, _IP]),
    "isst_debug_enable": (_I, [_P, _I]),
    "isst_debug_read": (_I, [_P, C.c_char_p, _P, C.c_int64, C.POINTER(C.c_int64)]),
    "isst_launch_count": (C.c_int64, [_P]),
    "isst_pages_free": (_I, [_P]),
    "isst_debug_shift_positions": (_I, [_P, _I, C.c_int64]),
    "isst_profile_enable": (_I, [_P, _I]),
    "isst_profile_reset": (_I, [_P]),
    "isst_profile_read": (_I, [_P, _I, C.c_char_p, _I, C.POINTER(C.c_int64), C.POINTER(C.c_double),
                               C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "isst_op_gemm": (_I, [_P, _P, _P, _I, _I, _I, _P, _I, _P, _I, _P, _I, _I, _I, _P]),
    "isst_path_count": (_I, [_P, C.c_char_p, C.POINTER(C.c_int64)]),
    "isst_debug_option": (_I, [_P, C.c_char_p, _I]),
    "isst_op_decode_attention_bench": (_I, [_P, _I, _I, _I, C.POINTER(C.c_float), _P]),
}

_lib = None


def lib_path() -> str:
    return _build.LIB


def load() -> C.CDLL:
    """Load (building first if needed and possible) the CUDA library.  Raises when it cannot."""
    global _lib
    if _lib is not None:
        return _lib
    if _build.needs_build():
        if shutil.which(os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")) is None and not os.path.exists(_build.LIB):
            raise ImportError("libinfinisst_b200.so is missing and nvcc is not available; run "
                              "`python -c 'import __graft_entry__ as g; g.build()'` on a box with CUDA 12.9. "
                              "infinisst_b200 has no CPU fallback.")
        if shutil.which(os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")) is not None:
            _build.build()
    lib = C.CDLL(_build.LIB)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)          # AttributeError here == the ABI and the header disagree
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class IsstError(RuntimeError):
    pass


def check(status: int) -> None:
    if status != 0:
        raise IsstError(load().isst_last_error().decode("utf-8", "replace"))
