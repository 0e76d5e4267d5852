"""ctypes binding of libavtex.so (the C ABI declared in include/avtex.h).

There is NO fallback: if the shared library is missing or a call fails, this raises.  The product
path never imports `oracle/`.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libavtex.so")
ABI_VERSION = 3

_p = C.c_void_p
_i64 = C.c_int64
_int = C.c_int
_f32 = C.c_float

# name -> argtypes (restype is always int unless noted); mirrors include/avtex.h one to one
SIGNATURES = {
    "avtex_abi_version": [],
    "avtex_device_info": [_int, C.POINTER(_int), C.POINTER(_int)],
    "avtex_zero": [_p, _i64, _int, _p],
    "avtex_pack_frames_u8": [_p, _i64, _i64, _i64, _p, _i64, _p, _p, _int, _p],
    "avtex_pack_frames_f32": [_p, _i64, _i64, _i64, _p, _i64, _p, _p, _p, _int, _p],
    "avtex_gram_l2_s8": [_p, _i64, _i64, _p, _i64, _i64, _int, _p, _i64, _p, _p, _int, _p],
    "avtex_gram_l2_u8": [_p, _i64, _i64, _i64, _p, _i64, _i64, _int, _p, _i64, _p, _p, _int, _p],
    "avtex_frame_norms_u8": [_p, _i64, _i64, _i64, _p, _p, _int, _p],
    "avtex_pairdist_direct_f32": [_p, _i64, _i64, _i64, _i64, _i64, _p, _i64, _p, _p, _int, _p],
    "avtex_pairdist_direct_u8": [_p, _i64, _i64, _i64, _i64, _i64, _p, _i64, _p, _p, _int, _p],
    "avtex_sum_nnz": [_p, _i64, _i64, _i64, _p, _p, _int, _p],
    "avtex_diag_filter_pow": [_p, _i64, _i64, _i64, C.POINTER(_f32), _int, _int, _i64, _i64, _i64,
                              _p, _i64, _p, _i64, _f32, _p, _p, _int, _p],
    "avtex_diag_filter_pow_sym": [_p, _i64, _i64, C.POINTER(_f32), _int, _int, _i64, _p, _i64, _p, _i64, _f32, _p, _p,
                                  _int, _p],
    "avtex_diag_filter_pow_res": [_p, _i64, _i64, _i64, _i64, C.POINTER(_f32), _int, _int, _i64, _i64, _i64, _p, _i64, _p,
                                  _i64, _f32, _p, _p, _int, _int, _p],
    "avtex_future_cost_sweep": [_p, _i64, _i64, _i64, _i64, _p, _p, _f32, _p, _p, _int, _p],
    "avtex_future_cost_fused": [_p, _i64, _i64, _f32, _f32, _int, _p, _i64, _p, _p, _p, _int, _p],
    "avtex_pow_matrix": [_p, _i64, _i64, _i64, _f32, _p, _i64, _int, _p],
    "avtex_frame_norms_u8_push": [_p, _i64, _i64, _i64, _i64, C.POINTER(_p), C.POINTER(_p), _int, _int, _p],
    "avtex_future_cost_finalize": [_p, _i64, _i64, _i64, _i64, _p, _f32, _p, _i64, _p, _p, _int, _p],
    "avtex_transition_probs": [_p, _i64, _i64, _i64, _f32, _int, _i64, _p, _i64, _f32, _p, _i64, _p, _int, _p],
    "avtex_row_nnz": [_p, _i64, _i64, _i64, _p, _int, _p],
    "avtex_csr_fill": [_p, _i64, _i64, _i64, _p, _p, _int, _p],
    "avtex_l2_normalize_rows": [_p, _i64, _i64, _i64, _p, _i64, _int, _p],
    "avtex_cosine_scores": [_p, _i64, _i64, _i64, _p, _f32, _p, _int, _p],
    "avtex_select_step": [_p, _p, _i64, _i64, _f32, _f32, _f32, _p, _p, _p, _int, _p],
    "avtex_audio_start": [_p, _i64, _i64, _i64, _p, _p, _p, _int, _p],
    "avtex_synthesis_loop": [_p, _i64, _i64, _i64, _p, _i64, _p, _i64, _i64, _p, _i64, _f32, _f32, _f32, _f32, _i64, _int,
                             _p, _p, _p, _p, _p, _p, _p, _p, _int, _p],
    "avtex_mt19937_randint_host": [_p, C.POINTER(_int), _p, _int, _p],
    "avtex_synthesis_step": [_p, _i64, _i64, _i64, _p, _p, _i64, _i64, _p, _f32, _i64, _f32, _f32, _f32, _p, _p, _p,
                             _p, _int, _p, _p, _p, _p, _int, _int, _int, _p],
    "avtex_gather_rows": [_p, _i64, _i64, _i64, _p, _i64, _p, _int, _p],
    "avtex_assemble_frames": [_p, _i64, _int, _int, _p, _p, _p, _int, _i64, _p, _int, _p],
    "avtex_logmel": [_p, _i64, _int, _int, _int, _int, _p, _p, _int, C.c_double, _p, _i64, _int, _p],
    "avtex_frame_examples": [_p, _i64, _int, _int, _int, _p, _i64, _int, _p],
    "avtex_gram_tile_schedule": [_int, _int, _int, C.POINTER(_int), C.POINTER(_int), _int],
    "avtex_gram_tile_schedule2": [_int, _int, _int, C.POINTER(_int), C.POINTER(_int), _int],
    "avtex_gram_tile_schedule2g": [_int, _int, _int, _int, C.POINTER(_int), C.POINTER(_int), _int],
}

class GramJob(C.Structure):
    """Mirror of `AvtexGramJob` (include/avtex.h)."""
    _fields_ = [("row0", _i64), ("rows", _i64), ("col0", _i64), ("cols", _i64),
                ("D", _p), ("d_row0", _i64), ("ldd", _i64),
                ("DT", _p), ("dt_row0", _i64), ("ldt", _i64),
                ("symmetric", _int), ("count_stats", _int),
                ("k_off", _i64), ("sq_off", _i64), ("sq_stride", _i64)]


SIGNATURES["avtex_future_cost_fused_peer"] = [_p, _i64, _i64, _i64, _i64, _f32, _f32, _int, _int, _int,
                                              C.POINTER(_p), _i64, C.POINTER(_p), C.POINTER(_p), C.c_uint,
                                              _p, _p, _p, _p, _int, _int, _p]
SIGNATURES["avtex_gram_l2_jobs_fused_norms"] = [_p, _i64, _i64, _i64, _i64, _i64, _p, _p, _p, C.POINTER(GramJob), _int, _p, _p,
                                               _int, _p]
SIGNATURES["avtex_gram_l2_jobs"] = [_p, _int, _i64, _i64, _i64, _p, C.POINTER(GramJob), _int, _p, _p, _p, _int, _p]

_lib = None


class AvtexError(RuntimeError):
    pass


def load():
    """Loads libavtex.so once.  Raises (never falls back) when it is missing or stale."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C audio_video_textures_b200/csrc`.  There is no CPU / PyTorch fallback.")
    lib = C.CDLL(LIB_PATH)
    lib.avtex_last_error.restype = C.c_char_p
    lib.avtex_last_error.argtypes = []
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is not exported
        fn.argtypes = argtypes
        fn.restype = _int
    if lib.avtex_abi_version() != ABI_VERSION:
        raise ImportError(f"libavtex.so ABI {lib.avtex_abi_version()} != expected {ABI_VERSION}: rebuild")
    _lib = lib
    return lib


def call(name, *args):
    lib = load()
    rc = getattr(lib, name)(*args)
    if rc != 0:
        raise AvtexError(f"{name} failed (rc={rc}): {lib.avtex_last_error().decode(errors='replace')}")


def ptr(t):
    """Device/host pointer of a tensor (None -> NULL)."""
    return None if t is None else C.c_void_p(t.data_ptr())
