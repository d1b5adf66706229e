"""ctypes binding of include/mvs_ncc.h.  Fails loudly when the library is missing:
there is no CPU or PyTorch fallback for any entry point."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libmvsncc.so")
ABI_VERSION = 1

# every symbol include/mvs_ncc.h declares: name -> (restype, argtypes)
_P = C.c_void_p
SYMBOLS = {
    "mvs_abi_version": (C.c_int, []),
    "mvs_last_error": (C.c_char_p, []),
    "mvs_create": (C.c_int, [C.POINTER(_P), C.c_int, C.c_int, C.c_int, C.c_int, _P, C.c_int, _P, _P, _P, _P]),
    "mvs_destroy": (C.c_int, [_P]),
    "mvs_get_info": (C.c_int, [_P, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int64)]),
    "mvs_download_gray": (C.c_int, [_P, _P]),
    "mvs_get_cameras": (C.c_int, [_P, _P, _P]),
    "mvs_score_batch": (C.c_int, [_P, C.c_int, C.c_int64, _P, _P, _P, C.c_double, C.c_int, _P, _P, _P, _P, _P,
                                  C.c_int, _P]),
    "mvs_compact_accepted_p2p": (C.c_int, [_P, C.c_int64, C.c_int64, _P, _P, _P, _P, _P, _P, _P, _P, C.c_int, _P, _P, C.c_int,
                                           C.c_int, C.c_int, C.c_int64, _P]),
    "mvs_wire_bytes": (C.c_int, [_P, C.c_int]),
    "mvs_records_expand": (C.c_int, [_P, C.c_int, _P, C.c_int64, _P, _P]),
    "mvs_score_pmvs": (C.c_int, [_P, C.c_int64, _P, _P, _P, _P, C.c_double, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P, _P,
                                 _P, _P, _P, _P, C.c_int, _P]),
    "mvs_select_best": (C.c_int, [_P, C.c_int64, C.c_int, _P, _P, C.c_int, _P, _P, _P]),
    "mvs_launch_count": (C.c_int64, [_P]),
    "mvs_profile_enable": (C.c_int, [_P, C.c_int]),
    "mvs_profile_probe": (C.c_int, [_P, C.c_int]),
    "mvs_profile_score_ms": (C.c_int, [_P, C.POINTER(C.c_float), C.POINTER(C.c_int)]),
    "mvs_record_bytes": (C.c_int, [_P]),
    "mvs_cells_init": (C.c_int, [_P, C.c_int, _P]),
    "mvs_cells_shape": (C.c_int, [_P, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "mvs_cells_download": (C.c_int, [_P, _P]),
    "mvs_cells_fill": (C.c_int, [_P, _P, C.c_int64, _P]),
    "mvs_round_generate": (C.c_int, [_P, _P, C.c_int64, C.POINTER(C.c_int64), _P]),
    "mvs_round_score": (C.c_int, [_P, _P, C.c_int64, C.c_int64, C.c_double, C.c_int, C.c_int, C.c_double, _P, C.c_int64,
                                  _P, _P]),
    "mvs_round_score_p2p": (C.c_int, [_P, _P, C.c_int64, C.c_int64, C.c_double, C.c_int, C.c_int, C.c_double, _P, _P, C.c_int,
                                      C.c_int, C.c_int, C.c_int64, _P]),
    "mvs_round_commit": (C.c_int, [_P, _P, C.c_int64, _P, _P, _P]),
    "mvs_round_candidates": (C.c_int, [_P, _P, _P, _P, _P, _P]),
    "mvs_exchange_bytes": (C.c_int64, [_P, C.c_int, C.c_int64]),
    "mvs_publish_accepted": (C.c_int, [_P, C.c_int64, _P, _P, _P, _P, C.c_int, _P, C.c_int, C.c_int, C.c_int64, C.c_int, _P]),
    "mvs_exchange_set_parts": (C.c_int, [_P, C.c_int, C.c_int64]),
    "mvs_score_publish": (C.c_int, [_P, C.c_int64, _P, _P, C.c_double, C.c_int, _P, _P, _P, _P, _P, C.c_int, _P, C.c_int, C.c_int,
                                    C.c_int64, C.c_int, _P]),
    "mvs_p2p_barrier": (C.c_int, [_P, _P, C.c_int, C.c_int, _P]),
    "mvs_p2p_barrier_failed": (C.c_int, [_P, _P]),
    "mvs_expand_run": (C.c_int, [_P, _P, C.c_int64, _P, _P, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int64), _P]),
    "mvs_expand_result": (C.c_int, [_P, _P, C.c_int64, C.c_int64, C.c_int, _P]),
    "mvs_cells_filter": (C.c_int, [_P, _P, C.c_int64, _P, _P, _P]),
    "mvs_triangulate": (C.c_int, [_P, C.c_int64, _P, _P, _P, _P, _P, _P]),
    "mvs_seed_stage": (C.c_int, [_P, C.c_int64, _P, _P, _P, C.c_double, C.c_int, C.c_int, _P, C.c_void_p, _P]),
    "mvs_ncc_pairs": (C.c_int, [C.c_int, C.c_int64, C.c_int, _P, _P, _P, C.c_int, _P]),
    "mvs_compact_accepted": (C.c_int, [_P, C.c_int64, C.c_int64, _P, _P, _P, _P, _P, _P, _P, _P, C.c_int, _P, C.c_int64,
                                       _P, _P]),
}



class ExpandParams(C.Structure):
    """mvs_expand_params (include/mvs_ncc.h)."""
    _fields_ = [("min_ncc", C.c_double), ("scale", C.c_double), ("wid", C.c_int32), ("bound", C.c_int32),
                ("max_rounds", C.c_int64), ("max_iterations", C.c_int64), ("max_patches", C.c_int64),
                ("rank", C.c_int32), ("world", C.c_int32), ("peer_inbox", C.c_void_p), ("peer_flags", C.c_void_p),
                ("capacity", C.c_int64), ("timing", C.c_int32), ("reserved", C.c_int32)]


class RoundStat(C.Structure):
    """mvs_round_stat (include/mvs_ncc.h)."""
    _fields_ = [("frontier", C.c_int64), ("candidates", C.c_int64), ("passed", C.c_int64), ("accepted", C.c_int64),
                ("ms", C.c_float), ("reserved", C.c_int32)]


_lib = None


def load():
    """dlopen libmvsncc.so and bind every declared symbol (no compute is run)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python __graft_entry__.py` (nvcc, sm_100a). "
            "There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)          # AttributeError if the .so lacks a declared symbol
        fn.restype = res
        fn.argtypes = args
    if lib.mvs_abi_version() != ABI_VERSION:
        raise ImportError(f"libmvsncc.so has ABI {lib.mvs_abi_version()}, binding expects {ABI_VERSION}: rebuild")
    _lib = lib
    return lib
