"""ctypes binding of the C ABI in include/gpusim_b200.h (libgpusim_b200.so, built in-tree by
``make`` / ``__graft_entry__.build()``).  There is no Python or CPU fallback: if the shared
library is missing every entry point raises."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libgpusim_b200.so")

GSB_OK, GSB_ERR_INVALID, GSB_ERR_CUDA, GSB_ERR_NOMEM, GSB_ERR_STATE, GSB_ERR_CORRUPT, GSB_ERR_IO = range(7)


class GsbError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"gpusim_b200 error {code}: {msg}")
        self.code = code


class Exchange(C.Structure):
    _fields_ = [("peer_base", C.c_uint64 * 16), ("rank", C.c_uint32), ("world", C.c_uint32), ("seq", C.c_uint64)]


class Sink(C.Structure):
    """gsb_sink: device-accessible result pointers of gsb_db_search_enqueue."""
    _fields_ = [("keys", C.c_void_p), ("rows", C.c_void_p), ("scores", C.c_void_p), ("n", C.c_void_p),
                ("approx", C.c_void_p), ("done", C.c_void_p), ("done_value", C.c_uint64)]


GSB_QUERY_STABLE = 1
GSB_COUNT_ERROR = 0xFFFFFFFF
GSB_METRIC_TANIMOTO, GSB_METRIC_DICE, GSB_METRIC_TVERSKY = 0, 1, 2
GSB_BATCH_LOOPED, GSB_BATCH_POPC, GSB_BATCH_SLICED = 0, 1, 2


class ScanInfo(C.Structure):
    _fields_ = [("device", C.c_int), ("grid", C.c_int), ("block", C.c_int), ("stages", C.c_int),
                ("tile_rows", C.c_uint32), ("tile_bytes", C.c_uint32), ("smem_bytes", C.c_uint32),
                ("cand_capacity", C.c_uint32), ("shard_rows", C.c_uint64),
                ("db_bytes_per_query", C.c_uint64)]


# name -> (restype, argtypes); must list every symbol the header declares (tests check it)
_P = C.c_void_p
SIGNATURES = {
    "gsb_last_error": (C.c_char_p, []),
    "gsb_version": (C.c_char_p, []),
    "gsb_device_count": (C.c_int, []),
    "gsb_device_free_bytes": (C.c_uint64, [C.c_int]),
    "gsb_available_device_bytes": (C.c_uint64, []),
    "gsb_next_device": (C.c_int, [C.c_uint64, C.POINTER(C.c_int)]),
    "gsb_devices_reset": (C.c_int, []),
    "gsb_layout_bytes": (C.c_uint64, [C.c_int, C.c_uint64, C.c_uint]),
    "gsb_db_create": (C.c_int, [C.POINTER(_P), C.POINTER(C.c_uint64), C.c_int, C.c_int, C.c_uint64,
                                C.POINTER(_P)]),
    "gsb_db_create_synthetic": (C.c_int, [C.c_int, C.c_int, C.c_uint64, C.c_uint64, C.c_uint64,
                                          C.c_uint32, C.POINTER(_P)]),
    "gsb_db_create_synthetic_sharded": (C.c_int, [C.POINTER(C.c_int), C.c_int, C.c_int, C.c_uint64, C.c_uint64,
                                                  C.c_uint64, C.c_uint32, C.POINTER(_P)]),
    "gsb_db_upload": (C.c_int, [_P, C.POINTER(C.c_int), C.c_int, C.c_uint]),
    "gsb_db_set_metric": (C.c_int, [_P, C.c_int, C.c_float, C.c_float]),
    "gsb_db_destroy": (None, [_P]),
    "gsb_db_count": (C.c_uint64, [_P]),
    "gsb_db_fp_bits": (C.c_int, [_P]),
    "gsb_db_data_bytes": (C.c_uint64, [_P]),
    "gsb_db_fold_factor": (C.c_uint, [_P]),
    "gsb_db_shard_count": (C.c_int, [_P]),
    "gsb_db_get_fingerprint": (C.c_int, [_P, C.c_uint64, _P]),
    "gsb_db_search": (C.c_int, [_P, _P, C.c_int, C.c_uint32, C.c_float, _P, _P,
                                C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)]),
    "gsb_db_search_async": (C.c_int, [_P, _P, C.c_int, C.c_uint32, C.c_float, C.POINTER(C.c_uint64)]),
    "gsb_db_search_wait": (C.c_int, [_P, C.c_uint64, _P, _P, C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)]),
    "gsb_db_search_batch": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_uint32, C.c_float, _P, _P, _P, _P]),
    "gsb_db_batch_mode": (C.c_int, [_P, C.c_uint32, C.c_int, C.c_float, C.POINTER(C.c_int), C.POINTER(C.c_uint32)]),
    "gsb_db_search_cpu": (C.c_int, [_P, _P, C.c_int, C.c_uint32, _P, _P, C.POINTER(C.c_uint32)]),
    "gsb_db_search_device": (C.c_int, [_P, _P, _P, C.c_uint32, C.c_float, _P, _P, _P]),
    "gsb_db_search_batch_device": (C.c_int, [_P, _P, _P, C.c_int, C.c_uint32, C.c_float, _P, _P, _P]),
    "gsb_db_batch_max_queries": (C.c_int, [_P, C.c_uint32, C.c_int, C.c_float, C.POINTER(C.c_uint32)]),
    "gsb_merge_batch_device": (C.c_int, [C.c_int, _P, _P, C.c_int, C.c_int, C.c_uint32, _P, _P, _P, _P]),
    "gsb_exchange_bytes": (C.c_int, [C.c_uint32, C.c_uint32, C.POINTER(C.c_uint64)]),
    "gsb_db_search_device_fused": (C.c_int, [_P, _P, _P, C.c_uint32, C.c_float, C.POINTER(Exchange), _P, _P, _P, _P]),
    "gsb_db_search_enqueue": (C.c_int, [_P, _P, _P, _P, C.c_uint32, C.c_uint32, C.c_float, C.POINTER(Exchange),
                                        C.POINTER(Sink)]),
    "gsb_wait_word": (C.c_int, [_P, C.c_uint64, C.c_uint64]),
    "gsb_merge_device": (C.c_int, [C.c_int, _P, _P, _P, C.c_int, C.c_uint32, C.c_uint32, _P, _P, _P]),
    "gsb_fsim_open": (C.c_int, [C.c_char_p, C.POINTER(_P)]),
    "gsb_fsim_close": (None, [_P]),
    "gsb_fsim_last_error": (C.c_char_p, []),
    "gsb_fsim_dbkey": (C.c_char_p, [_P]),
    "gsb_fsim_fp_bits": (C.c_int, [_P]),
    "gsb_fsim_fp_count": (C.c_uint64, [_P]),
    "gsb_fsim_chunk_count": (C.c_int, [_P]),
    "gsb_fsim_chunk_data": (_P, [_P, C.c_int]),
    "gsb_fsim_chunk_bytes": (C.c_uint64, [_P, C.c_int]),
    "gsb_fsim_string_count": (C.c_uint64, [_P, C.c_int]),
    "gsb_fsim_string": (C.c_char_p, [_P, C.c_int, C.c_uint64]),
    "gsb_fsim_create_db": (C.c_int, [_P, C.POINTER(_P)]),
    "gsb_server_create": (C.c_int, [C.POINTER(C.c_char_p), C.c_int, C.c_int, C.c_int, C.POINTER(_P)]),
    "gsb_server_destroy": (None, [_P]),
    "gsb_server_recover": (C.c_int, [_P]),
    "gsb_server_last_error": (C.c_char_p, []),
    "gsb_server_set_use_gpu": (None, [_P, C.c_int]),
    "gsb_server_using_gpu": (C.c_int, [_P]),
    "gsb_server_fold_factor": (C.c_uint, [_P]),
    "gsb_server_database_count": (C.c_int, [_P]),
    "gsb_server_get_fingerprint": (C.c_int, [_P, C.c_char_p, C.c_uint64, _P]),
    "gsb_server_handle_request": (C.c_int, [_P, _P, C.c_uint64, C.POINTER(_P), C.POINTER(C.c_uint64)]),
    "gsb_server_free": (None, [_P]),
    "gsb_server_handle_batch": (C.c_int, [_P, C.POINTER(_P), C.POINTER(C.c_uint64), C.c_int, C.POINTER(_P),
                                          C.POINTER(C.c_uint64)]),
    "gsb_server_listen": (C.c_int, [_P, C.c_char_p]),
    "gsb_server_serve": (C.c_int, [_P, C.c_uint64]),
    "gsb_server_stop": (None, [_P]),
    "gsb_fold_fingerprint": (C.c_int, [_P, C.c_int, C.c_int, _P]),
    "gsb_db_scan_info": (C.c_int, [_P, C.c_int, C.c_uint32, C.POINTER(ScanInfo)]),
    "gsb_selftest_division": (C.c_int, [C.c_int, C.POINTER(C.c_uint64)]),
    "gsb_launch_count": (C.c_uint64, []),
}

_lib = None


def lib() -> C.CDLL:
    """The loaded library.  Raises ImportError (loudly) when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `make` or `python -c 'import __graft_entry__ as g; "
                "g.build()'`.  gpusimilarity_b200 has no fallback path.")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(rc: int) -> None:
    if rc != GSB_OK:
        raise GsbError(rc, lib().gsb_last_error().decode())
