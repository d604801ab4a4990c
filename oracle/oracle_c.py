"""ctypes bindings for the C oracle (oracle/_build/liboracle.so) and for the reference's own
sources compiled verbatim (oracle/_ref/libgpusim_ref.so).  TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional, Tuple

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_ORACLE_SO = os.path.join(HERE, "_build", "liboracle.so")
_REF_SO = os.path.join(HERE, "_ref", "libgpusim_ref.so")


def build(with_ref: bool = True) -> None:
    """Compile the checkers (``make -C oracle``).  The reference-based one is rebuilt only
    where /root/reference exists; elsewhere the prebuilt file is used as is."""
    subprocess.run(["make", "-s", "-C", HERE, "all" if with_ref else "_build/liboracle.so"],
                   check=True)


_oracle = None
_ref = None


def oracle_lib():
    global _oracle
    if _oracle is None:
        if not os.path.exists(_ORACLE_SO):
            build(with_ref=False)
        lib = C.CDLL(_ORACLE_SO)
        lib.oracle_score.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_uint64, C.c_void_p, C.c_int]
        lib.oracle_score.restype = None
        lib.oracle_search.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_uint64, C.c_uint64,
                                      C.c_uint32, C.c_float, C.c_void_p, C.c_void_p,
                                      C.POINTER(C.c_uint32), C.POINTER(C.c_uint64), C.c_int]
        lib.oracle_search.restype = None
        lib.oracle_fold.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        lib.oracle_fold.restype = None
        lib.oracle_synth_rows.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_int, C.c_uint32,
                                          C.c_void_p, C.c_int]
        lib.oracle_synth_rows.restype = None
        lib.oracle_stream_search.argtypes = [C.c_void_p, C.c_int, C.c_uint64, C.c_uint32, C.c_uint64,
                                             C.c_uint64, C.c_uint32, C.c_float, C.c_void_p, C.c_void_p,
                                             C.POINTER(C.c_uint32), C.POINTER(C.c_uint64), C.c_int]
        lib.oracle_stream_search.restype = None
        lib.oracle_stream_search_multi.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_uint64, C.c_uint32, C.c_uint64,
                                                   C.c_uint64, C.c_uint32, C.c_float, C.c_void_p, C.c_void_p,
                                                   C.c_void_p, C.c_void_p, C.c_int]
        lib.oracle_stream_search_multi.restype = None
        _oracle = lib
    return _oracle


def _i32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.int32)


def c_score(query, db, n_threads: int = 0) -> np.ndarray:
    q, d = _i32(query), _i32(db)
    out = np.empty(d.shape[0], dtype=np.float32)
    oracle_lib().oracle_score(q.ctypes.data, q.shape[0], d.ctypes.data, d.shape[0],
                              out.ctypes.data, n_threads or (os.cpu_count() or 1))
    return out


def c_search(query, db, k: int, cutoff: float, row_base: int = 0,
             n_threads: int = 0) -> Tuple[np.ndarray, np.ndarray, int]:
    q, d = _i32(query), _i32(db)
    rows = np.empty(max(k, 1), dtype=np.uint32)
    scores = np.empty(max(k, 1), dtype=np.float32)
    n, approx = C.c_uint32(0), C.c_uint64(0)
    oracle_lib().oracle_search(q.ctypes.data, q.shape[0], d.ctypes.data, d.shape[0], row_base, k,
                               cutoff, rows.ctypes.data, scores.ctypes.data, C.byref(n),
                               C.byref(approx), n_threads or (os.cpu_count() or 1))
    return rows[:n.value].astype(np.int64), scores[:n.value], int(approx.value)


def c_synth_db(seed: int, n_rows: int, words: int = 32, plant_period: int = 0, row_base: int = 0,
               n_threads: int = 0) -> np.ndarray:
    """Threaded C twin of oracle.synth_db (checked against it in tests/test_oracle.py)."""
    out = np.empty((n_rows, words), dtype=np.int32)
    oracle_lib().oracle_synth_rows(seed, n_rows, row_base, words, plant_period, out.ctypes.data,
                                   n_threads or (os.cpu_count() or 1))
    return out


def c_stream_search(query, seed: int, plant_period: int, n_rows: int, k: int, cutoff: float,
                    row_base: int = 0, n_threads: int = 0) -> Tuple[np.ndarray, np.ndarray, int]:
    """oracle_search over rows [row_base, row_base + n_rows) of the synthetic database, generated on
    the fly inside the scan (nothing is materialised): the full-size oracle for 1 B-row searches."""
    q = _i32(query)
    rows = np.empty(max(k, 1), dtype=np.uint32)
    scores = np.empty(max(k, 1), dtype=np.float32)
    n, approx = C.c_uint32(0), C.c_uint64(0)
    oracle_lib().oracle_stream_search(q.ctypes.data, q.shape[0], seed, plant_period, n_rows, row_base, k,
                                      cutoff, rows.ctypes.data, scores.ctypes.data, C.byref(n),
                                      C.byref(approx), n_threads or (os.cpu_count() or 1))
    return rows[:n.value].astype(np.int64), scores[:n.value], int(approx.value)


def c_stream_search_multi(queries, seed: int, plant_period: int, n_rows: int, k: int, cutoff: float,
                          row_base: int = 0, n_threads: int = 0):
    """c_stream_search for a batch of queries in one pass over the generated rows.  Returns a list of
    (rows, scores, approx), one per query."""
    qs = np.ascontiguousarray(queries, dtype=np.int32)
    nq = qs.shape[0]
    rows = np.zeros((nq, max(k, 1)), dtype=np.uint32)
    scores = np.zeros((nq, max(k, 1)), dtype=np.float32)
    n = np.zeros(nq, dtype=np.uint32)
    approx = np.zeros(nq, dtype=np.uint64)
    oracle_lib().oracle_stream_search_multi(qs.ctypes.data, nq, qs.shape[1], seed, plant_period, n_rows, row_base, k,
                                            cutoff, rows.ctypes.data, scores.ctypes.data, n.ctypes.data,
                                            approx.ctypes.data, n_threads or (os.cpu_count() or 1))
    return [(rows[i, :n[i]].astype(np.int64), scores[i, :n[i]].copy(), int(approx[i])) for i in range(nq)]


def c_fold(fp, factor: int) -> np.ndarray:
    f = _i32(fp)
    out = np.empty(f.shape[0] // factor, dtype=np.int32)
    oracle_lib().oracle_fold(f.ctypes.data, f.shape[0], factor, out.ctypes.data)
    return out


# --------------------------------------------------------------------------- reference itself
def ref_available() -> bool:
    return os.path.exists(_REF_SO)


def ref_lib():
    global _ref
    if _ref is None:
        lib = C.CDLL(_REF_SO)
        lib.ref_last_error.restype = C.c_char_p
        lib.ref_gpu_count.restype = C.c_int
        lib.ref_db_create.argtypes = [C.c_int, C.c_int, C.c_char_p, C.POINTER(C.c_void_p),
                                      C.POINTER(C.c_uint64), C.c_int]
        lib.ref_db_create.restype = C.c_void_p
        lib.ref_db_destroy.argtypes = [C.c_void_p]
        lib.ref_db_copy_to_gpu.argtypes = [C.c_void_p, C.c_uint]
        lib.ref_db_search.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_char_p, C.c_uint,
                                      C.c_float, C.c_int, C.c_void_p, C.c_void_p,
                                      C.POINTER(C.c_ulong)]
        lib.ref_db_search.restype = C.c_int
        lib.ref_db_get_fingerprint.argtypes = [C.c_void_p, C.c_uint, C.c_void_p, C.c_int]
        lib.ref_bubble_sort.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        lib.ref_fold.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        lib.ref_score_cpu.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        _ref = lib
    return _ref


class RefDB:
    """The reference's FingerprintDB (compiled from its own sources) behind row-number results."""

    APPROX_UNSET = 0xDEAD0000DEAD

    def __init__(self, chunks, fp_bitcount: int, dbkey: str = "pass"):
        self._chunks = [np.ascontiguousarray(c) for c in chunks]
        self.words = fp_bitcount // 32
        n = sum(c.nbytes for c in self._chunks) // (fp_bitcount // 8)
        ptrs = (C.c_void_p * len(self._chunks))(*[c.ctypes.data for c in self._chunks])
        sizes = (C.c_uint64 * len(self._chunks))(*[c.nbytes for c in self._chunks])
        self._key = dbkey.encode()
        self._h = ref_lib().ref_db_create(fp_bitcount, n, self._key, ptrs, sizes, len(self._chunks))
        if not self._h:
            raise RuntimeError(ref_lib().ref_last_error().decode())
        self.count = n

    def copy_to_gpu(self, fold_factor: int = 1) -> None:
        if ref_lib().ref_db_copy_to_gpu(self._h, fold_factor) != 0:
            raise RuntimeError(ref_lib().ref_last_error().decode())

    def search(self, query, k: int, cutoff: float, cpu: bool = False,
               dbkey: Optional[str] = None):
        q = _i32(query)
        rows = np.empty(max(k, 1) + 8, dtype=np.uint32)
        scores = np.empty(max(k, 1) + 8, dtype=np.float32)
        approx = C.c_ulong(self.APPROX_UNSET)
        n = ref_lib().ref_db_search(self._h, q.ctypes.data, q.shape[0],
                                    (dbkey.encode() if dbkey is not None else self._key), k,
                                    cutoff, 1 if cpu else 0, rows.ctypes.data, scores.ctypes.data,
                                    C.byref(approx))
        if n < 0:
            raise RuntimeError(ref_lib().ref_last_error().decode())
        return rows[:n].astype(np.int64), scores[:n].copy(), int(approx.value)

    def get_fingerprint(self, row: int) -> np.ndarray:
        out = np.empty(self.words, dtype=np.int32)
        ref_lib().ref_db_get_fingerprint(self._h, row, out.ctypes.data, self.words)
        return out

    def close(self):
        if self._h:
            ref_lib().ref_db_destroy(self._h)
            self._h = None

    def __del__(self):
        self.close()


def ref_bubble_sort(indices, scores, number_required: int):
    i = np.ascontiguousarray(indices, dtype=np.int32).copy()
    s = np.ascontiguousarray(scores, dtype=np.float32).copy()
    ref_lib().ref_bubble_sort(i.ctypes.data, s.ctypes.data, i.shape[0], number_required)
    return i, s


def ref_fold(fp, factor: int) -> np.ndarray:
    f = _i32(fp)
    out = np.zeros(f.shape[0] // factor, dtype=np.int32)
    ref_lib().ref_fold(f.ctypes.data, f.shape[0], factor, out.ctypes.data)
    return out


def ref_score_cpu(query, db, n_threads: int = 0) -> np.ndarray:
    q, d = _i32(query), _i32(db)
    out = np.empty(d.shape[0], dtype=np.float32)
    ref_lib().ref_score_cpu(q.ctypes.data, q.shape[0], d.ctypes.data, d.shape[0], out.ctypes.data,
                            n_threads or (os.cpu_count() or 1))
    return out
