"""Host-side mirror of the reference's ``gpusim::FingerprintDB`` (fingerprintdb_cuda.h:53-147)
over the C ABI: same method names, argument meaning and error behaviour, so that the parity
tests read like reference test/test_gpusim.cpp.  All compute happens in libgpusim_b200.so."""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import _lib
from ._lib import GsbError, ScanInfo, check, lib


def get_gpu_count() -> int:
    """reference get_gpu_count(), fingerprintdb_cuda.cu:41-52."""
    return lib().gsb_device_count()


def get_next_gpu(required_memory: int) -> int:
    """reference get_next_gpu(), fingerprintdb_cuda.cu:54-68 (raises like its runtime_error)."""
    dev = C.c_int(0)
    check(lib().gsb_next_device(required_memory, C.byref(dev)))
    return dev.value


def get_available_gpu_memory() -> int:
    """reference get_available_gpu_memory(), fingerprintdb_cuda.cu:401-413."""
    return lib().gsb_available_device_bytes()


def top_results_bubble_sort(indices: List[int], scores: List[float], number_required: int) -> None:
    """reference top_results_bubble_sort, fingerprintdb_cuda.cpp:92-103 (in place)."""
    count = len(indices)
    for i in range(number_required):
        for j in range(count - 1, i, -1):
            if scores[j] > scores[j - 1]:
                indices[j], indices[j - 1] = indices[j - 1], indices[j]
                scores[j], scores[j - 1] = scores[j - 1], scores[j]


def fold_fingerprint(fp: Sequence[int], factor: int) -> np.ndarray:
    """reference FoldFingerprintFunctorCPU, calculation_functors.cpp:22-41."""
    words = np.ascontiguousarray(fp, dtype=np.int32)
    out = np.empty(words.shape[0] // max(factor, 1), dtype=np.int32)
    check(lib().gsb_fold_fingerprint(words.ctypes.data, words.shape[0], factor, out.ctypes.data))
    return out


def _as_i32(q) -> np.ndarray:
    return np.ascontiguousarray(q, dtype=np.int32)


class FingerprintDB:
    """``FingerprintDB(fp_bitcount, fp_count, dbkey, data, smiles_vector, ids_vector)``.

    ``data`` is the list of raw fingerprint chunks (bytes or arrays) exactly as
    ``GPUSimServer::extractData`` hands them over (gpusim.cpp:101-112).  Like the reference
    constructor (.cu:164-165) the smiles / ids lists are *taken*: the caller's lists are emptied.
    """

    def __init__(self, fp_bitcount: int, fp_count: int, dbkey: str, data: Sequence,
                 smiles_vector: Optional[list] = None, ids_vector: Optional[list] = None):
        self._h = C.c_void_p()
        chunks = [np.frombuffer(c, dtype=np.uint8) if isinstance(c, (bytes, bytearray, memoryview))
                  else np.ascontiguousarray(c).view(np.uint8).reshape(-1) for c in data]
        ptrs = (C.c_void_p * max(len(chunks), 1))(*[c.ctypes.data for c in chunks])
        sizes = (C.c_uint64 * max(len(chunks), 1))(*[c.nbytes for c in chunks])
        check(lib().gsb_db_create(ptrs, sizes, len(chunks), fp_bitcount, fp_count, C.byref(self._h)))
        self.m_dbkey = dbkey
        self.m_smiles: list = []
        self.m_ids: list = []
        if smiles_vector is not None:
            self.m_smiles, smiles_vector[:] = list(smiles_vector), []
        if ids_vector is not None:
            self.m_ids, ids_vector[:] = list(ids_vector), []

    @classmethod
    def synthetic(cls, n_rows: int, device: int = 0, fp_bitcount: int = 1024, row_base: int = 0,
                  seed: int = 0x5EED5EED, plant_period: int = 0, dbkey: str = "pass") -> "FingerprintDB":
        """A shard generated in device memory (already on the GPU); rows have no SMILES / ids."""
        self = cls.__new__(cls)
        self._h = C.c_void_p()
        check(lib().gsb_db_create_synthetic(device, fp_bitcount, n_rows, row_base, seed, plant_period,
                                            C.byref(self._h)))
        self.m_dbkey, self.m_smiles, self.m_ids = dbkey, [], []
        return self

    @classmethod
    def synthetic_sharded(cls, n_rows: int, devices: Sequence[int], fp_bitcount: int = 1024, row_base: int = 0,
                          seed: int = 0x5EED5EED, plant_period: int = 0, dbkey: str = "pass") -> "FingerprintDB":
        """The same synthetic rows as contiguous, equal shards over ``devices``, driven by this ONE
        process (the reference's own multi-GPU mode, fingerprintdb_cuda.cu:176-183)."""
        self = cls.__new__(cls)
        self._h = C.c_void_p()
        arr = (C.c_int * len(devices))(*devices)
        check(lib().gsb_db_create_synthetic_sharded(arr, len(devices), fp_bitcount, n_rows, row_base, seed,
                                                    plant_period, C.byref(self._h)))
        self.m_dbkey, self.m_smiles, self.m_ids = dbkey, [], []
        return self

    def setMetric(self, metric: str = "tanimoto", alpha: float = 1.0, beta: float = 1.0) -> None:
        """Similarity metric of later searches: 'tanimoto' (the reference's), 'dice', 'tversky'."""
        code = {"tanimoto": _lib.GSB_METRIC_TANIMOTO, "dice": _lib.GSB_METRIC_DICE,
                "tversky": _lib.GSB_METRIC_TVERSKY}[metric]
        check(lib().gsb_db_set_metric(self._h, code, alpha, beta))

    # -- life cycle ---------------------------------------------------------------------
    def copyToGPU(self, fold_factor: int = 1, devices: Optional[Sequence[int]] = None) -> None:
        if devices:
            arr = (C.c_int * len(devices))(*devices)
            check(lib().gsb_db_upload(self._h, arr, len(devices), fold_factor))
        else:
            check(lib().gsb_db_upload(self._h, None, 0, fold_factor))

    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h:
            lib().gsb_db_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- accessors ----------------------------------------------------------------------
    def count(self) -> int:
        return lib().gsb_db_count(self._h)

    def getFingerprintDataSize(self) -> int:
        return lib().gsb_db_data_bytes(self._h)

    def getFingerprintBitcount(self) -> int:
        return lib().gsb_db_fp_bits(self._h)

    def foldFactor(self) -> int:
        return lib().gsb_db_fold_factor(self._h)

    def shardCount(self) -> int:
        return lib().gsb_db_shard_count(self._h)

    def getFingerprint(self, index: int) -> np.ndarray:
        out = np.empty(self.getFingerprintBitcount() // 32, dtype=np.int32)
        check(lib().gsb_db_get_fingerprint(self._h, index, out.ctypes.data))
        return out

    def getSmiles(self, index: int):
        return self.m_smiles[index]

    def getID(self, index: int):
        return self.m_ids[index]

    def scan_info(self, k: int, shard: int = 0) -> ScanInfo:
        info = ScanInfo()
        check(lib().gsb_db_scan_info(self._h, shard, k, C.byref(info)))
        return info

    # -- search -------------------------------------------------------------------------
    def search_rows(self, query, max_return_count: int, similarity_cutoff: float
                    ) -> Tuple[np.ndarray, np.ndarray, int]:
        """Row-level result of FingerprintDB::search: (global rows, f32 scores, approximate count)."""
        q = _as_i32(query)
        k = int(max_return_count)
        rows = np.empty(max(k, 1), dtype=np.uint32)
        scores = np.empty(max(k, 1), dtype=np.float32)
        n, approx = C.c_uint32(0), C.c_uint64(0)
        check(lib().gsb_db_search(self._h, q.ctypes.data, q.shape[0], k, similarity_cutoff,
                                  rows.ctypes.data, scores.ctypes.data, C.byref(n), C.byref(approx)))
        return rows[:n.value].astype(np.int64), scores[:n.value].copy(), int(approx.value)

    def search_rows_async(self, query, max_return_count: int, similarity_cutoff: float) -> int:
        """Queue one search (gsb_db_search_async); returns the ticket for ``search_rows_wait``.  Up to
        four searches per database may be in flight; consecutive ones overlap on the device."""
        q = _as_i32(query)
        ticket = C.c_uint64(0)
        check(lib().gsb_db_search_async(self._h, q.ctypes.data, q.shape[0], int(max_return_count), similarity_cutoff,
                                        C.byref(ticket)))
        self._pending_k = getattr(self, "_pending_k", {})
        self._pending_k[ticket.value] = int(max_return_count)
        return ticket.value

    def search_rows_wait(self, ticket: int) -> Tuple[np.ndarray, np.ndarray, int]:
        k = self._pending_k.pop(ticket)
        rows = np.empty(max(k, 1), dtype=np.uint32)
        scores = np.empty(max(k, 1), dtype=np.float32)
        n, approx = C.c_uint32(0), C.c_uint64(0)
        check(lib().gsb_db_search_wait(self._h, ticket, rows.ctypes.data, scores.ctypes.data, C.byref(n),
                                       C.byref(approx)))
        return rows[:n.value].astype(np.int64), scores[:n.value].copy(), int(approx.value)

    def batch_mode(self, max_return_count: int, n_queries: int, similarity_cutoff: float = 0.0) -> Tuple[int, int]:
        """(GSB_BATCH_* mode, queries per pass) gsb_db_search_batch would use."""
        mode, per = C.c_int(0), C.c_uint32(0)
        check(lib().gsb_db_batch_mode(self._h, int(max_return_count), n_queries, similarity_cutoff, C.byref(mode),
                                      C.byref(per)))
        return mode.value, per.value

    def search(self, query, dbkey: str, max_return_count: int, similarity_cutoff: float,
               results_smiles: list, results_ids: list, results_scores: list) -> Optional[int]:
        """reference FingerprintDB::search (.cu:341-381): results are APPENDED to the three lists;
        returns the approximate result count, or None (and appends nothing) on a dbkey mismatch —
        the reference leaves its out-parameter untouched in that case (.cu:349-352)."""
        if dbkey != self.m_dbkey:
            return None
        rows, scores, approx = self.search_rows(query, max_return_count, similarity_cutoff)
        for r, s in zip(rows, scores):
            results_smiles.append(self.m_smiles[r] if self.m_smiles else int(r))
            results_ids.append(self.m_ids[r] if self.m_ids else int(r))
            results_scores.append(float(s))
        return approx

    def search_batch_rows(self, queries, max_return_count: int, similarity_cutoff: float):
        q = np.ascontiguousarray(queries, dtype=np.int32)
        nq, k = q.shape[0], int(max_return_count)
        rows = np.zeros((nq, max(k, 1)), dtype=np.uint32)
        scores = np.zeros((nq, max(k, 1)), dtype=np.float32)
        n = np.zeros(nq, dtype=np.uint32)
        approx = np.zeros(nq, dtype=np.uint64)
        check(lib().gsb_db_search_batch(self._h, q.ctypes.data, q.shape[1], nq, k, similarity_cutoff,
                                        rows.ctypes.data, scores.ctypes.data, n.ctypes.data,
                                        approx.ctypes.data))
        return [(rows[i, :n[i]].astype(np.int64), scores[i, :n[i]].copy(), int(approx[i])) for i in range(nq)]

    def search_batch_rows_raw(self, queries, max_return_count: int, similarity_cutoff: float):
        """gsb_db_search_batch with the result arrays as they come: (rows [nq][k] int64, scores [nq][k],
        n [nq], approx [nq])."""
        q = np.ascontiguousarray(queries, dtype=np.int32)
        nq, k = q.shape[0], int(max_return_count)
        rows = np.zeros((nq, max(k, 1)), dtype=np.uint32)
        scores = np.zeros((nq, max(k, 1)), dtype=np.float32)
        n = np.zeros(nq, dtype=np.uint32)
        approx = np.zeros(nq, dtype=np.uint64)
        check(lib().gsb_db_search_batch(self._h, q.ctypes.data, q.shape[1], nq, k, similarity_cutoff,
                                        rows.ctypes.data, scores.ctypes.data, n.ctypes.data,
                                        approx.ctypes.data))
        return rows.astype(np.int64), scores, n, approx

    def search_cpu_rows(self, query, max_return_count: int) -> Tuple[np.ndarray, np.ndarray]:
        q = _as_i32(query)
        k = int(max_return_count)
        rows = np.empty(max(k, 1), dtype=np.uint32)
        scores = np.empty(max(k, 1), dtype=np.float32)
        n = C.c_uint32(0)
        check(lib().gsb_db_search_cpu(self._h, q.ctypes.data, q.shape[0], k, rows.ctypes.data,
                                      scores.ctypes.data, C.byref(n)))
        return rows[:n.value].astype(np.int64), scores[:n.value].copy()

    def search_cpu(self, query, dbkey: str, max_return_count: int, similarity_cutoff: float,
                   results_smiles: list, results_ids: list, results_scores: list) -> None:
        """reference FingerprintDB::search_cpu (fingerprintdb_cuda.cpp:20-54): the cutoff is ignored
        and no approximate count is produced (.cpp:38-39)."""
        if dbkey != self.m_dbkey:
            return None
        rows, scores = self.search_cpu_rows(query, max_return_count)
        for r, s in zip(rows, scores):
            results_smiles.append(self.m_smiles[r] if self.m_smiles else int(r))
            results_ids.append(self.m_ids[r] if self.m_ids else int(r))
            results_scores.append(float(s))
        return None

    # -- device-resident entry points (one process per GPU) -------------------------------
    def search_device(self, stream: int, d_query: int, k: int, cutoff: float, d_out_keys: int,
                      d_out_n: int, d_out_survivors: int) -> None:
        check(lib().gsb_db_search_device(self._h, stream, d_query, k, cutoff, d_out_keys, d_out_n,
                                         d_out_survivors))


def merge_device(device: int, stream: int, d_keys: int, d_counts: Optional[int], n_lists: int,
                 list_stride: int, k: int, d_out_rows: int, d_out_scores: int, d_out_n: int) -> None:
    check(lib().gsb_merge_device(device, stream, d_keys, d_counts, n_lists, list_stride, k, d_out_rows,
                                 d_out_scores, d_out_n))


def launch_count() -> int:
    return lib().gsb_launch_count()


__all__ = ["FingerprintDB", "GsbError", "get_gpu_count", "get_next_gpu", "get_available_gpu_memory",
           "top_results_bubble_sort", "fold_fingerprint", "merge_device", "launch_count", "ScanInfo"]
