"""Row-sharded search, one process per GPU (torch.distributed; NCCL over NVLink on the GPU box).

Rows are independent and top-k is a monoid, so the database shards with no data-path
collective: rank r owns the contiguous rows ``shard_range(N, r, world)``, scans them with the
fused kernel (global row ids), and the only exchange is one all-gather of each shard's k best
candidates (k x 8 B + two counters) followed by ``gsb_merge_device`` on every rank.  The
reference merges per-chunk results on the host with std::sort (fingerprintdb_cuda.cu:366)."""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np

RECORD_EXTRA = 2  # [k keys][survivors][n]


def shard_range(total_rows: int, rank: int, world: int) -> Tuple[int, int]:
    """(row_base, n_rows) of rank's contiguous, equal shard (the last one may be short)."""
    per = (total_rows + world - 1) // world
    base = min(total_rows, rank * per)
    return base, min(per, total_rows - base)


def pack_key(score_bits: np.ndarray, rows: np.ndarray) -> np.ndarray:
    """Candidate key of include/gpusim_b200.h: (f32 bits << 32) | (0xFFFFFFFF - row)."""
    return (score_bits.astype(np.uint64) << np.uint64(32)) | (np.uint64(0xFFFFFFFF) - rows.astype(np.uint64))


def unpack_key(keys: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    keys = keys.astype(np.uint64)
    rows = np.uint64(0xFFFFFFFF) - (keys & np.uint64(0xFFFFFFFF))
    scores = (keys >> np.uint64(32)).astype(np.uint32).view(np.float32)
    return rows.astype(np.int64), scores


def exchange_candidates(record, world: int, dist):
    """All-gather every rank's candidate record ([k keys][survivors][n], int64) into one tensor
    laid out rank-major.  Works for NCCL (device tensors) and gloo (host tensors) alike."""
    import torch
    gathered = torch.empty(world * record.numel(), dtype=record.dtype, device=record.device)
    dist.all_gather_into_tensor(gathered, record)
    return gathered


class ShardedSearcher:
    """This rank's shard of a database plus the buffers of the per-query exchange."""

    def __init__(self, db, k: int, local_device: int, dist=None, world: int = 1, rank: int = 0,
                 fused: bool = False):
        import torch
        self.torch, self.db, self.k, self.dist, self.world = torch, db, k, dist, world
        self.rank, self.fused, self.seq = rank, False, 0
        self.device_index = local_device
        dev = torch.device("cuda", local_device)
        self.rec = torch.zeros(k + RECORD_EXTRA, dtype=torch.int64, device=dev)
        self.gathered = torch.zeros(world * (k + RECORD_EXTRA), dtype=torch.int64, device=dev)
        self.out_rows = torch.zeros(k, dtype=torch.int32, device=dev)
        self.out_scores = torch.zeros(k, dtype=torch.float32, device=dev)
        self.out_n = torch.zeros(1, dtype=torch.int32, device=dev)
        self.h_rows = torch.zeros(k, dtype=torch.int32).pin_memory()
        self.h_scores = torch.zeros(k, dtype=torch.float32).pin_memory()
        self.h_n = torch.zeros(1, dtype=torch.int32).pin_memory()
        self.out_approx = torch.zeros(1, dtype=torch.int64, device=dev)
        if fused and world > 1:
            self._setup_fused(dev)

    def _setup_fused(self, dev) -> None:
        """Exchange buffers in NVLink peer-mapped (symmetric) memory for the one-launch query path
        (gsb_db_search_device_fused).  Falls back to the NCCL all-gather path if the platform cannot
        map peer memory; both are product paths."""
        import ctypes as C
        import torch
        from ._lib import Exchange, check, lib
        try:
            import torch.distributed._symmetric_memory as symm_mem
            nbytes = C.c_uint64(0)
            check(lib().gsb_exchange_bytes(self.world, self.k, C.byref(nbytes)))
            self.xbuf = symm_mem.empty(int(nbytes.value), dtype=torch.uint8, device=dev)
            self.xbuf.zero_()
            hdl = symm_mem.rendezvous(self.xbuf, self.dist.group.WORLD)
            self.xchg = Exchange()
            for r in range(self.world):
                self.xchg.peer_base[r] = int(hdl.buffer_ptrs[r])
            self.xchg.rank, self.xchg.world = self.rank, self.world
            torch.cuda.synchronize()
            self.dist.barrier()          # every buffer is zeroed before anybody stores into it
            self.fused = True
        except Exception as e:  # pragma: no cover - depends on the platform
            import sys
            print(f"[gpusimilarity_b200] fused exchange unavailable ({e!r}); using NCCL all-gather", file=sys.stderr)
            self.fused = False

    def search_local(self, d_query_ptr: int, cutoff: float, stream) -> None:
        """One fused scan+select launch over this rank's shard; the record stays in HBM."""
        p = self.rec.data_ptr()
        self.db.search_device(stream.cuda_stream, d_query_ptr, self.k, cutoff, p, p + 8 * (self.k + 1),
                              p + 8 * self.k)

    def search_device(self, d_query_ptr: int, cutoff: float, stream) -> None:
        """Scan, exchange, merge — all asynchronous on ``stream``; results stay in HBM
        (out_rows / out_scores / out_n)."""
        from .fingerprintdb import merge_device
        if self.fused:
            import ctypes as C
            from ._lib import check, lib
            self.seq += 1
            self.xchg.seq = self.seq
            check(lib().gsb_db_search_device_fused(self.db._h, stream.cuda_stream, d_query_ptr, self.k, cutoff,
                                                   C.byref(self.xchg), self.out_rows.data_ptr(),
                                                   self.out_scores.data_ptr(), self.out_n.data_ptr(),
                                                   self.out_approx.data_ptr()))
            return
        self.search_local(d_query_ptr, cutoff, stream)
        if self.world > 1:
            self.dist.all_gather_into_tensor(self.gathered, self.rec)
            src, n_lists = self.gathered, self.world
        else:
            src, n_lists = self.rec, 1
        merge_device(self.device_index, stream.cuda_stream, src.data_ptr(), None, n_lists,
                     self.k + RECORD_EXTRA, self.k, self.out_rows.data_ptr(), self.out_scores.data_ptr(),
                     self.out_n.data_ptr())

    def approx_count(self) -> int:
        if self.fused:
            return int(self.out_approx.item())
        src = self.gathered if self.world > 1 else self.rec
        return int(src.view(self.world if self.world > 1 else 1, self.k + RECORD_EXTRA)[:, self.k].sum().item())

    def search_host(self, d_query, q_pinned, cutoff: float, stream):
        """End to end with host buffers: pinned query in, rows / scores out."""
        d_query.copy_(q_pinned, non_blocking=True)
        self.search_device(d_query.data_ptr(), cutoff, stream)
        self.h_rows.copy_(self.out_rows, non_blocking=True)
        self.h_scores.copy_(self.out_scores, non_blocking=True)
        self.h_n.copy_(self.out_n, non_blocking=True)
        stream.synchronize()
        n = int(self.h_n[0])
        return self.h_rows[:n].numpy().astype(np.int64) & 0xFFFFFFFF, self.h_scores[:n].numpy().copy()


class ShardedBatchSearcher:
    """Multi-query variant (BASELINE config "1024 queries, top-100, 8 GPUs"): every rank scores its
    shard against up to 1024 queries per pass (gsb_db_search_batch_device: bit-sliced kernel; 256
    where only the POPC kernel applies, see max_queries), the per-rank records
    ([nq][k] keys, [nq] survivors, [nq] counts) are all-gathered and merged with one CTA per query."""

    MAX_QUERIES = 1024

    def __init__(self, db, k: int, local_device: int, dist=None, world: int = 1):
        import torch
        self.torch, self.db, self.k, self.dist, self.world, self.device_index = torch, db, k, dist, world, local_device
        dev = torch.device("cuda", local_device)
        nq = self.MAX_QUERIES
        self.rec = torch.zeros(nq * (k + 2), dtype=torch.int64, device=dev)
        self.gathered = torch.zeros(world * nq * (k + 2), dtype=torch.int64, device=dev)
        self.out_rows = torch.zeros(nq * k, dtype=torch.int32, device=dev)
        self.out_scores = torch.zeros(nq * k, dtype=torch.float32, device=dev)
        self.out_n = torch.zeros(nq, dtype=torch.int32, device=dev)
        self.out_approx = torch.zeros(nq, dtype=torch.int64, device=dev)

    def max_queries(self, n_queries: int, cutoff: float) -> int:
        """Queries one search_device call accepts for this batch size and cutoff."""
        import ctypes as C
        from ._lib import check, lib
        out = C.c_uint32(0)
        check(lib().gsb_db_batch_max_queries(self.db._h, self.k, n_queries, cutoff, C.byref(out)))
        return int(out.value)

    def search_device(self, d_queries_ptr: int, n_queries: int, cutoff: float, stream) -> None:
        """Results for queries [0, n_queries) stay in HBM: out_rows/out_scores [nq][k], out_n, out_approx."""
        from ._lib import check, lib
        nq, k = n_queries, self.k
        rec_len = nq * (k + 2)
        p = self.rec.data_ptr()
        check(lib().gsb_db_search_batch_device(self.db._h, stream.cuda_stream, d_queries_ptr, nq, k, cutoff, p,
                                               p + 8 * nq * (k + 1), p + 8 * nq * k))
        src = self.rec
        if self.world > 1:
            self.dist.all_gather_into_tensor(self.gathered[:self.world * rec_len], self.rec[:rec_len])
            src = self.gathered
        check(lib().gsb_merge_batch_device(self.device_index, stream.cuda_stream, src.data_ptr(), self.world, nq, k,
                                           self.out_rows.data_ptr(), self.out_scores.data_ptr(),
                                           self.out_n.data_ptr(), self.out_approx.data_ptr()))
