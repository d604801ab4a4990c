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
    """This rank's shard of a database plus the buffers of the per-query exchange.

    Three ways to run a query, all through gsb_db_search_enqueue (include/gpusim_b200.h):

    * ``search_local``  — this shard only, record stays in HBM (the kernel the roofline is quoted on);
    * ``search_device`` — scan + cross-rank exchange + merge, results stay in HBM.  Fused (default):
      ONE launch per rank, the exchange runs inside the kernel over NVLink peer memory.  Otherwise
      NCCL all-gather + merge kernel;
    * ``submit_host`` / ``wait_host`` — the same with HOST buffers: the query travels as a kernel
      parameter, the last CTA stores rows / scores straight into pinned host memory and raises a
      completion word the host polls (no cudaMemcpy, no stream synchronize).  ``DEPTH`` queries may
      be in flight; consecutive launches overlap on the device (programmatic dependent launch).
    """

    DEPTH = 2

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
        self.out_approx = torch.zeros(1, dtype=torch.int64, device=dev)
        # host sinks, one per query in flight: [rows k i32 | scores k f32 | keys k i64] [n, approx, done]
        self.h_rows = [torch.zeros(k, dtype=torch.int32).pin_memory() for _ in range(self.DEPTH)]
        self.h_scores = [torch.zeros(k, dtype=torch.float32).pin_memory() for _ in range(self.DEPTH)]
        self.h_keys = [torch.zeros(k, dtype=torch.int64).pin_memory() for _ in range(self.DEPTH)]
        self.h_meta = [torch.zeros(4, dtype=torch.int64).pin_memory() for _ in range(self.DEPTH)]  # n, approx, done
        self.h_events = [None] * self.DEPTH
        self.host_seq = 0
        if fused and world > 1:
            self._setup_fused(dev)

    def _setup_fused(self, dev) -> None:
        """Exchange buffers in NVLink peer-mapped (symmetric) memory for the one-launch query path.
        Falls back to the NCCL all-gather path if the platform cannot map peer memory; both are
        product paths."""
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

    # -- launches -------------------------------------------------------------------------
    def _enqueue(self, stream, h_query_ptr, d_query_ptr, stable, cutoff, xchg, sink) -> None:
        import ctypes as C
        from ._lib import GSB_QUERY_STABLE, check, lib
        check(lib().gsb_db_search_enqueue(self.db._h, stream.cuda_stream, h_query_ptr, d_query_ptr,
                                          GSB_QUERY_STABLE if stable else 0, self.k, cutoff,
                                          C.byref(xchg) if xchg is not None else None, C.byref(sink)))

    def search_local(self, d_query_ptr: int, cutoff: float, stream, stable: bool = True) -> None:
        """One fused scan+select launch over this rank's shard; the record stays in HBM.  ``stable``:
        nothing queued on the stream since the previous search writes the query buffer."""
        from ._lib import Sink
        p = self.rec.data_ptr()
        sink = Sink(keys=p, n=p + 8 * (self.k + 1), approx=p + 8 * self.k)
        self._enqueue(stream, None, d_query_ptr, stable, cutoff, None, sink)

    def search_device(self, d_query_ptr: int, cutoff: float, stream, stable: bool = True) -> None:
        """Scan, exchange, merge — all asynchronous on ``stream``; results stay in HBM
        (out_rows / out_scores / out_n / out_approx)."""
        from ._lib import Sink
        from .fingerprintdb import merge_device
        if self.fused:
            self.seq += 1
            self.xchg.seq = self.seq
            sink = Sink(rows=self.out_rows.data_ptr(), scores=self.out_scores.data_ptr(), n=self.out_n.data_ptr(),
                        approx=self.out_approx.data_ptr())
            self._enqueue(stream, None, d_query_ptr, stable, cutoff, self.xchg, sink)
            return
        with self.torch.cuda.stream(stream):     # the collective follows torch's current stream
            self.search_local(d_query_ptr, cutoff, stream, stable)
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

    # -- host buffers in and out ------------------------------------------------------------
    def submit_host(self, query_np: np.ndarray, cutoff: float, stream) -> int:
        """Queue one query given in HOST memory; returns a handle for ``wait_host``."""
        from ._lib import Sink
        q = np.ascontiguousarray(query_np, dtype=np.int32)
        self.host_seq += 1
        slot = self.host_seq % self.DEPTH
        meta = self.h_meta[slot].data_ptr()
        if self.fused:
            self.seq += 1
            self.xchg.seq = self.seq
            sink = Sink(rows=self.h_rows[slot].data_ptr(), scores=self.h_scores[slot].data_ptr(), n=meta,
                        approx=meta + 8, done=meta + 16, done_value=self.host_seq)
            self._enqueue(stream, q.ctypes.data, None, True, cutoff, self.xchg, sink)
        elif self.world == 1:
            sink = Sink(keys=self.h_keys[slot].data_ptr(), n=meta, approx=meta + 8, done=meta + 16,
                        done_value=self.host_seq)
            self._enqueue(stream, q.ctypes.data, None, True, cutoff, None, sink)
        else:                                    # NCCL path: H2D, collective, merge kernel, D2H
            with self.torch.cuda.stream(stream):
                if not hasattr(self, "_d_q"):
                    self._d_q = self.torch.zeros(q.shape[0], dtype=self.torch.int32, device=self.rec.device)
                    self._h_q = self.torch.zeros(q.shape[0], dtype=self.torch.int32).pin_memory()
                self._h_q.numpy()[:] = q
                self._d_q.copy_(self._h_q, non_blocking=True)
                self.search_device(self._d_q.data_ptr(), cutoff, stream, stable=False)
                self.h_rows[slot].copy_(self.out_rows, non_blocking=True)
                self.h_scores[slot].copy_(self.out_scores, non_blocking=True)
                self.h_meta[slot][:1].copy_(self.out_n.to(self.torch.int64), non_blocking=True)
                ev = self.torch.cuda.Event()
                ev.record(stream)
                self.h_events[slot] = ev
        return self.host_seq

    def wait_host(self, handle: int):
        """(global rows, scores, approximate count or None) of a ``submit_host`` query."""
        from ._lib import GSB_COUNT_ERROR, GsbError, check, lib
        slot = handle % self.DEPTH
        meta = self.h_meta[slot]
        if self.fused or self.world == 1:
            check(lib().gsb_wait_word(meta.data_ptr() + 16, handle, 120_000_000))
            n = int(meta[0]) & 0xFFFFFFFF
            if n == GSB_COUNT_ERROR:
                raise GsbError(2, "the search kernel reported a failure (barrier or peer flag timed out)")
            approx = int(meta[1])
            if self.fused:
                return (self.h_rows[slot][:n].numpy().astype(np.int64) & 0xFFFFFFFF,
                        self.h_scores[slot][:n].numpy().copy(), approx)
            rows, scores = unpack_key(self.h_keys[slot][:n].numpy().view(np.uint64))
            return rows, scores.copy(), approx
        self.h_events[slot].synchronize()
        n = int(meta[0])
        return self.h_rows[slot][:n].numpy().astype(np.int64) & 0xFFFFFFFF, self.h_scores[slot][:n].numpy().copy(), None

    def search_host(self, query_np: np.ndarray, cutoff: float, stream):
        """End to end with host buffers, one query at a time."""
        return self.wait_host(self.submit_host(query_np, cutoff, stream))


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
        with self.torch.cuda.stream(stream):     # the collective follows torch's current stream
            check(lib().gsb_db_search_batch_device(self.db._h, stream.cuda_stream, d_queries_ptr, nq, k, cutoff, p,
                                                   p + 8 * nq * (k + 1), p + 8 * nq * k))
            src = self.rec
            if self.world > 1:
                self.dist.all_gather_into_tensor(self.gathered[:self.world * rec_len], self.rec[:rec_len])
                src = self.gathered
            check(lib().gsb_merge_batch_device(self.device_index, stream.cuda_stream, src.data_ptr(), self.world, nq, k,
                                               self.out_rows.data_ptr(), self.out_scores.data_ptr(),
                                               self.out_n.data_ptr(), self.out_approx.data_ptr()))

    def search_host(self, queries_np: np.ndarray, cutoff: float, stream):
        """End to end with HOST buffers: pinned queries in (H2D), one pass over every shard, exchange,
        merge, results back into pinned host arrays (D2H).  Returns (rows [nq][k], scores, n, approx)
        as numpy views of the pinned buffers; ``bytes_h2d`` / ``bytes_d2h`` hold the traffic."""
        torch = self.torch
        q = np.ascontiguousarray(queries_np, dtype=np.int32)
        nq, k = q.shape[0], self.k
        if not hasattr(self, "_h_q") or self._h_q.shape[0] < nq or self._h_q.shape[1] != q.shape[1]:
            self._h_q = torch.zeros((max(nq, self.MAX_QUERIES), q.shape[1]), dtype=torch.int32).pin_memory()
            self._d_q = torch.zeros_like(self._h_q, device=self.rec.device)
            self._h_rows = torch.zeros(self.MAX_QUERIES * k, dtype=torch.int32).pin_memory()
            self._h_scores = torch.zeros(self.MAX_QUERIES * k, dtype=torch.float32).pin_memory()
            self._h_n = torch.zeros(self.MAX_QUERIES, dtype=torch.int32).pin_memory()
            self._h_approx = torch.zeros(self.MAX_QUERIES, dtype=torch.int64).pin_memory()
        self._h_q.numpy()[:nq] = q
        with torch.cuda.stream(stream):
            self._d_q[:nq].copy_(self._h_q[:nq], non_blocking=True)
            self.search_device(self._d_q.data_ptr(), nq, cutoff, stream)
            self._h_rows[:nq * k].copy_(self.out_rows[:nq * k], non_blocking=True)
            self._h_scores[:nq * k].copy_(self.out_scores[:nq * k], non_blocking=True)
            self._h_n[:nq].copy_(self.out_n[:nq], non_blocking=True)
            self._h_approx[:nq].copy_(self.out_approx[:nq], non_blocking=True)
        stream.synchronize()
        self.bytes_h2d, self.bytes_d2h = q.nbytes, nq * k * 8 + nq * 12
        return (self._h_rows[:nq * k].numpy().reshape(nq, k).astype(np.int64) & 0xFFFFFFFF,
                self._h_scores[:nq * k].numpy().reshape(nq, k), self._h_n[:nq].numpy(), self._h_approx[:nq].numpy())
