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

    def __init__(self, db, k: int, local_device: int, dist=None, world: int = 1):
        import torch
        self.torch, self.db, self.k, self.dist, self.world = torch, db, k, dist, world
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

    def search_local(self, d_query_ptr: int, cutoff: float, stream) -> None:
        """One fused scan+select launch over this rank's shard; the record stays in HBM."""
        p = self.rec.data_ptr()
        self.db.search_device(stream.cuda_stream, d_query_ptr, self.k, cutoff, p, p + 8 * (self.k + 1),
                              p + 8 * self.k)

    def search_device(self, d_query_ptr: int, cutoff: float, stream) -> None:
        """Scan, exchange, merge — all asynchronous on ``stream``; results stay in HBM
        (out_rows / out_scores / out_n)."""
        from .fingerprintdb import merge_device
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
