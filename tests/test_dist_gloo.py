"""world_size-2 gloo tests (CPU) of the host-side logic of the row-sharded search: shard ranges,
candidate-record packing and the all-gather exchange.  Per-shard candidates come from the oracle
here (no GPU); the merge of gathered records is checked against the whole-database oracle."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT
from gpusimilarity_b200.dist import RECORD_EXTRA, exchange_candidates, pack_key, shard_range, unpack_key
from oracle import oracle as O

N, K, SEED, PLANT = 30011, 200, 9, 17


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, cutoff, out):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    base, n = shard_range(N, rank, world)
    rows_np = O.synth_db(SEED, n, 32, PLANT, row_base=base)
    q = O.synth_template(SEED, 32)
    r, s, approx = O.search_gpu(q, rows_np, K, cutoff, row_base=base)
    rec = np.zeros(K + RECORD_EXTRA, dtype=np.uint64)
    rec[:len(r)] = pack_key(s.view(np.uint32), r)
    rec[K], rec[K + 1] = approx, len(r)
    gathered = exchange_candidates(torch.from_numpy(rec.view(np.int64)), world, dist)
    g = gathered.numpy().view(np.uint64).reshape(world, K + RECORD_EXTRA)
    keys = np.concatenate([g[i, :int(g[i, K + 1])] for i in range(world)])
    keys = np.sort(keys)[::-1][:K]                      # canonical order == descending key order
    rows, scores = unpack_key(keys)
    if rank == 0:
        out.put((rows, scores.view(np.uint32).copy(), int(g[:, K].sum())))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("cutoff", [0.0, 0.3])
def test_sharded_exchange_world2(cutoff):
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, cutoff, out)) for r in range(2)]
    for p in procs:
        p.start()
    rows, score_bits, approx = out.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    whole = O.synth_db(SEED, N, 32, PLANT)
    w_rows, w_scores, w_approx = O.search_gpu(O.synth_template(SEED, 32), whole, K, cutoff)
    assert approx == w_approx
    assert np.array_equal(rows, w_rows) and np.array_equal(score_bits, w_scores.view(np.uint32))


def test_shard_ranges_cover_rows_exactly():
    for total in (0, 1, 7, 1000, 10 ** 9, 10 ** 9 + 3):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(total, r, world) for r in range(world)]
            assert sum(n for _, n in spans) == total
            pos = 0
            for base, n in spans:
                assert base == min(pos, total) and n >= 0
                pos += n


def test_key_packing_orders_canonically():
    scores = np.array([0.5, 0.5, 0.25, 1.0, 0.0], dtype=np.float32)
    rows = np.array([7, 3, 1, 9, 2], dtype=np.int64)
    keys = pack_key(scores.view(np.uint32), rows)
    order = np.argsort(keys)[::-1]
    assert list(rows[order]) == [9, 3, 7, 1, 2]          # score desc, then row asc
    r2, s2 = unpack_key(keys)
    assert np.array_equal(r2, rows) and np.array_equal(s2, scores)
