"""Tensor-core multi-query kernel (gsb_tensor.cuh, tcgen05.mma kind::i8) against the oracle and the
other multi-query kernels: rows and f32 score bits must be identical (bit-exact bar of the path:
popcounts are integers, the score is the reference's one IEEE divide, fingerprintdb_cuda.cu:89-103)."""
import ctypes as C

import numpy as np
import pytest

import gpusimilarity_b200 as gsb
from oracle import oracle as O
from oracle import oracle_c as OC
from test_gpu_parity import _sliced_queries, assert_same, make_db

pytestmark = pytest.mark.gpu


def _dense_queries(rng, n, bits):
    qs = np.zeros((n, 32), np.uint32)
    for j in range(n):
        pos = rng.choice(1024, bits, replace=False)
        np.bitwise_or.at(qs[j], pos // 32, np.uint32(1) << (pos % 32).astype(np.uint32))
    return qs.view(np.int32)


@pytest.mark.parametrize("n_rows", [1, 33, 127, 128, 129, 4096, 37889, 300001])
def test_tensor_kernel_sizes_and_ragged_tiles(monkeypatch, n_rows):
    """Forced on (GSB_BATCH_KERNEL=4): ragged last batch and last 128-row tile, fewer tiles than
    SMs and several tiles per CTA, k above and below the row count, cutoffs that follow the
    threshold (<= 0) and the cutoff (> 0); sparse, empty, all-ones and half-dense queries."""
    monkeypatch.setenv("GSB_BATCH_KERNEL", "4")
    rows_np = OC.c_synth_db(177 + n_rows, n_rows, 32, 53)
    db = make_db(rows_np)
    qs = _sliced_queries(rows_np, 177 + n_rows, 14)
    for k in (1, 10, 100):
        for cutoff in (0.0, -1.0, 0.08, 0.3, 0.6, 1.5):
            res = db.search_batch_rows(qs, k, cutoff)
            for j, got in enumerate(res):
                assert_same(got, OC.c_search(qs[j], rows_np, k, cutoff), f"tensor n={n_rows} q{j} k={k} c={cutoff}")


@pytest.mark.parametrize("n_queries,k,grid", [(37, 100, 0), (128, 100, 0), (300, 100, 0), (1030, 10, 0), (40, 512, 0),
                                              (64, 100, 8), (129, 100, 3)])
def test_tensor_kernel_ties_groups_and_small_grids(monkeypatch, n_queries, k, grid):
    """Tie groups, empty rows, more than 128 queries (one launch per 128) and more than 1024 (two
    calls), k = 512, and small grids: many tiles per CTA, long candidate lists (select rounds),
    thresholds shared between few CTAs."""
    monkeypatch.setenv("GSB_BATCH_KERNEL", "4")
    if grid:
        monkeypatch.setenv("GSB_GRID", str(grid))
    rows_np = OC.c_synth_db(1900 + n_queries, 700_000, 32, 97)
    rows_np[1000:1100] = rows_np[5]                                 # a tie group
    rows_np[2000:2050] = 0                                          # empty fingerprints
    rows_np[650_000:650_600] = rows_np[7]                           # a tie group larger than k
    db = make_db(rows_np)
    qs = _sliced_queries(rows_np, n_queries, n_queries - 6)
    qs[6], qs[7] = rows_np[5], rows_np[7]
    step = max(1, n_queries // 40)                                  # oracle-check a spread of the queries
    for cutoff in (0.0, 0.4):
        res = db.search_batch_rows(qs, k, cutoff)
        assert len(res) == n_queries
        for j in list(range(0, n_queries, step)) + [n_queries - 1]:
            assert_same(res[j], OC.c_search(qs[j], rows_np, k, cutoff), f"tensor query {j} cutoff {cutoff}")


def test_tensor_kernel_dense_queries_full_grid(monkeypatch):
    """3 M rows (one CTA per SM, ~160 tiles each), queries of 32 / 128 / 512 / 1000 set bits: the
    tensor-core kernel against the bit-sliced kernel on every query and the oracle on a few."""
    rows_np = OC.c_synth_db(88, 3_000_000, 32, 1500)
    db = make_db(rows_np)
    rng = np.random.default_rng(88)
    qs = np.concatenate([_dense_queries(rng, 40, b) for b in (32, 128, 512, 1000)] + [_sliced_queries(rows_np, 88, 26)])
    for cutoff in (0.0, 0.25):
        monkeypatch.setenv("GSB_BATCH_KERNEL", "3")
        want = db.search_batch_rows(qs, 100, cutoff)
        monkeypatch.setenv("GSB_BATCH_KERNEL", "4")
        got = db.search_batch_rows(qs, 100, cutoff)
        for j in range(len(qs)):
            assert_same(got[j], want[j], f"tensor vs bit-sliced, query {j}, cutoff {cutoff}")
        for j in (0, 45, 90, 130, 161):
            assert_same(got[j], OC.c_search(qs[j], rows_np, 100, cutoff), f"oracle, query {j}, cutoff {cutoff}")


def test_tensor_kernel_full_size(monkeypatch):
    """200 M synthetic rows in HBM (10 500 tiles per CTA: thresholds converge, lists are pruned, the
    pipeline runs for tens of milliseconds): dense and sparse queries through the tensor-core kernel
    against the streamed oracle (every row scored on the host cores), the single-query kernel and the
    bit-sliced kernel, with and without a cutoff."""
    n, seed, plant, k = 200_000_000, 77, 50000, 100
    whole = gsb.FingerprintDB.synthetic(n, device=0, seed=seed, plant_period=plant)
    rng = np.random.default_rng(n)
    qs = np.concatenate([O.synth_template(seed, 32)[None, :],
                         np.stack([whole.getFingerprint(int(r)) for r in rng.integers(0, n, 11)]),
                         _dense_queries(rng, 6, 200), _dense_queries(rng, 6, 512)])
    for cutoff in (0.0, 0.1):
        monkeypatch.setenv("GSB_BATCH_KERNEL", "4")
        got = whole.search_batch_rows(qs, k, cutoff)
        monkeypatch.setenv("GSB_BATCH_KERNEL", "3")
        want = whole.search_batch_rows(qs, k, cutoff)
        for j in range(len(qs)):
            assert_same(got[j], want[j], f"full size, tensor vs bit-sliced, query {j}, cutoff {cutoff}")
        for j in (0, 5, 14, 20):
            assert_same(got[j], whole.search_rows(qs[j], k, cutoff), f"full size, tensor vs single-query kernel, query {j}")
        for j in ((0, 13) if cutoff == 0.0 else (1,)):
            assert_same(got[j], OC.c_stream_search(qs[j], seed, plant, n, k, cutoff), f"full size, streamed oracle, query {j}")
    whole.close()


def test_tensor_kernel_choice(monkeypatch):
    """Automatic dispatch: the set bits of the batch decide between the bit-sliced kernel (cost per set
    bit) and the tensor-core kernel (cost per 128 queries); other row widths and metrics never get it."""
    from gpusimilarity_b200._lib import check, lib, GSB_BATCH_SLICED

    def mode(db, qs, k=100, cutoff=0.0):
        m, per = C.c_int(0), C.c_uint32(0)
        check(lib().gsb_db_batch_mode(db._h, k, len(qs), cutoff, C.byref(m), C.byref(per)))
        return m.value
    monkeypatch.delenv("GSB_BATCH_KERNEL", raising=False)
    rows_np = OC.c_synth_db(5, 50_000, 32, 53)
    db = make_db(rows_np)
    rng = np.random.default_rng(5)
    dense = _dense_queries(rng, 128, 512)
    sparse = _dense_queries(rng, 128, 24)
    assert mode(db, dense) == GSB_BATCH_SLICED                      # (the mode query does not see the queries)
    # results are the same whichever kernel the library picks
    for qs in (dense, sparse):
        auto = db.search_batch_rows(qs, 50, 0.0)
        monkeypatch.setenv("GSB_BATCH_KERNEL", "3")
        want = db.search_batch_rows(qs, 50, 0.0)
        monkeypatch.delenv("GSB_BATCH_KERNEL")
        for j in range(0, 128, 9):
            assert_same(auto[j], want[j], f"automatic choice, query {j}")
    monkeypatch.setenv("GSB_BATCH_KERNEL", "4")
    narrow = make_db(OC.c_synth_db(6, 20_000, 8, 31))               # 256-bit rows: no tensor-core kernel
    q8 = np.stack([narrow.getFingerprint(i) for i in range(10)])
    res = narrow.search_batch_rows(q8, 10, 0.0)
    for j in range(10):
        assert_same(res[j], OC.c_search(q8[j], OC.c_synth_db(6, 20_000, 8, 31), 10, 0.0), f"narrow rows, query {j}")


def test_tensor_kernel_pipeline_timeout_is_an_error_not_a_hang(monkeypatch):
    """A launch whose pipeline cannot make progress (test hook: CTA 0 never loads its first tile)
    reports GSB_ERR_CUDA through the error word instead of hanging or trapping, and the very next
    search on the same context is right again."""
    monkeypatch.setenv("GSB_BATCH_KERNEL", "4")
    rows_np = OC.c_synth_db(9, 200_000, 32, 53)
    db = make_db(rows_np)
    qs = _sliced_queries(rows_np, 9, 10)
    good = db.search_batch_rows(qs, 10, 0.0)
    monkeypatch.setenv("GSB_SPIN_TIMEOUT_MS", "20")
    monkeypatch.setenv("GSB_TC_FAULT", "1")
    with pytest.raises(gsb.GsbError):
        db.search_batch_rows(qs, 10, 0.0)
    monkeypatch.delenv("GSB_TC_FAULT")
    monkeypatch.delenv("GSB_SPIN_TIMEOUT_MS")
    again = db.search_batch_rows(qs, 10, 0.0)
    for j in range(len(qs)):
        assert_same(again[j], good[j], f"after the timeout, query {j}")
