"""GPU parity tests: the CUDA path, called through the C ABI (ctypes mirror of the reference's
FingerprintDB), against the CPU oracle on identical inputs.  Bar: bit-exact rows and f32 score
bit patterns in canonical order (score desc, row asc), exact approximate counts."""
import os

import numpy as np
import pytest

from conftest import f32bits
from gpusimilarity_b200.dist import unpack_key
from oracle import oracle as O
from oracle import oracle_c as OC

pytestmark = pytest.mark.gpu

import gpusimilarity_b200 as gsb  # noqa: E402


def make_db(rows: np.ndarray, fold: int = 1, dbkey: str = "pass", devices=None) -> gsb.FingerprintDB:
    rows = np.ascontiguousarray(rows, dtype=np.int32)
    db = gsb.FingerprintDB(rows.shape[1] * 32, rows.shape[0], dbkey, [rows])
    db.copyToGPU(fold, devices)
    return db


def assert_same(got, want, what=""):
    g_rows, g_scores, g_approx = got
    w_rows, w_scores, w_approx = want
    assert g_approx == w_approx, f"{what}: approx {g_approx} != {w_approx}"
    assert len(g_rows) == len(w_rows), f"{what}: {len(g_rows)} results, want {len(w_rows)}"
    assert np.array_equal(f32bits(g_scores), f32bits(w_scores)), f"{what}: score bits differ"
    assert np.array_equal(np.asarray(g_rows), np.asarray(w_rows)), f"{what}: rows differ"


def check(db, rows_np, query, k, cutoff, what=""):
    got = db.search_rows(query, k, cutoff)
    want = OC.c_search(query, rows_np, k, cutoff)
    assert_same(got, want, f"{what} k={k} cutoff={cutoff}")


def assert_appendix_d(got, ref, query, rows_np, cutoff, what=""):
    """SURVEY App. D parity rules against the reference's own CUDA path on a multi-chunk database,
    where the reference's survivors of a tie group at the k boundary depend on heap addresses:
    (i) score vectors equal as f32 bit patterns, (ii) the same rows above the k-th score,
    (iii) every returned row at the k-th score really has that score; plus the approximate count."""
    g_rows, g_scores, g_approx = got
    r_rows, r_scores, r_approx = ref
    assert g_approx == r_approx, f"{what}: approx {g_approx} != {r_approx}"
    assert np.array_equal(f32bits(g_scores), f32bits(r_scores)), f"{what}: (i) score vectors differ"
    if len(g_scores) == 0:
        return
    s_k = g_scores[-1]
    assert sorted(g_rows[g_scores > s_k].tolist()) == sorted(r_rows[r_scores > s_k].tolist()), \
        f"{what}: (ii) rows above the k-th score differ"
    for rows, scores in ((g_rows, g_scores), (r_rows, r_scores)):
        at = rows[scores == s_k]
        assert len(set(at.tolist())) == len(at)
        true = O.tanimoto_scores_gpu(query, rows_np[at], cutoff)
        assert np.all(f32bits(true) == f32bits(np.float32(s_k))), f"{what}: (iii) boundary rows mis-scored"


# ------------------------------------------------------------------ the reference's own tests
def test_reference_similarity_cutoff(golden, small_fsim, small_db):
    """reference test/test_gpusim.cpp:101-128."""
    db = gsb.FingerprintDB(1024, 100, "pass", small_fsim.fp_chunks, list(small_fsim.smiles),
                           list(small_fsim.ids))
    db.copyToGPU(1)
    t = golden["reference_tests"]["TestSimilarityCutoff"]
    fp = db.getFingerprint(t["query_row"])
    for cutoff, n_res, n_approx in zip(t["cutoffs"], t["result_counts"], t["approximate_counts"]):
        smiles, ids, scores = [], [], []
        approx = db.search(fp, "pass", t["k"], cutoff, smiles, ids, scores)
        assert len(smiles) == n_res and approx == n_approx


def test_reference_compare_gpu_to_cpu(golden, small_fsim):
    """reference test/test_gpusim.cpp:29-69: GPU result order == search_cpu order."""
    db = gsb.FingerprintDB(1024, 100, "pass", small_fsim.fp_chunks, list(small_fsim.smiles),
                           list(small_fsim.ids))
    db.copyToGPU(1)
    t = golden["reference_tests"]["CompareGPUtoCPU"]
    fp = db.getFingerprint(t["query_row"])
    for k in t["return_counts"]:
        g_smiles, g_ids, g_scores = [], [], []
        db.search(fp, "pass", k, 0.0, g_smiles, g_ids, g_scores)
        c_smiles, c_ids, c_scores = [], [], []
        db.search_cpu(fp, "pass", k, 0.0, c_smiles, c_ids, c_scores)
        assert len(g_smiles) == k and g_smiles == c_smiles
    # key mismatch: silently empty (fingerprintdb_cuda.cu:349-352)
    s, i, f = [], [], []
    assert db.search(fp, "wrong", 10, 0.0, s, i, f) is None and s == []


def test_golden_scores_small_fsim(golden, small_fsim, small_db):
    """Frozen outputs of the reference's own build: rows and f32 bit patterns."""
    db = gsb.FingerprintDB(1024, 100, "pass", small_fsim.fp_chunks)
    db.copyToGPU(1)
    for g in golden["search_cpu"]:
        rows, scores, approx = db.search_rows(small_db[g["query_row"]], g["k"], 0.0)
        assert list(rows) == g["rows"] and list(f32bits(scores)) == g["score_bits"] and approx == 100


def test_golden_synthetic(golden):
    s = golden["synthetic"]
    rows_np = O.synth_db(s["seed"], s["rows"], 32, s["plant_period"])
    db = make_db(rows_np)
    queries = {"template": O.synth_template(s["seed"], 32), "row123": rows_np[123],
               "zero": np.zeros(32, np.int32)}
    for g in s["queries"]:
        rows, scores, _ = db.search_rows(queries[g["name"]], g["k"], 0.0)
        assert list(rows) == g["rows"] and list(f32bits(scores)) == g["score_bits"]


def test_division_is_bit_exact_for_every_popcount_pair():
    """The kernel's FMA division vs the IEEE divide the reference uses, for all 33.5 M
    (common, union) pairs with union <= 8192."""
    import ctypes
    from gpusimilarity_b200 import _lib
    bad = ctypes.c_uint64(1)
    _lib.check(_lib.lib().gsb_selftest_division(0, ctypes.byref(bad)))
    assert bad.value == 0


# ------------------------------------------------------------------ oracle parity sweeps
@pytest.mark.parametrize("n_rows", [1, 31, 32, 33, 255, 256, 257, 1000, 37889, 300001])
def test_sizes_and_ragged_tails(n_rows):
    rows_np = O.synth_db(1234 + n_rows, n_rows, 32, 41)
    db = make_db(rows_np)
    tmpl = O.synth_template(1234 + n_rows, 32)
    for q in (tmpl, rows_np[n_rows // 2]):
        for k in (1, 10, 1000):
            for cutoff in (0.0, 0.08, 0.5):
                check(db, rows_np, q, k, cutoff, f"n={n_rows}")
    check(db, rows_np, tmpl, n_rows + 5, 0.0, "k>N")        # k > N returns every row
    check(db, rows_np, tmpl, 3, -1.0, "negative cutoff")
    check(db, rows_np, tmpl, 3, 1.5, "cutoff>1")             # nothing survives


def test_ties_zero_rows_and_nan():
    rng = np.random.default_rng(5)
    rows_np = O.synth_db(99, 50000, 32, 0)
    rows_np[rng.integers(0, 50000, 4000)] = rows_np[17]      # a 4000-row tie group at score 1.0
    rows_np[rng.integers(0, 50000, 3000)] = 0                # all-zero rows
    db = make_db(rows_np)
    zero = np.zeros(32, np.int32)
    for k in (5, 1000, 4500):
        check(db, rows_np, rows_np[17], k, 0.0, "ties")
        check(db, rows_np, rows_np[17], k, 0.9, "ties+cutoff")
        check(db, rows_np, zero, k, 0.0, "zero query: 0/0 -> 0")
        check(db, rows_np, zero, k, 0.1, "zero query with cutoff")


def test_adversarial_order_forces_compactions():
    """Rows sorted by ascending score: every row beats the running threshold."""
    base = O.synth_db(7, 120000, 32, 3)
    q = O.synth_template(7, 32)
    order = np.argsort(O.tanimoto_scores_cpu(q, base), kind="stable")
    rows_np = np.ascontiguousarray(base[order])
    db = make_db(rows_np)
    for k in (10, 1000, 3000):
        check(db, rows_np, q, k, 0.0, "ascending")
    rows_desc = np.ascontiguousarray(rows_np[::-1])
    db2 = make_db(rows_desc)
    check(db2, rows_desc, q, 1000, 0.0, "descending")


@pytest.mark.parametrize("k", [2000, 5000, 12000, 20000, 70000])
def test_large_k(k):
    rows_np = O.synth_db(31, 200000, 32, 11)
    db = make_db(rows_np)
    check(db, rows_np, O.synth_template(31, 32), k, 0.0, "large k")
    check(db, rows_np, O.synth_template(31, 32), k, 0.2, "large k + cutoff")


@pytest.mark.parametrize("bits", [128, 256, 512, 2048, 4096, 96, 768, 1536])
def test_other_widths(bits):
    words = bits // 32
    rows_np = O.synth_db(bits, 20011, words, 23)
    db = make_db(rows_np)
    assert db.getFingerprintBitcount() == bits
    for q in (O.synth_template(bits, words), rows_np[77]):
        check(db, rows_np, q, 100, 0.0, f"bits={bits}")
        check(db, rows_np, q, 100, 0.3, f"bits={bits}")
    assert np.array_equal(db.getFingerprint(77), rows_np[77])


@pytest.mark.parametrize("metric,alpha,beta", [("dice", 1.0, 1.0), ("tversky", 1.0, 1.0), ("tversky", 0.5, 0.5),
                                               ("tversky", 0.9, 0.1), ("tversky", 0.0, 1.0), ("tversky", 2.0, 0.3)])
def test_dice_and_tversky(metric, alpha, beta):
    """the other metrics share the scan, only the epilogue differs.  Single-query kernel,
    multi-query kernel (the POPC one: the bit-sliced filter is Tanimoto's), folded re-score and
    search_cpu against the oracle's definition; Tversky(1,1) must equal Tanimoto score for score."""
    rows_np = O.synth_db(404, 90_000, 32, 37)
    rows_np[100:140] = 0
    db = make_db(rows_np)
    db.setMetric(metric, alpha, beta)
    qs = [O.synth_template(404, 32), rows_np[7], np.zeros(32, np.int32), np.full(32, -1, np.int32)]
    for q in qs:
        for k, cutoff in ((10, 0.0), (1000, 0.0), (200, 0.25), (50, 0.7)):
            assert_same(db.search_rows(q, k, cutoff), O.search_gpu_metric(q, rows_np, k, cutoff, metric, alpha, beta),
                        f"{metric}({alpha},{beta}) k={k} cutoff={cutoff}")
    if metric == "tversky" and alpha == beta == 1.0:
        assert_same(db.search_rows(qs[0], 500, 0.1), OC.c_search(qs[0], rows_np, 500, 0.1), "tversky(1,1) == tanimoto")
    mode, per = db.batch_mode(100, len(qs))
    for got, q in zip(db.search_batch_rows(np.stack(qs), 100, 0.05), qs):
        assert_same(got, O.search_gpu_metric(q, rows_np, 100, 0.05, metric, alpha, beta), f"{metric} batch (mode {mode})")
    many = np.stack([rows_np[i * 31] for i in range(40)])          # 40 queries would pick the bit-sliced kernel
    assert db.batch_mode(100, 40)[0] == 1                           # ... but not for another metric: POPC kernel
    for got, q in zip(db.search_batch_rows(many, 100, 0.0), many):
        assert_same(got, O.search_gpu_metric(q, rows_np, 100, 0.0, metric, alpha, beta), f"{metric} 40-query batch")
    db.setMetric("tanimoto")
    check(db, rows_np, qs[0], 100, 0.0, "back to tanimoto")


def test_async_tickets_keep_four_queries_in_flight():
    """gsb_db_search_async / _wait: four tickets outstanding, waited out of order; the fifth is refused
    until one is waited for; folded databases and very large k answer through the same two calls."""
    rows_np = O.synth_db(515, 250_000, 32, 29)
    db = make_db(rows_np)
    qs = [O.synth_template(515, 32), rows_np[1], rows_np[249_999], np.zeros(32, np.int32), rows_np[77]]
    want = [OC.c_search(q, rows_np, 1000, 0.1) for q in qs]
    for rep in range(3):
        tickets = [db.search_rows_async(q, 1000, 0.1) for q in qs[:4]]
        with pytest.raises(gsb.GsbError):
            db.search_rows_async(qs[4], 1000, 0.1)
        assert_same(db.search_rows(qs[4], 1000, 0.1), want[4], "sync call between async ones")
        for i in (2, 0, 3, 1):
            assert_same(db.search_rows_wait(tickets[i]), want[i], f"rep {rep} ticket {i}")
    # pipelined: always two in flight
    t_prev = db.search_rows_async(qs[0], 1000, 0.1)
    for i in range(1, 40):
        t = db.search_rows_async(qs[i % 5], 1000, 0.1)
        assert_same(db.search_rows_wait(t_prev), want[(i - 1) % 5], f"pipelined {i}")
        t_prev = t
    assert_same(db.search_rows_wait(t_prev), want[39 % 5], "pipelined tail")
    t = db.search_rows_async(qs[0], 30_000, 0.0)                    # more than one launch can select
    assert_same(db.search_rows_wait(t), OC.c_search(qs[0], rows_np, 30_000, 0.0), "deferred large k")
    folded = make_db(rows_np, fold=4)
    t = folded.search_rows_async(qs[0], 20, 0.3)
    r, s_, a = folded.search_rows_wait(t)
    w = O.search_gpu_folded(qs[0], rows_np, 20, 0.3, 4)
    assert a == w[2] and np.array_equal(r, w[0]) and np.array_equal(f32bits(s_), f32bits(w[1]))


def test_searches_on_different_streams_are_ordered():
    """two device searches of one database on different streams share the shard's
    workspace; the library orders them itself (event on the earlier stream's tail)."""
    import torch
    from gpusimilarity_b200.dist import ShardedSearcher
    n, k, seed = 3_000_000, 1000, 626
    rows_np = OC.c_synth_db(seed, n, 32, 900)
    db = make_db(rows_np)
    dev = torch.device("cuda", 0)
    qs = [O.synth_template(seed, 32), rows_np[5], rows_np[n - 1]]
    d_qs = [torch.from_numpy(np.ascontiguousarray(q)).to(dev) for q in qs]
    streams = [torch.cuda.Stream() for _ in range(3)]
    searchers = [ShardedSearcher(db, k, 0) for _ in range(3)]
    torch.cuda.synchronize()
    for rep in range(5):
        for i in range(3):                         # three launches back to back on three streams
            searchers[i].search_local(d_qs[i].data_ptr(), 0.0, streams[i])
        got_sync = db.search_rows(qs[rep % 3], k, 0.0)   # and the host-buffer call on the library's own stream
        torch.cuda.synchronize()
        for i in range(3):
            rec = searchers[i].rec.cpu().numpy().view(np.uint64)
            cnt = int(rec[k + 1]) & 0xffffffff
            rows, scores = unpack_key(rec[:cnt])
            assert_same((rows, scores, int(rec[k])), OC.c_search(qs[i], rows_np, k, 0.0), f"stream {i} rep {rep}")
        assert_same(got_sync, OC.c_search(qs[rep % 3], rows_np, k, 0.0), "host-buffer call in between")


def test_peer_flag_timeout_is_an_error_not_a_dead_context(monkeypatch):
    """Fault injection: a fused search whose peer never shows up must end in an error
    value — GSB_COUNT_ERROR in *n, raised by wait_host — and leave the context and the database usable."""
    import ctypes as C
    import torch
    from gpusimilarity_b200._lib import Exchange, Sink, check, lib
    monkeypatch.setenv("GSB_SPIN_TIMEOUT_MS", "200")
    n, k = 200_000, 100
    rows_np = O.synth_db(909, n, 32, 0)
    shard = gsb.FingerprintDB.synthetic(n, device=0, seed=909)
    dev = torch.device("cuda", 0)
    nbytes = C.c_uint64(0)
    check(lib().gsb_exchange_bytes(2, k, C.byref(nbytes)))
    mine, peer = (torch.zeros(nbytes.value, dtype=torch.uint8, device=dev) for _ in range(2))
    x = Exchange()
    x.peer_base[0], x.peer_base[1] = mine.data_ptr(), peer.data_ptr()
    x.rank, x.world, x.seq = 0, 2, 1
    out_rows = torch.zeros(k, dtype=torch.int32, device=dev)
    out_scores = torch.zeros(k, dtype=torch.float32, device=dev)
    out_n = torch.zeros(1, dtype=torch.int32, device=dev)
    out_approx = torch.zeros(1, dtype=torch.int64, device=dev)
    q = O.synth_template(909, 32)
    sink = Sink(rows=out_rows.data_ptr(), scores=out_scores.data_ptr(), n=out_n.data_ptr(), approx=out_approx.data_ptr())
    check(lib().gsb_db_search_enqueue(shard._h, torch.cuda.current_stream().cuda_stream, q.ctypes.data, None, 0, k, 0.0,
                                      C.byref(x), C.byref(sink)))
    torch.cuda.synchronize()                                   # returns: the kernel gave up after 200 ms
    assert (int(out_n.item()) & 0xffffffff) == 0xffffffff      # GSB_COUNT_ERROR
    # the context is alive and the same shard answers the next query correctly
    assert_same(shard.search_rows(q, k, 0.0), OC.c_search(q, rows_np, k, 0.0), "after the injected fault")


def test_rowpop_layout_variant(monkeypatch):
    monkeypatch.setenv("GSB_ROWPOP", "0")
    rows_np = O.synth_db(555, 70001, 32, 19)
    db = make_db(rows_np)
    for k, cutoff in ((10, 0.0), (1000, 0.0), (1000, 0.1)):
        check(db, rows_np, O.synth_template(555, 32), k, cutoff, "rowpop")
        check(db, rows_np, rows_np[69999], k, cutoff, "rowpop")


def test_tuning_knobs_do_not_change_results(monkeypatch):
    rows_np = O.synth_db(808, 150000, 32, 29)
    q = O.synth_template(808, 32)
    want = OC.c_search(q, rows_np, 1000, 0.0)
    for env in ({"GSB_STAGES": "2"}, {"GSB_STAGES": "4", "GSB_WARPS": "8"}, {"GSB_GRID": "7"},
                {"GSB_GRID": "1", "GSB_WARPS": "12"}, {"GSB_ROWPOP": "0", "GSB_WARPS": "4"},
                {"GSB_ROWPOP": "0"}, {"GSB_MIN_CAP": "8192", "GSB_WARPS": "8"}, {"GSB_SHARE_HIST": "1"},
                {"GSB_SHARE_HIST": "1", "GSB_GRID": "9"}, {"GSB_PDL": "0"}, {"GSB_PDL": "0", "GSB_SHARE_HIST": "1"},
                {"GSB_TAIL": "1"}, {"GSB_TAIL": "1", "GSB_PDL": "0"}, {"GSB_TAIL": "1", "GSB_GRID": "5"}):
        for k_, v in env.items():
            monkeypatch.setenv(k_, v)
        db = make_db(rows_np)
        assert_same(db.search_rows(q, 1000, 0.0), want, str(env))
        db.close()
        for k_ in env:
            monkeypatch.delenv(k_)


@pytest.mark.gpu
@pytest.mark.parametrize("tail", ["1", "0"])
def test_select_with_and_without_grid_barrier(monkeypatch, tail):
    """GSB_TAIL=0 (default): grid-wide arrival counter, every CTA adds its share of the final list;
    GSB_TAIL=1: the CTAs leave their candidates in global memory and go, the CTA with the last ticket
    selects out of all the lists (tail_select).  3 M rows (every SM busy), a tie group around k (crowded boundary
    bucket: the streaming path), cutoffs, k from 1 to the peeling passes, repeated and back-to-back
    asynchronous queries (both control sets, lists and histograms must be left clean)."""
    monkeypatch.setenv("GSB_TAIL", tail)
    n = 3_000_000
    rows_np = OC.c_synth_db(1234, n, 32, 1500)
    rows_np[700_000:706_000] = rows_np[11]                         # 6000 equal rows: more ties than the buffer holds
    db = make_db(rows_np)
    for rep in range(2):
        for q in (O.synth_template(1234, 32), rows_np[11], rows_np[n - 1]):
            for k, cutoff in ((1000, 0.0), (1, 0.0), (1000, 0.12), (4000, 0.0), (20000, 0.0)):
                check(db, rows_np, q, k, cutoff, f"tail {tail} rep {rep}")
    qs = [rows_np[i] for i in (3, 11, 2_999_999, 1_500_000)]
    tickets = [db.search_rows_async(q, 500, 0.0) for q in qs]
    for q, t in zip(qs, tickets):
        assert_same(db.search_rows_wait(t), OC.c_search(q, rows_np, 500, 0.0), f"tail {tail} async")
    db.close()


@pytest.mark.gpu
def test_grid_wide_threshold_sharing(monkeypatch):
    """GSB_SHARE_HIST=1: every candidate that goes through a select is counted in a grid-wide histogram
    whose k-th bucket bounds the threshold of every CTA.  3 M rows (every SM busy), ties, cutoffs, k up
    to the peeling passes (key ceiling), repeated queries (the histogram must be left clean)."""
    monkeypatch.setenv("GSB_SHARE_HIST", "1")
    n = 3_000_000
    rows_np = OC.c_synth_db(4711, n, 32, 1500)
    rows_np[500_000:500_900] = rows_np[9]                          # a tie group close to k
    db = make_db(rows_np)
    for rep in range(2):
        for q in (O.synth_template(4711, 32), rows_np[9], rows_np[n - 1]):
            for k, cutoff in ((1000, 0.0), (10, 0.0), (1000, 0.12), (3000, 0.0), (20000, 0.0)):
                check(db, rows_np, q, k, cutoff, f"shared thresholds rep {rep}")


def test_synthetic_device_generator_matches_host_twin():
    n = 100000
    for plant in (0, 37):
        db = gsb.FingerprintDB.synthetic(n, device=0, seed=42, plant_period=plant)
        rows_np = O.synth_db(42, n, 32, plant)
        for r in (0, 1, 255, 256, 31337, n - 1):
            assert np.array_equal(db.getFingerprint(r), rows_np[r]), (plant, r)
        check(db, rows_np, O.synth_template(42, 32), 1000, 0.0, f"synthetic plant={plant}")
    # a shard with a row base returns global row ids
    db = gsb.FingerprintDB.synthetic(5000, device=0, seed=42, plant_period=37, row_base=70000)
    rows_np = O.synth_db(42, 5000, 32, 37, row_base=70000)
    got = db.search_rows(O.synth_template(42, 32), 50, 0.0)
    assert_same(got, OC.c_search(O.synth_template(42, 32), rows_np, 50, 0.0, row_base=70000), "row_base")


def test_repeat_queries_are_idempotent():
    """The launch-persistent control block must be left clean by every launch."""
    rows_np = O.synth_db(3, 90000, 32, 13)
    db = make_db(rows_np)
    qs = [O.synth_template(3, 32), rows_np[5], rows_np[89999], np.zeros(32, np.int32)]
    for rep in range(3):
        for q in qs:
            for k, cutoff in ((1000, 0.0), (7, 0.25)):
                check(db, rows_np, q, k, cutoff, f"rep {rep}")
    res = db.search_batch_rows(np.stack(qs), 100, 0.05)
    for q, got in zip(qs, res):
        assert_same(got, OC.c_search(q, rows_np, 100, 0.05), "batch")


@pytest.mark.parametrize("fold", [2, 3, 4, 8, 32])
def test_folded_search(fold):
    """reference copyToGPU(fold) + re-score path, fingerprintdb_cuda.cu:170-173,246-331."""
    rows_np = O.synth_db(61, 60000, 32, 17)
    db = make_db(rows_np, fold=fold)
    assert db.foldFactor() == O.effective_fold_factor(32, fold)
    q = O.synth_template(61, 32)
    for k, cutoff in ((10, 0.0), (20, 0.3), (100, 0.6)):
        rows, scores, approx = db.search_rows(q, k, cutoff)
        w_rows, w_scores, w_approx = O.search_gpu_folded(q, rows_np, k, cutoff, fold)
        assert approx == w_approx
        assert np.array_equal(f32bits(scores), f32bits(w_scores)) and np.array_equal(rows, w_rows)


def test_config1_10m_rows_top1000():
    """BASELINE config[1]: 10 M x 1024 bit, one query, top-1000 (oracle: threaded C port)."""
    n = 10_000_000
    rows_np = OC.c_synth_db(0x5EED5EED, n, 32, 4096)
    db = make_db(rows_np)
    q = O.synth_template(0x5EED5EED, 32)
    for k, cutoff in ((1000, 0.0), (1000, 0.3), (10, 0.0)):
        check(db, rows_np, q, k, cutoff, "10M")
    check(db, rows_np, rows_np[9_999_999], 1000, 0.0, "10M row query")
    # the same database generated on the device gives the same answer
    dbs = gsb.FingerprintDB.synthetic(n, device=0, seed=0x5EED5EED, plant_period=4096)
    assert_same(dbs.search_rows(q, 1000, 0.0), OC.c_search(q, rows_np, 1000, 0.0), "10M synthetic")


@pytest.mark.skipif(not OC.ref_available(), reason="oracle/_ref not built")
def test_against_reference_cuda_path():
    """The reference's own Thrust/CUDA search (its .cu compiled verbatim for sm_100a) on the same
    rows.  Single chunk (<= 2^23 rows): identical row SET and score vector; the reference orders
    equal scores by pointer (fingerprintdb_cuda.cu:366), so rows are compared after canonical
    re-ordering."""
    n = 3_000_000
    rows_np = OC.c_synth_db(2024, n, 32, 1500)
    ref = OC.RefDB([rows_np], 1024)
    ref.copy_to_gpu(1)
    db = make_db(rows_np)
    for q in (O.synth_template(2024, 32), rows_np[12345]):
        for k, cutoff in ((10, 0.0), (1000, 0.0), (1000, 0.25), (50, 0.6)):
            r_rows, r_scores, r_approx = ref.search(q, k, cutoff)
            rows, scores, approx = db.search_rows(q, k, cutoff)
            assert approx == r_approx
            assert np.array_equal(f32bits(scores), f32bits(r_scores))
            order = O.canonical_order(r_scores, r_rows)
            assert np.array_equal(rows, r_rows[order])


@pytest.mark.skipif(not OC.ref_available(), reason="oracle/_ref not built")
@pytest.mark.parametrize("fold", [2, 4, 8])
def test_folded_search_against_reference_cuda_path(fold):
    """reference copyToGPU(F) + search (fingerprintdb_cuda.cu:168-195, 246-331) running live on this
    GPU against gsb_db_search on the same rows, and the numpy restatement against both.  One chunk:
    the result SET and the score vector must match; the reference orders equal scores by pointer."""
    n = 400_000
    rows_np = OC.c_synth_db(6100 + fold, n, 32, 250)
    ref = OC.RefDB([rows_np], 1024)
    ref.copy_to_gpu(fold)
    db = make_db(rows_np, fold=fold)
    for q in (O.synth_template(6100 + fold, 32), rows_np[4242]):
        for k, cutoff in ((10, 0.0), (100, 0.0), (20, 0.3), (100, 0.55), (1000, 0.2)):
            r_rows, r_scores, r_approx = ref.search(q, k, cutoff)
            rows, scores, approx = db.search_rows(q, k, cutoff)
            what = f"fold {fold} k {k} cutoff {cutoff}"
            assert approx == r_approx, what
            assert np.array_equal(f32bits(scores), f32bits(r_scores)), what
            assert np.array_equal(rows[O.canonical_order(scores, rows)], r_rows[O.canonical_order(r_scores, r_rows)]), what
            w_rows, w_scores, w_approx = O.search_gpu_folded(q, rows_np, k, cutoff, fold)
            assert w_approx == r_approx and np.array_equal(f32bits(w_scores), f32bits(r_scores)), what
            assert sorted(w_rows.tolist()) == sorted(r_rows.tolist()), what


@pytest.mark.skipif(not OC.ref_available(), reason="oracle/_ref not built")
@pytest.mark.parametrize("n", [10_000_000, 20_000_000])
def test_multi_chunk_against_reference_cuda_path(n):
    """BASELINE configs[1] the way gpusimserver holds it: chunks of 2^23 rows (1 GiB, what
    gpusim_createdb writes), searched by the reference's own CUDA path and by this engine from the
    same chunk list.  Checked with the cross-chunk rules of SURVEY App. D (i)-(iii); a tie group
    planted across two chunk boundaries makes the k-th score a tie on purpose."""
    chunk = 1 << 23
    rows_np = OC.c_synth_db(n, n, 32, 4096)
    q = O.synth_template(n, 32)
    dup = rows_np[777].copy()                        # 60 identical rows straddling both chunk boundaries
    for b in range(1, (n + chunk - 1) // chunk):
        rows_np[b * chunk - 30:b * chunk + 30] = dup
    chunks = [rows_np[a:a + chunk] for a in range(0, n, chunk)]
    ref = OC.RefDB(chunks, 1024)
    ref.copy_to_gpu(1)
    db = gsb.FingerprintDB(1024, n, "pass", chunks)
    db.copyToGPU(1)
    for query, cases in ((q, ((1000, 0.0), (1000, 0.3), (10, 0.0))), (dup, ((40, 0.0), (1000, 0.0), (25, 0.9)))):
        for k, cutoff in cases:
            got = db.search_rows(query, k, cutoff)
            assert_appendix_d(got, ref.search(query, k, cutoff), query, rows_np, cutoff, f"n={n} k={k} cutoff={cutoff}")
            assert_same(got, OC.c_search(query, rows_np, k, cutoff), f"oracle n={n} k={k} cutoff={cutoff}")
    for r in (0, chunk - 1, chunk, chunk + 1, n - 1):   # row lookup across chunk boundaries (a8)
        assert np.array_equal(db.getFingerprint(r), rows_np[r])


@pytest.mark.parametrize("n", [200_000_000, 1_000_000_000])
def test_full_size_against_streamed_oracle(n):
    """BASELINE configs[2] (1 B x 1024 bit, one query, top-1000) bit for bit: the oracle streams the
    same synthetic rows through its scorer on the host cores (generated on the fly, never stored) and
    must return the same rows, f32 score bits and approximate counts as the one-launch GPU search."""
    seed, plant, k = 77, 50000, 1000
    q = O.synth_template(seed, 32)
    whole = gsb.FingerprintDB.synthetic(n, device=0, seed=seed, plant_period=plant)
    for query, cutoff in ((q, 0.0), (q, 0.08), (whole.getFingerprint(n - 7), 0.0)):
        got = whole.search_rows(query, k, cutoff)
        assert_same(got, OC.c_stream_search(query, seed, plant, n, k, cutoff), f"streamed oracle n={n} cutoff={cutoff}")
    whole.close()


@pytest.mark.parametrize("n", [200_000_000, 1_000_000_000])
def test_partition_invariance_at_full_size(n):
    """Size-independent properties at sizes no oracle run covers (1 B rows = BASELINE's headline
    configuration, 130 GB in HBM): the top-k of the whole database equals the merge of the top-k
    of its parts, results are sorted and unique, and every probed score re-scores exactly."""
    parts, k = 4, 1000
    seed, plant = 77, 50000
    q = O.synth_template(seed, 32)
    whole = gsb.FingerprintDB.synthetic(n, device=0, seed=seed, plant_period=plant)
    rows, scores, approx = whole.search_rows(q, k, 0.0)
    assert approx == n and len(rows) == k
    assert np.all(np.diff(scores) <= 0) and len(set(rows.tolist())) == k
    sample = np.concatenate([rows[:20], rows[-20:]])
    fps = np.stack([whole.getFingerprint(int(r)) for r in sample])
    assert np.array_equal(fps, O.synth_rows(seed, sample, 32, plant))
    assert np.array_equal(f32bits(O.tanimoto_scores_gpu(q, fps, 0.0)),
                          f32bits(np.concatenate([scores[:20], scores[-20:]])))
    c_rows, c_scores, c_approx = whole.search_rows(q, k, 0.08)       # survivors add up over the parts
    # BASELINE configs[4] at this size: one pass of the bit-sliced multi-query kernel must give, query
    # by query, what the single-query kernel gives (rows, score bits, survivor counts)
    rng = np.random.default_rng(n)
    qs = np.stack([q] + [whole.getFingerprint(int(r)) for r in rng.integers(0, n, 23)])
    for cutoff in (0.0, 0.08):
        batched = whole.search_batch_rows(qs, 100, cutoff)
        for j in range(0, len(qs), 3 if cutoff else 1):
            assert_same(batched[j], whole.search_rows(qs[j], 100, cutoff), f"full-size batch, query {j}, cutoff {cutoff}")
    whole.close()
    per = n // parts
    cand_rows, cand_scores, part_approx = [], [], 0
    for p in range(parts):
        part = gsb.FingerprintDB.synthetic(per, device=0, seed=seed, plant_period=plant, row_base=p * per)
        r, s, _ = part.search_rows(q, k, 0.0)
        cand_rows.append(r)
        cand_scores.append(s)
        part_approx += part.search_rows(q, k, 0.08)[2]
        part.close()
    cr, cs = np.concatenate(cand_rows), np.concatenate(cand_scores)
    order = O.canonical_order(cs, cr)[:k]
    assert np.array_equal(cr[order], rows) and np.array_equal(f32bits(cs[order]), f32bits(scores))
    assert part_approx == c_approx and 0 < c_approx < n and np.all(c_scores >= np.float32(0.08))


@pytest.mark.parametrize("bits,n_queries,k", [(1024, 37, 100), (1024, 300, 100), (256, 64, 10), (1024, 5, 512)])
def test_multi_query_kernel(bits, n_queries, k):
    """BASELINE config 5 in small: gsb_db_search_batch (one pass over the database for up to 256
    queries) must equal n_queries independent searches, i.e. the oracle per query."""
    words = bits // 32
    rows_np = OC.c_synth_db(bits + n_queries, 200_000, words, 97)
    rows_np[1000:1100] = rows_np[5]                                 # a tie group
    rows_np[2000:2050] = 0                                          # empty fingerprints
    db = make_db(rows_np)
    rng = np.random.default_rng(n_queries)
    qs = np.stack([O.synth_template(bits + n_queries, words), np.zeros(words, np.int32), rows_np[5]] +
                  [rows_np[i] for i in rng.integers(0, 200_000, n_queries - 3)])
    for cutoff in (0.0, 0.15):
        res = db.search_batch_rows(qs, k, cutoff)
        assert len(res) == n_queries
        for j, got in enumerate(res):
            assert_same(got, OC.c_search(qs[j], rows_np, k, cutoff), f"batch query {j} cutoff {cutoff}")


def _sliced_queries(rows_np, seed, n_random):
    """Query mix for the bit-sliced kernel: sparse rows, the template, an empty query, an all-ones
    query, half-dense random queries (long lists, 8 extra counter planes)."""
    rng = np.random.default_rng(seed)
    n = rows_np.shape[0]
    qs = [O.synth_template(seed, 32), np.zeros(32, np.int32), np.full(32, -1, np.int32),
          rng.integers(-2**31, 2**31, 32).astype(np.int32), rows_np[0], rows_np[n - 1]]
    qs += [rows_np[i] for i in rng.integers(0, n, n_random)]
    return np.stack(qs)


@pytest.mark.parametrize("n_rows", [1, 33, 1024, 1025, 37889, 300001])
def test_sliced_kernel_sizes_and_ragged_tiles(monkeypatch, n_rows):
    """Bit-sliced multi-query kernel (gsb_sliced.cuh) forced on: ragged last batch and last tile,
    k above and below the row count, cutoffs that follow the threshold (<= 0) and the cutoff (> 0)."""
    monkeypatch.setenv("GSB_BATCH_KERNEL", "3")
    rows_np = OC.c_synth_db(77 + n_rows, n_rows, 32, 53)
    db = make_db(rows_np)
    qs = _sliced_queries(rows_np, 77 + n_rows, 14)
    for k in (1, 10, 100):
        for cutoff in (0.0, -1.0, 0.08, 0.3, 0.6, 1.5):
            res = db.search_batch_rows(qs, k, cutoff)
            for j, got in enumerate(res):
                assert_same(got, OC.c_search(qs[j], rows_np, k, cutoff), f"sliced n={n_rows} q{j} k={k} c={cutoff}")


@pytest.mark.parametrize("n_queries,k,grid", [(37, 100, 0), (300, 100, 0), (1030, 10, 0), (40, 512, 0), (64, 100, 8)])
def test_sliced_kernel_ties_blocks_and_groups(monkeypatch, n_queries, k, grid):
    """Tie groups, empty rows, more than 1024 queries (two passes), k = 512, and with GSB_GRID=8 a
    small grid: many tiles per CTA, thresholds shared between few CTAs."""
    monkeypatch.setenv("GSB_BATCH_KERNEL", "3")
    if grid:
        monkeypatch.setenv("GSB_GRID", str(grid))
    rows_np = OC.c_synth_db(900 + n_queries, 700_000, 32, 97)
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
            assert_same(res[j], OC.c_search(qs[j], rows_np, k, cutoff), f"sliced query {j} cutoff {cutoff}")


def test_batch_kernel_choice(monkeypatch):
    """Host-side dispatch of multi-query searches: queries per pass over the database."""
    import ctypes as C
    from gpusimilarity_b200._lib import check, lib

    def max_queries(db, k, nq, cutoff):
        out = C.c_uint32(0)
        check(lib().gsb_db_batch_max_queries(db._h, k, nq, cutoff, C.byref(out)))
        return out.value
    monkeypatch.delenv("GSB_BATCH_KERNEL", raising=False)
    wide = make_db(O.synth_db(1, 100, 32, 0))
    narrow = make_db(O.synth_db(1, 100, 8, 0))
    assert max_queries(wide, 100, 1024, 0.0) == 1024      # bit-sliced kernel
    assert max_queries(wide, 100, 1024, 0.1) == 1024      # ... for every cutoff
    assert max_queries(wide, 100, 8, 0.0) == 1024
    assert max_queries(wide, 100, 4, 0.0) == 256          # few queries: POPC kernel
    assert max_queries(narrow, 100, 1024, 0.0) == 1024    # narrow rows too
    monkeypatch.setenv("GSB_BATCH_KERNEL", "2")
    assert max_queries(wide, 100, 1024, 0.0) == 256
    monkeypatch.setenv("GSB_BATCH_KERNEL", "3")
    assert max_queries(wide, 100, 2, 0.0) == 1024


@pytest.mark.parametrize("bits", [128, 256, 480, 512])
def test_sliced_kernel_narrow_rows(monkeypatch, bits):
    """Rows narrower than 1024 bits (16, 8 or 4 device words; 480 bits are padded to 512): smaller
    tiles, gangs of batches per warp in the transposition.  Both multi-query kernels and the oracle."""
    words = bits // 32
    rows_np = OC.c_synth_db(bits, 150_000, words, 59)
    rows_np[500:560] = rows_np[3]                                   # a tie group
    rows_np[900:910] = 0
    db = make_db(rows_np)
    rng = np.random.default_rng(bits)
    qs = np.stack([O.synth_template(bits, words), np.zeros(words, np.int32), np.full(words, -1, np.int32),
                   rng.integers(-2**31, 2**31, words).astype(np.int32), rows_np[3]] +
                  [rows_np[i] for i in rng.integers(0, 150_000, 19)])
    for cutoff in (0.0, 0.3):
        monkeypatch.setenv("GSB_BATCH_KERNEL", "3")
        sliced = db.search_batch_rows(qs, 50, cutoff)
        monkeypatch.setenv("GSB_BATCH_KERNEL", "2")
        popc = db.search_batch_rows(qs, 50, cutoff)
        for j in range(len(qs)):
            want = OC.c_search(qs[j], rows_np, 50, cutoff)
            assert_same(sliced[j], want, f"{bits}-bit rows, bit-sliced, query {j}, cutoff {cutoff}")
            assert_same(popc[j], want, f"{bits}-bit rows, POPC kernel, query {j}, cutoff {cutoff}")
    # many sparse queries: coarse scores, huge tie groups, lists long enough for the CTA-wide select
    # rounds (whose staging area borrows the tile buffer — smaller than that for 128-bit rows)
    many = np.stack([rows_np[i] for i in rng.integers(0, 150_000, 200)])
    monkeypatch.setenv("GSB_BATCH_KERNEL", "3")
    sliced = db.search_batch_rows(many, 100, 0.0)
    monkeypatch.setenv("GSB_BATCH_KERNEL", "2")
    popc = db.search_batch_rows(many, 100, 0.0)
    for j in range(len(many)):
        assert_same(sliced[j], popc[j], f"{bits}-bit rows, 200 sparse queries, query {j}")


def test_sliced_kernel_padded_width(monkeypatch):
    """992-bit fingerprints live in 1024-bit device rows (zero padded): the bit-sliced kernel applies."""
    rows_np = OC.c_synth_db(17, 120_000, 31, 43)
    db = make_db(rows_np)
    rng = np.random.default_rng(17)
    qs = np.stack([O.synth_template(17, 31), np.zeros(31, np.int32), np.full(31, -1, np.int32)] +
                  [rows_np[i] for i in rng.integers(0, 120_000, 21)])
    for cutoff in (0.0, 0.2):
        res = db.search_batch_rows(qs, 64, cutoff)
        for j, got in enumerate(res):
            assert_same(got, OC.c_search(qs[j], rows_np, 64, cutoff), f"992-bit query {j} cutoff {cutoff}")


def test_sliced_kernel_dense_queries_span_list_blocks(monkeypatch):
    """96 half-dense queries: ~49 k list entries, three shared-memory list blocks per tile."""
    monkeypatch.setenv("GSB_BATCH_KERNEL", "3")
    rows_np = OC.c_synth_db(31, 150_000, 32, 61)
    db = make_db(rows_np)
    rng = np.random.default_rng(31)
    qs = rng.integers(-2**31, 2**31, (96, 32)).astype(np.int32)
    qs[10] = rows_np[99]
    res = db.search_batch_rows(qs, 50, 0.0)
    for j in range(0, 96, 5):
        assert_same(res[j], OC.c_search(qs[j], rows_np, 50, 0.0), f"dense query {j}")


def test_sliced_kernel_full_grid_against_popc_kernel(monkeypatch):
    """3 M rows, one CTA per SM, every warp count: the bit-sliced kernel against the POPC multi-query
    kernel and the oracle."""
    rows_np = OC.c_synth_db(8, 3_000_000, 32, 1500)
    db = make_db(rows_np)
    qs = _sliced_queries(rows_np, 8, 26)
    monkeypatch.setenv("GSB_BATCH_KERNEL", "2")
    want = db.search_batch_rows(qs, 100, 0.0)
    monkeypatch.setenv("GSB_BATCH_KERNEL", "3")
    for warps in ("32", "24", "16"):
        monkeypatch.setenv("GSB_SLICED_WARPS", warps)
        got = db.search_batch_rows(qs, 100, 0.0)
        for j in range(len(qs)):
            assert_same(got[j], want[j], f"{warps} warps vs popc, query {j}")
    for j in (0, 1, 2, 3, 9):
        assert_same(want[j], OC.c_search(qs[j], rows_np, 100, 0.0), f"oracle, query {j}")


def test_multi_query_matches_single_query_path(monkeypatch):
    rows_np = OC.c_synth_db(4, 500_000, 32, 211)
    db = make_db(rows_np)
    qs = np.stack([rows_np[i * 1000] for i in range(20)])
    batched = db.search_batch_rows(qs, 100, 0.0)
    monkeypatch.setenv("GSB_BATCH_KERNEL", "0")
    looped = db.search_batch_rows(qs, 100, 0.0)
    for a, b in zip(batched, looped):
        assert_same(a, b, "batch kernel vs loop")


def test_device_entry_points_and_merge_kernel():
    """The one-process-per-GPU path on one GPU: three shards with row bases, searched through
    gsb_db_search_device on a torch stream, merged by gsb_merge_device."""
    import torch
    from gpusimilarity_b200.dist import ShardedSearcher, shard_range, RECORD_EXTRA
    n, k, seed, plant = 900_000, 1000, 21, 700
    whole_np = OC.c_synth_db(seed, n, 32, plant)
    q_np = O.synth_template(seed, 32)
    dev = torch.device("cuda", 0)
    d_q = torch.from_numpy(q_np.copy()).to(dev)
    stream = torch.cuda.current_stream()
    for cutoff in (0.0, 0.2):
        recs, approx = [], 0
        for r in range(3):
            base, rows = shard_range(n, r, 3)
            shard = gsb.FingerprintDB.synthetic(rows, device=0, seed=seed, plant_period=plant, row_base=base)
            s = ShardedSearcher(shard, k, 0)
            s.search_local(d_q.data_ptr(), cutoff, stream)
            torch.cuda.synchronize()
            recs.append(s.rec.clone())
            approx += s.approx_count()
            shard.close()
        gathered = torch.cat(recs)
        out_rows = torch.zeros(k, dtype=torch.int32, device=dev)
        out_scores = torch.zeros(k, dtype=torch.float32, device=dev)
        out_n = torch.zeros(1, dtype=torch.int32, device=dev)
        gsb.merge_device(0, stream.cuda_stream, gathered.data_ptr(), None, 3, k + RECORD_EXTRA, k,
                         out_rows.data_ptr(), out_scores.data_ptr(), out_n.data_ptr())
        torch.cuda.synchronize()
        cnt = int(out_n.item())
        got = (out_rows[:cnt].cpu().numpy().astype(np.int64) & 0xffffffff, out_scores[:cnt].cpu().numpy(), approx)
        assert_same(got, OC.c_search(q_np, whole_np, k, cutoff), f"3 shards cutoff={cutoff}")


def test_sharded_multi_query_merge():
    """Config 5 plumbing on one GPU: three shards searched with the multi-query kernel, records
    concatenated as an all-gather would, merged by gsb_merge_batch_device (one CTA per query)."""
    import torch
    from gpusimilarity_b200.dist import ShardedBatchSearcher, shard_range
    from gpusimilarity_b200._lib import check, lib
    n, k, seed, plant, nq, world = 450_000, 100, 55, 300, 24, 3
    whole_np = OC.c_synth_db(seed, n, 32, plant)
    qs = np.stack([O.synth_template(seed, 32)] + [whole_np[i * 997] for i in range(nq - 1)])
    dev = torch.device("cuda", 0)
    d_q = torch.from_numpy(qs.copy()).to(dev)
    stream = torch.cuda.current_stream()
    for cutoff in (0.0, 0.25):
        recs = []
        for r in range(world):
            base, rows = shard_range(n, r, world)
            shard = gsb.FingerprintDB.synthetic(rows, device=0, seed=seed, plant_period=plant, row_base=base)
            s = ShardedBatchSearcher(shard, k, 0)
            p = s.rec.data_ptr()
            check(lib().gsb_db_search_batch_device(shard._h, stream.cuda_stream, d_q.data_ptr(), nq, k, cutoff, p,
                                                   p + 8 * nq * (k + 1), p + 8 * nq * k))
            torch.cuda.synchronize()
            recs.append(s.rec[:nq * (k + 2)].clone())
            shard.close()
        gathered = torch.cat(recs)
        merger = ShardedBatchSearcher.__new__(ShardedBatchSearcher)
        out_rows = torch.zeros(nq * k, dtype=torch.int32, device=dev)
        out_scores = torch.zeros(nq * k, dtype=torch.float32, device=dev)
        out_n = torch.zeros(nq, dtype=torch.int32, device=dev)
        out_approx = torch.zeros(nq, dtype=torch.int64, device=dev)
        check(lib().gsb_merge_batch_device(0, stream.cuda_stream, gathered.data_ptr(), world, nq, k, out_rows.data_ptr(),
                                           out_scores.data_ptr(), out_n.data_ptr(), out_approx.data_ptr()))
        torch.cuda.synchronize()
        rows_h = out_rows.cpu().numpy().astype(np.int64).reshape(nq, k) & 0xffffffff
        scores_h = out_scores.cpu().numpy().reshape(nq, k)
        for j in range(nq):
            c = int(out_n[j].item())
            assert_same((rows_h[j, :c], scores_h[j, :c], int(out_approx[j].item())),
                        OC.c_search(qs[j], whole_np, k, cutoff), f"sharded batch query {j}")


def test_fused_peer_exchange_two_ranks_on_one_gpu(monkeypatch):
    """gsb_db_search_device_fused with two 'ranks' in one process: two shards, two streams, small
    grids so that both launches are resident at once; the exchange buffers are plain device
    allocations here (on the 8-GPU box they are NVLink peer mappings)."""
    import ctypes as C
    import torch
    from gpusimilarity_b200._lib import Exchange, check, lib
    from gpusimilarity_b200.dist import shard_range
    monkeypatch.setenv("GSB_GRID", "12")
    n, k, seed, plant, world = 600_000, 1000, 33, 500, 2
    whole_np = OC.c_synth_db(seed, n, 32, plant)
    dev = torch.device("cuda", 0)
    nbytes = C.c_uint64(0)
    check(lib().gsb_exchange_bytes(world, k, C.byref(nbytes)))
    xbufs = [torch.zeros(nbytes.value, dtype=torch.uint8, device=dev) for _ in range(world)]
    shards, outs, streams = [], [], []
    for r in range(world):
        base, rows = shard_range(n, r, world)
        shards.append(gsb.FingerprintDB.synthetic(rows, device=0, seed=seed, plant_period=plant, row_base=base))
        outs.append((torch.zeros(k, dtype=torch.int32, device=dev), torch.zeros(k, dtype=torch.float32, device=dev),
                     torch.zeros(1, dtype=torch.int32, device=dev), torch.zeros(1, dtype=torch.int64, device=dev)))
        streams.append(torch.cuda.Stream())
    # allocate every shard's workspace up front: cudaMalloc inside the first search would wait
    # for the other rank's spinning kernel (one process, one GPU here)
    warm = torch.zeros(k + 2, dtype=torch.int64, device=dev)
    d_q0 = torch.from_numpy(O.synth_template(seed, 32).copy()).to(dev)
    for r in range(world):
        shards[r].search_device(0, d_q0.data_ptr(), k, 0.0, warm.data_ptr(), warm.data_ptr() + 8 * (k + 1),
                                warm.data_ptr() + 8 * k)
    torch.cuda.synchronize()
    seq = 0
    for q_np, cutoff in ((O.synth_template(seed, 32), 0.0), (whole_np[77], 0.0), (O.synth_template(seed, 32), 0.3),
                         (whole_np[599_999], 0.05)):
        seq += 1
        d_q = torch.from_numpy(np.ascontiguousarray(q_np)).to(dev)
        torch.cuda.synchronize()
        for r in range(world):
            x = Exchange()
            for j in range(world):
                x.peer_base[j] = xbufs[j].data_ptr()
            x.rank, x.world, x.seq = r, world, seq
            o = outs[r]
            check(lib().gsb_db_search_device_fused(shards[r]._h, streams[r].cuda_stream, d_q.data_ptr(), k, cutoff,
                                                   C.byref(x), o[0].data_ptr(), o[1].data_ptr(), o[2].data_ptr(),
                                                   o[3].data_ptr()))
        torch.cuda.synchronize()
        want = OC.c_search(q_np, whole_np, k, cutoff)
        for r in range(world):
            o = outs[r]
            cnt = int(o[2].item())
            got = (o[0][:cnt].cpu().numpy().astype(np.int64) & 0xffffffff, o[1][:cnt].cpu().numpy(), int(o[3].item()))
            assert_same(got, want, f"fused rank {r} seq {seq}")


def _shard_device_lists():
    lists = [[0, 0, 0], [0, 0, 0, 0, 0]]     # several shards on ONE device: what a 1-GPU box can run
    if gsb.get_gpu_count() >= 2:
        lists.append(list(range(min(gsb.get_gpu_count(), 4))))
    return lists


@pytest.mark.parametrize("devices", _shard_device_lists(), ids=lambda d: "dev" + "".join(map(str, d)))
def test_single_process_multi_device_shards(devices):
    """The reference's own multi-GPU mode: one process, chunks spread over the visible devices
    (fingerprintdb_cuda.cu:176-183), per-device results merged (:366).  A device may be listed more
    than once: every entry is a shard with its own row range, stream and workspace, so the sharding,
    the per-shard launches and the merge are exercised on a one-GPU box too."""
    n = 2_000_001
    rows_np = OC.c_synth_db(71, n, 32, 900)
    rows_np[n // 3 - 20:n // 3 + 20] = rows_np[5]          # a tie group across a shard boundary
    db = make_db(rows_np, devices=devices)
    assert db.shardCount() == len(devices)
    qs = [O.synth_template(71, 32), rows_np[0], rows_np[n - 1]]
    qs.append(rows_np[5])
    for q in qs:
        for k, cutoff in ((10, 0.0), (30, 0.0), (1000, 0.0), (1000, 0.2), (20000, 0.0)):
            check(db, rows_np, q, k, cutoff, f"{len(devices)} devices")
    for r in (0, n // 3, n - 1):
        assert np.array_equal(db.getFingerprint(r), rows_np[r])
    res = db.search_batch_rows(np.stack(qs), 100, 0.1)
    for q, got in zip(qs, res):
        assert_same(got, OC.c_search(q, rows_np, 100, 0.1), "multi-device batch")
    db2 = make_db(rows_np, fold=2, devices=devices)
    rows, scores, approx = db2.search_rows(qs[0], 10, 0.3)
    w_rows, w_scores, w_approx = O.search_gpu_folded(qs[0], rows_np, 10, 0.3, 2)
    assert approx == w_approx and np.array_equal(f32bits(scores), f32bits(w_scores))
    assert sorted(rows.tolist()) == sorted(w_rows.tolist())


def test_multi_chunk_upload():
    """Host chunks that do not line up with 32-row batches, pieces or shards."""
    n = 700_003
    rows_np = OC.c_synth_db(88, n, 32, 401)
    cuts = [0, 5, 100_001, 100_002, 333_333, n]
    db = gsb.FingerprintDB(1024, n, "pass", [rows_np[a:b] for a, b in zip(cuts[:-1], cuts[1:])])
    db.copyToGPU(1)
    for q in (O.synth_template(88, 32), rows_np[100_001], rows_np[n - 1]):
        check(db, rows_np, q, 100, 0.0, "multi-chunk")
    db.copyToGPU(4)                                                   # folded upload from the same chunks
    rows, scores, approx = db.search_rows(O.synth_template(88, 32), 10, 0.3)
    w = O.search_gpu_folded(O.synth_template(88, 32), rows_np, 10, 0.3, 4)
    assert np.array_equal(rows, w[0]) and np.array_equal(f32bits(scores), f32bits(w[1])) and approx == w[2]


def test_degenerate_inputs():
    rows_np = O.synth_db(2, 3000, 32, 0)
    db = make_db(rows_np)
    rows, scores, approx = db.search_rows(rows_np[1], 0, 0.0)             # k = 0: nothing, count still right
    assert len(rows) == 0 and approx == 3000
    rows, scores, approx = db.search_rows(rows_np[1], 0, 0.2)
    assert len(rows) == 0 and approx == OC.c_search(rows_np[1], rows_np, 1, 0.2)[2]
    empty = gsb.FingerprintDB(1024, 0, "pass", [])
    empty.copyToGPU(1)
    rows, scores, approx = empty.search_rows(rows_np[1], 10, 0.0)
    assert len(rows) == 0 and approx == 0
    one = make_db(rows_np[:1])
    check(one, rows_np[:1], rows_np[0], 5, 0.0, "one row")
    ones = np.full((200, 32), -1, dtype=np.int32)                         # every bit set: union 1024
    full = make_db(ones)
    check(full, ones, ones[0], 10, 0.0, "all ones")
    check(full, ones, rows_np[0], 10, 0.0, "all ones db, sparse query")


def test_concurrent_searches_from_threads():
    """One search in flight per database (like the reference), but callers may come from several
    threads and several databases may be searched at the same time."""
    import threading
    a_np, b_np = O.synth_db(8, 120000, 32, 31), O.synth_db(9, 90000, 32, 17)
    dbs = [(make_db(a_np), a_np, O.synth_template(8, 32)), (make_db(b_np), b_np, O.synth_template(9, 32))]
    want = {i: OC.c_search(q, rows, 500, 0.0) for i, (_, rows, q) in enumerate(dbs)}
    errors = []

    def worker(i):
        try:
            db, rows, q = dbs[i % 2]
            for _ in range(10):
                assert_same(db.search_rows(q, 500, 0.0), want[i % 2], f"thread {i}")
        except Exception as e:  # noqa: BLE001
            errors.append(e)

    threads = [threading.Thread(target=worker, args=(i,)) for i in range(6)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors[0]


def test_errors_are_loud():
    rows_np = O.synth_db(1, 1000, 32, 0)
    db = gsb.FingerprintDB(1024, 1000, "pass", [rows_np])
    with pytest.raises(gsb.GsbError):                       # search before copyToGPU
        db.search_rows(rows_np[0], 10, 0.0)
    db.copyToGPU(1)
    with pytest.raises(gsb.GsbError):                       # wrong query width
        db.search_rows(rows_np[0][:16], 10, 0.0)
    with pytest.raises(gsb.GsbError):                       # count mismatch (reference .cu:153-156)
        gsb.FingerprintDB(1024, 999, "pass", [rows_np])
    with pytest.raises(gsb.GsbError):                       # row out of range
        db.getFingerprint(1000)
    assert gsb.get_gpu_count() >= 1
    g = gsb.get_gpu_count()
    assert [gsb.get_next_gpu(1) for _ in range(2 * g)] is not None  # reference test getNextGPU
