"""Pins the CPU oracle (oracle/oracle.py numpy restatement, oracle/tanimoto_oracle.c) to the
reference: its test-suite's known answers (reference test/test_gpusim.cpp) and the frozen
outputs of its own sources compiled verbatim (tests/golden/make_golden.py)."""
import hashlib

import numpy as np
import pytest

from conftest import f32bits
from oracle import oracle as O
from oracle import oracle_c as OC


def test_fixture_identity(golden, small_fsim):
    assert golden["fp_sha256"] == hashlib.sha256(b"".join(small_fsim.fp_chunks)).hexdigest()
    assert (small_fsim.dbkey, small_fsim.fp_bitcount, small_fsim.fp_count) == ("pass", 1024, 100)
    assert len(small_fsim.smiles) == len(small_fsim.ids) == 100
    assert small_fsim.ids[0] == b"ZINC00000007"


def test_reference_cutoff_counts(golden, small_db):
    """reference test/test_gpusim.cpp:101-128 (TestSimilarityCutoff)."""
    t = golden["reference_tests"]["TestSimilarityCutoff"]
    for cutoff, n_res, n_approx in zip(t["cutoffs"], t["result_counts"], t["approximate_counts"]):
        for search in (O.search_gpu, OC.c_search):
            rows, scores, approx = search(small_db[t["query_row"]], small_db, t["k"], cutoff)
            assert len(rows) == n_res and approx == n_approx


def test_reference_gpu_equals_cpu_order(golden, small_db):
    """reference test/test_gpusim.cpp:29-69 (CompareGPUtoCPU): GPU order == CPU order."""
    t = golden["reference_tests"]["CompareGPUtoCPU"]
    for k in t["return_counts"]:
        g_rows, _, _ = O.search_gpu(small_db[t["query_row"]], small_db, k, 0.0)
        c_rows, _ = O.search_cpu(small_db[t["query_row"]], small_db, k)
        assert len(g_rows) == k and list(g_rows) == list(c_rows)


def test_reference_search_multiple(golden, small_fsim, small_db):
    """reference test/test_gpusim.cpp:71-99 (TestSearchMultiple): two copies of the DB."""
    t = golden["reference_tests"]["TestSearchMultiple"]
    rows, scores, _ = O.search_gpu(small_db[t["query_row"]], small_db, t["k"], 0.0)
    one = ([small_fsim.smiles[r] for r in rows], [small_fsim.ids[r] for r in rows], list(scores))
    smiles, ids, sc = O.search_databases([one, one], t["k"])
    assert len(smiles) == t["k"]
    assert ids[0].decode() == t["top_id"]


def test_reference_cpusort_and_fold(golden):
    t = golden["reference_tests"]["CPUSort"]
    idx, sc = list(t["indices"]), [float(x) for x in t["scores"]]
    O.top_results_bubble_sort(idx, sc, t["k"])
    assert (idx[0], sc[0], idx[2], sc[2]) == (t["idx0"], t["score0"], t["idx2"], t["score2"])
    f = golden["reference_tests"]["FoldFingerprint"]
    fp = np.array(f["fp"], dtype=np.int32)
    assert list(O.fold_fingerprint(fp, 2)) == f["x2"] and list(OC.c_fold(fp, 2)) == f["x2"]
    assert list(O.fold_fingerprint(fp, 4)) == f["x4"] and list(OC.c_fold(fp, 4)) == f["x4"]


def test_golden_search_cpu(golden, small_db):
    """Frozen outputs of the reference's own search_cpu / TanimotoFunctorCPU."""
    for g in golden["search_cpu"]:
        q = small_db[g["query_row"]]
        rows, scores = O.search_cpu(q, small_db, g["k"])
        assert list(rows) == g["rows"] and list(f32bits(scores)) == g["score_bits"]
        rows, scores, _ = OC.c_search(q, small_db, g["k"], 0.0)
        assert list(rows) == g["rows"] and list(f32bits(scores)) == g["score_bits"]
    for q, bits in golden["scores_cpu"].items():
        assert list(f32bits(O.tanimoto_scores_cpu(small_db[int(q)], small_db))) == bits
        assert list(f32bits(OC.c_score(small_db[int(q)], small_db, 2))) == bits


def test_golden_bubble_and_fold(golden, small_db):
    for g in golden["bubble"]:
        idx, sc = list(range(len(g["scores"]))), list(g["scores"])
        O.top_results_bubble_sort(idx, sc, g["k"])
        assert idx[:g["k"]] == g["idx"] and sc[:g["k"]] == g["sorted"]
        # the bubble sort's first k == stable (score desc, index asc) prefix
        order = O.canonical_order(np.array(g["scores"], np.float32), np.arange(len(idx)))
        assert list(order[:g["k"]]) == g["idx"]
    for g in golden["fold"]:
        assert list(O.fold_fingerprint(small_db[g["row"]], g["factor"])) == g["folded"]
        assert list(OC.c_fold(small_db[g["row"]], g["factor"])) == g["folded"]


def test_golden_synthetic(golden):
    s = golden["synthetic"]
    db = O.synth_db(s["seed"], s["rows"], 32, s["plant_period"])
    assert hashlib.sha256(db.tobytes()).hexdigest() == s["db_sha256"]
    assert list(O.synth_template(s["seed"], 32)) == s["template"]
    queries = {"template": O.synth_template(s["seed"], 32), "row123": db[123],
               "zero": np.zeros(32, np.int32)}
    for g in s["queries"]:
        q = queries[g["name"]]
        assert hashlib.sha256(O.tanimoto_scores_cpu(q, db).tobytes()).hexdigest() == g["scores_sha256"]
        assert hashlib.sha256(OC.c_score(q, db, 3).tobytes()).hexdigest() == g["scores_sha256"]
        for rows, scores in (O.search_cpu(q, db, g["k"]), OC.c_search(q, db, g["k"], 0.0)[:2]):
            assert list(rows) == g["rows"] and list(f32bits(scores)) == g["score_bits"]


def test_c_synth_generator_matches_numpy():
    for words, plant, base in ((32, 0, 0), (32, 13, 0), (32, 13, 123456789), (8, 5, 77), (128, 3, 0)):
        a = O.synth_db(42, 3001, words, plant, row_base=base)
        b = OC.c_synth_db(42, 3001, words, plant, row_base=base, n_threads=3)
        assert np.array_equal(a, b), (words, plant, base)


def test_c_oracle_matches_numpy_semantics():
    """Edge cases the reference tests never touch: all-zero rows (0/0), k > N, ties, cutoff."""
    rng = np.random.default_rng(3)
    db = O.synth_db(11, 5000, 32, 50)
    db[7] = 0
    db[4000:4010] = db[100]                       # a tie group
    zero = np.zeros(32, np.int32)
    for q, cutoff, k in ((db[100], 0.0, 20), (db[100], 0.2, 20), (zero, 0.0, 5), (db[9], -1.0, 7000),
                         (db[9], 0.05, 7000), (db[9], 1.5, 10)):
        r1, s1, a1 = O.search_gpu(q, db, k, cutoff)
        r2, s2, a2 = OC.c_search(q, db, k, cutoff, n_threads=int(rng.integers(1, 6)))
        assert list(r1) == list(r2) and list(f32bits(s1)) == list(f32bits(s2)) and a1 == a2
    # 0/0: NaN on the CPU functor, 0 on the GPU path (fingerprintdb_cuda.cu:102)
    assert np.isnan(O.tanimoto_scores_cpu(zero, db)[7]) and np.isnan(OC.c_score(zero, db, 1)[7])
    assert O.tanimoto_scores_gpu(zero, db, 0.0)[7] == 0


def test_row_base_and_fold_search():
    db = O.synth_db(5, 3000, 32, 30)
    q = O.synth_template(5, 32)
    r, s, a = O.search_gpu(q, db, 10, 0.0, row_base=1000)
    r2, s2, a2 = OC.c_search(q, db, 10, 0.0, row_base=1000)
    assert list(r) == list(r2) and r.min() >= 1000
    for f in (2, 3, 4, 8):
        rows, scores, approx = O.search_gpu_folded(q, db, 10, 0.3, f)
        full = O.tanimoto_scores_cpu(q, db)
        assert np.all(full[rows] == scores) and np.all(scores >= np.float32(0.3))
        assert np.all(np.diff(scores) <= 0)
    assert O.effective_fold_factor(32, 3) == 4 and O.fold_candidate_count(10 ** 6, 10, 4) == 120


@pytest.mark.skipif(not OC.ref_available(), reason="oracle/_ref not built")
def test_reference_build_live(small_fsim, small_db):
    """The reference's own sources, run live, against the restatement (CPU entry points)."""
    ref = OC.RefDB(small_fsim.fp_chunks, 1024, "pass")
    for q in (0, 3, 50):
        rows, scores, approx = ref.search(small_db[q], 12, 0.0, cpu=True)
        r2, s2 = O.search_cpu(small_db[q], small_db, 12)
        assert list(rows) == list(r2) and list(f32bits(scores)) == list(f32bits(s2))
        assert approx == OC.RefDB.APPROX_UNSET          # search_cpu never writes it (cpp:38-39)
    rows, scores, approx = ref.search(small_db[0], 5, 0.0, cpu=True, dbkey="wrong")
    assert len(rows) == 0                                # key mismatch: silent empty (cpp:28-31)
    assert list(ref.get_fingerprint(0)[:4]) == [4104, 2, 1073807360, 0]
