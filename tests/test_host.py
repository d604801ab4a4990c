"""CPU-side tests: the C-ABI library loads and exports every declared symbol, host-only entry
points (fold, search_cpu, fsim reader/writer) work, and GPU entry points fail loudly without a
device instead of falling back."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT, f32bits
from oracle import oracle as O

import gpusimilarity_b200 as gsb
from gpusimilarity_b200 import _lib
from gpusimilarity_b200.fsim import FsimError, read_fsim, write_fsim


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "gpusim_b200.h")).read()
    declared = set(re.findall(r"\b(gsb_[a-z_0-9]+)\s*\(", header))
    declared -= {"gsb_db", "gsb_key", "gsb_scan_info"}
    handle = C.CDLL(_lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(handle, name), f"{name} declared in gpusim_b200.h but not exported"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert b"sm_100a" in _lib.lib().gsb_version()


def test_fold_matches_reference_known_answers(golden):
    f = golden["reference_tests"]["FoldFingerprint"]
    assert list(gsb.fold_fingerprint(f["fp"], 2)) == f["x2"]
    assert list(gsb.fold_fingerprint(f["fp"], 4)) == f["x4"]
    with pytest.raises(gsb.GsbError):
        gsb.fold_fingerprint(f["fp"], 3)


def test_cpusort_known_answer(golden):
    t = golden["reference_tests"]["CPUSort"]
    idx, sc = list(t["indices"]), [float(x) for x in t["scores"]]
    gsb.top_results_bubble_sort(idx, sc, t["k"])
    assert (idx[0], sc[0], idx[2], sc[2]) == (t["idx0"], t["score0"], t["idx2"], t["score2"])


def test_search_cpu_entry_point(golden, small_fsim, small_db):
    smiles, ids = list(small_fsim.smiles), list(small_fsim.ids)
    db = gsb.FingerprintDB(1024, 100, "pass", small_fsim.fp_chunks, smiles, ids)
    assert smiles == [] and ids == []                      # ctor takes the vectors (.cu:164-165)
    assert db.count() == 100 and db.getFingerprintBitcount() == 1024
    assert db.getFingerprintDataSize() == 12800
    assert db.getID(0) == b"ZINC00000007"
    for g in golden["search_cpu"]:
        rows, scores = db.search_cpu_rows(small_db[g["query_row"]], g["k"])
        assert list(rows) == g["rows"] and list(f32bits(scores)) == g["score_bits"]
    s, i, f = [], [], []
    db.search_cpu(small_db[3], "pass", 10, 0.0, s, i, f)
    assert i[0] == b"ZINC00000022" and len(s) == 10
    assert list(db.getFingerprint(0)[:4]) == [4104, 2, 1073807360, 0]


def test_gpu_entry_points_fail_loudly_without_device(small_fsim, small_db):
    if gsb.get_gpu_count() > 0:
        pytest.skip("a GPU is present")
    db = gsb.FingerprintDB(1024, 100, "pass", small_fsim.fp_chunks)
    with pytest.raises(gsb.GsbError) as e:
        db.copyToGPU(1)
    assert e.value.code == _lib.GSB_ERR_CUDA
    with pytest.raises(gsb.GsbError):
        db.search_rows(small_db[0], 10, 0.0)
    with pytest.raises(gsb.GsbError):
        db.search_batch_rows(np.stack([small_db[0]] * 20), 10, 0.0)   # multi-query path: no fallback either
    with pytest.raises(gsb.GsbError):
        gsb.FingerprintDB.synthetic(1000)


def test_batch_max_queries_needs_an_uploaded_database():
    db = gsb.FingerprintDB(1024, 10, "k", [O.synth_db(1, 10, 32, 0)])
    out = C.c_uint32(0)
    rc = _lib.lib().gsb_db_batch_max_queries(db._h, 100, 1024, 0.0, C.byref(out))
    assert rc == _lib.GSB_ERR_STATE and b"upload" in _lib.lib().gsb_last_error()


def test_constructor_validation():
    rows = O.synth_db(1, 10, 32, 0)
    with pytest.raises(gsb.GsbError) as e:
        gsb.FingerprintDB(1024, 11, "k", [rows])
    assert e.value.code == _lib.GSB_ERR_CORRUPT
    with pytest.raises(gsb.GsbError):
        gsb.FingerprintDB(1000, 10, "k", [rows])           # not a multiple of 32 bits


def test_fsim_round_trip(tmp_path, small_fsim):
    rows = O.synth_db(9, 1000, 32, 0)
    smiles = [b"C" * (i % 7 + 1) for i in range(1000)]
    ids = [b"ID%06d" % i for i in range(1000)]
    path = str(tmp_path / "t.fsim")
    write_fsim(path, rows, smiles, ids, dbkey="secret", chunk_bytes=40000)
    d = read_fsim(path)
    assert (d.dbkey, d.fp_bitcount, d.fp_count) == ("secret", 1024, 1000)
    assert len(d.fp_chunks) == 4 and np.array_equal(d.fingerprints(), rows)
    assert d.smiles == smiles and d.ids == ids
    # re-writing the reference fixture reproduces its content
    p2 = str(tmp_path / "small.fsim")
    write_fsim(p2, small_fsim.fingerprints(), small_fsim.smiles, small_fsim.ids, dbkey="pass")
    d2 = read_fsim(p2)
    assert d2.fp_chunks == small_fsim.fp_chunks and d2.ids == small_fsim.ids
    write_fsim(path, rows, smiles, ids, version=2)
    with pytest.raises(FsimError):
        read_fsim(path)


def test_multi_chunk_host_storage():
    """Several fingerprint chunks (as a > 1 GiB .fsim has): rows at the chunk boundaries come back
    right (the reference's getStorageAndLocalIndex is off by one there, fingerprintdb_cuda.cu:204)
    and search_cpu covers every chunk (the reference only searches the first, cpp:38,44)."""
    rows = O.synth_db(21, 1000, 32, 13)
    cuts = [0, 1, 257, 600, 1000]
    chunks = [rows[a:b] for a, b in zip(cuts[:-1], cuts[1:])]
    db = gsb.FingerprintDB(1024, 1000, "k", chunks)
    for r in (0, 1, 2, 256, 257, 258, 599, 600, 601, 999):
        assert np.array_equal(db.getFingerprint(r), rows[r]), r
    q = O.synth_template(21, 32)
    got_rows, got_scores = db.search_cpu_rows(q, 50)
    want_rows, want_scores = O.search_cpu(q, rows, 50)
    assert list(got_rows) == list(want_rows) and list(f32bits(got_scores)) == list(f32bits(want_scores))
    with pytest.raises(gsb.GsbError):
        db.getFingerprint(1000)


def test_sliced_kernel_math_on_cpu():
    """tests/cpp/test_sliced_math.cpp: one warp of the bit-sliced multi-query kernel emulated with
    the kernel's own integer helpers (tile layout, in-place transposition, bank-conflict freedom,
    carry-save counts, bit-sliced compare, filter bound) against plain popcounts."""
    import subprocess
    from conftest import ROOT
    subprocess.run(["make", "-C", ROOT, "tests/cpp/test_sliced_math"], check=True, capture_output=True)
    res = subprocess.run([os.path.join(ROOT, "tests", "cpp", "test_sliced_math")], capture_output=True, text=True,
                         timeout=300)
    assert res.returncode == 0 and "sliced math ok" in res.stdout, res.stdout + res.stderr


def test_tensor_kernel_operands_on_cpu():
    """tests/cpp/test_tensor_math.cpp: both operands of the tensor-core multi-query kernel built with
    the kernel's own helpers (bit -> byte expansion with weights 2^plane / 2^(7-plane), slab and
    tensor-memory layouts), read back the way tcgen05.mma reads the canonical K-major layout:
    D = 128 * popc(q & d) for every pair; conflict-free stores; the epilogue filter never rejects a
    pair the exact score comparison accepts."""
    import subprocess
    from conftest import ROOT
    subprocess.run(["make", "-C", ROOT, "tests/cpp/test_tensor_math"], check=True, capture_output=True)
    res = subprocess.run([os.path.join(ROOT, "tests", "cpp", "test_tensor_math")], capture_output=True, text=True,
                         timeout=300)
    assert res.returncode == 0 and "\nok (" in res.stdout, res.stdout + res.stderr
