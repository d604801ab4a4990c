"""Generate tests/golden/* from the reference itself (run in the build container, where
/root/reference exists):

    python tests/golden/make_golden.py

* copies the reference's own fixture test/small.fsim (binary test data, not source),
* runs the reference's FingerprintDB::search_cpu / TanimotoFunctorCPU / bubble sort / fold,
  compiled verbatim into oracle/_ref/libgpusim_ref.so, on that fixture and on a seeded
  synthetic database, and freezes the outputs as JSON so the GPU box (which has no
  /root/reference) can check the oracle and the CUDA path against them.
"""
import hashlib
import json
import os
import shutil
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from gpusimilarity_b200.fsim import read_fsim  # noqa: E402
from oracle import oracle as O, oracle_c as OC  # noqa: E402

REF = os.environ.get("GSB_REFERENCE", "/root/reference")
SYNTH_SEED, SYNTH_ROWS, SYNTH_PLANT = 0x5EED5EED, 20000, 97


def bits(a):
    return [int(x) for x in np.asarray(a, dtype=np.float32).view(np.uint32)]


def main():
    OC.build(with_ref=True)
    shutil.copyfile(os.path.join(REF, "test", "small.fsim"), os.path.join(HERE, "small.fsim"))
    d = read_fsim(os.path.join(HERE, "small.fsim"))
    db = d.fingerprints()
    ref = OC.RefDB(d.fp_chunks, d.fp_bitcount, d.dbkey)
    out = {
        "fixture_sha256": hashlib.sha256(open(os.path.join(HERE, "small.fsim"), "rb").read()).hexdigest(),
        "fp_sha256": hashlib.sha256(b"".join(d.fp_chunks)).hexdigest(),
        "dbkey": d.dbkey, "fp_bitcount": d.fp_bitcount, "fp_count": d.fp_count,
        "search_cpu": [], "scores_cpu": {}, "bubble": [], "fold": [],
        # the reference test-suite's own expectations for this path (test/test_gpusim.cpp)
        "reference_tests": {
            "TestSimilarityCutoff": {"query_row": 0, "k": 10, "cutoffs": [0, 0.1, 0.3, 0.4],
                                     "result_counts": [10, 10, 3, 1],
                                     "approximate_counts": [100, 86, 3, 1]},
            "TestSearchMultiple": {"query_row": 3, "k": 10, "top_id": "ZINC00000022;:;ZINC00000022"},
            "CompareGPUtoCPU": {"query_row": 3, "return_counts": [10, 15]},
            "CPUSort": {"indices": [0, 1, 2, 3, 4, 5], "scores": [1, 3, 2, 4, 0, 7], "k": 3,
                        "idx0": 5, "score0": 7, "idx2": 1, "score2": 3},
            "FoldFingerprint": {"fp": [32, 24, 11, 7], "x2": [43, 31], "x4": [63]},
        },
    }
    for q in (0, 3, 17, 42, 99):
        for k in (1, 10, 15, 100):
            rows, scores, _ = ref.search(db[q], k, 0.0, cpu=True)
            out["search_cpu"].append({"query_row": q, "k": k, "rows": [int(r) for r in rows],
                                      "score_bits": bits(scores)})
        out["scores_cpu"][str(q)] = bits(OC.ref_score_cpu(db[q], db, 1))
    rng = np.random.default_rng(7)
    for n, k in ((6, 3), (50, 10), (200, 25)):
        sc = rng.integers(0, 12, n).astype(np.float32)
        idx, s = OC.ref_bubble_sort(np.arange(n), sc, k)
        out["bubble"].append({"scores": [float(x) for x in sc], "k": k,
                              "idx": [int(x) for x in idx[:k]], "sorted": [float(x) for x in s[:k]]})
    for f in (2, 4, 8, 16, 32):
        out["fold"].append({"row": 5, "factor": f, "folded": [int(x) for x in OC.ref_fold(db[5], f)]})

    # seeded synthetic database: pins the generator + scoring at a size beyond the fixture
    sdb = O.synth_db(SYNTH_SEED, SYNTH_ROWS, 32, SYNTH_PLANT)
    tmpl = O.synth_template(SYNTH_SEED, 32)
    sref = OC.RefDB([sdb], 1024)
    synth = {"seed": SYNTH_SEED, "rows": SYNTH_ROWS, "plant_period": SYNTH_PLANT,
             "db_sha256": hashlib.sha256(sdb.tobytes()).hexdigest(),
             "template": [int(x) for x in tmpl], "queries": []}
    for name, q in (("template", tmpl), ("row123", sdb[123]), ("zero", np.zeros(32, np.int32))):
        sc = OC.ref_score_cpu(q, sdb, 1)
        rows, scores, _ = sref.search(q, 50, 0.0, cpu=True)
        synth["queries"].append({"name": name, "scores_sha256": hashlib.sha256(sc.tobytes()).hexdigest(),
                                 "k": 50, "rows": [int(r) for r in rows], "score_bits": bits(scores)})
    out["synthetic"] = synth
    with open(os.path.join(HERE, "small_fsim_golden.json"), "w") as fh:
        json.dump(out, fh, indent=1)
    print("wrote", os.path.join(HERE, "small_fsim_golden.json"))


if __name__ == "__main__":
    main()
