"""The reference's own Thrust/CUDA search path (its fingerprintdb_cuda.cu compiled verbatim for
sm_100a, oracle/_ref) timed on this B200 next to the B200-native engine, same rows, same query.
usage: python tools/ref_cuda_bench.py [rows ...]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import gpusimilarity_b200 as gsb
from oracle import oracle as O, oracle_c as OC

CHUNK = 1 << 23            # rows per chunk: what a 1 GiB .fsim chunk holds (createdb.py:14,65)
sizes = [int(a) for a in sys.argv[1:]] or [10_000_000, 100_000_000]
K = 1000
for n in sizes:
    rows = OC.c_synth_db(0x5EED5EED, n, 32, max(64, n // 4000))
    q = O.synth_template(0x5EED5EED, 32)
    chunks = [rows[i:i + CHUNK] for i in range(0, n, CHUNK)]
    t = time.perf_counter(); ref = OC.RefDB(chunks, 1024); ref.copy_to_gpu(1); t_load = time.perf_counter() - t
    for _ in range(2):
        ref.search(q, K, 0.0)
    reps = 5
    t = time.perf_counter()
    for _ in range(reps):
        r_rows, r_scores, r_approx = ref.search(q, K, 0.0)
    t_ref = (time.perf_counter() - t) / reps
    ref.close()
    db = gsb.FingerprintDB(1024, n, "pass", chunks)
    t = time.perf_counter(); db.copyToGPU(1, devices=[0]); t_up = time.perf_counter() - t
    for _ in range(3):
        db.search_rows(q, K, 0.0)
    t = time.perf_counter()
    for _ in range(20):
        g_rows, g_scores, g_approx = db.search_rows(q, K, 0.0)
    t_b200 = (time.perf_counter() - t) / 20
    same_scores = bool(np.array_equal(np.sort(r_scores)[::-1].view(np.uint32), g_scores.view(np.uint32)))
    print(f"rows={n} ({len(chunks)} chunks) k={K}: reference Thrust/CUDA {t_ref*1e3:8.2f} ms/query "
          f"({n*128/t_ref/1e9:7.1f} GB/s, load {t_load:.1f}s) | gpusim_b200 {t_b200*1e3:7.3f} ms/query "
          f"({n*128/t_b200/1e9:7.1f} GB/s, upload {t_up:.1f}s) | speed-up {t_ref/t_b200:6.1f}x | "
          f"score vectors identical: {same_scores}, approx {r_approx}=={g_approx}", flush=True)
    db.close()
