"""Load-time path: host rows -> HBM layout (gsb_db_upload), unfolded and folded."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import gpusimilarity_b200 as gsb
from oracle import oracle_c as OC
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20_000_000
rows = OC.c_synth_db(1, n, 32, 1000)
t = time.perf_counter(); db = gsb.FingerprintDB(1024, n, "k", [rows]); t_create = time.perf_counter() - t
for fold in (1, 1, 2, 4):
    t = time.perf_counter(); db.copyToGPU(fold); dt = time.perf_counter() - t
    print(f"rows={n} fold={fold}: create(copy) {t_create:.2f}s  upload {dt*1e3:.1f} ms = {n*128/dt/1e9:.1f} GB/s of source rows", flush=True)
