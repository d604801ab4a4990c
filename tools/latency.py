"""Latency vs database size: separates the fixed cost of a query from the per-row cost.
usage: python tools/latency.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import gpusimilarity_b200 as gsb
from oracle import oracle as O

dev = torch.device("cuda", 0)
q_np = O.synth_template(0x5EED5EED, 32)
q = torch.from_numpy(q_np.copy()).to(dev)
stream = torch.cuda.current_stream()
print("rows        k    device_ms   GB/s(alg)   host_api_ms")
for rows in (100_000, 1_000_000, 2_000_000, 5_000_000, 10_000_000, 20_000_000, 50_000_000):
    db = gsb.FingerprintDB.synthetic(rows, device=0, seed=0x5EED5EED, plant_period=max(64, rows // 4000))
    for K in (10, 100, 1000):
        rec = torch.zeros(K + 2, dtype=torch.int64, device=dev)
        def run():
            db.search_device(stream.cuda_stream, q.data_ptr(), K, 0.0, rec.data_ptr(), rec.data_ptr() + 8 * (K + 1),
                             rec.data_ptr() + 8 * K)
        for _ in range(5):
            run()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 20
        a.record()
        for _ in range(n):
            run()
        b.record(); b.synchronize()
        ms = a.elapsed_time(b) / n
        for _ in range(3):
            db.search_rows(q_np, K, 0.0)
        t = time.perf_counter()
        for _ in range(n):
            db.search_rows(q_np, K, 0.0)
        host_ms = (time.perf_counter() - t) / n * 1e3
        print(f"{rows:10d} {K:5d} {ms:9.4f} {rows * 128 / ms / 1e6:10.1f} {host_ms:11.4f}", flush=True)
    db.close()
