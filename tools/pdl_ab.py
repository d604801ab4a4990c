"""Back-to-back single-query launches on one stream, with and without programmatic dependent launch
(GSB_PDL), at the sizes where the fixed cost per query shows: 10 M rows (BASELINE configs[1]) and a
125 M-row shard (configs[3]).  Device query, results in HBM, CUDA events around 50 launches.
usage: python tools/pdl_ab.py [rows ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gpusimilarity_b200 as gsb
from gpusimilarity_b200.dist import ShardedSearcher
from oracle import oracle as O
dev = torch.device("cuda", 0)
q = torch.from_numpy(O.synth_template(0x5EED5EED, 32).copy()).to(dev)
st = torch.cuda.current_stream()
peak = 6537.0
sizes = [int(a) for a in sys.argv[1:]] or [10_000_000, 125_000_000]
for rows in sizes:
    db = gsb.FingerprintDB.synthetic(rows, device=0, seed=0x5EED5EED, plant_period=max(64, rows // 4000))
    for K in (1000, 10):
        s = ShardedSearcher(db, K, 0)
        for pdl, stable in (("1", True), ("1", False), ("0", True)):
            os.environ["GSB_PDL"] = pdl
            run = lambda: s.search_local(q.data_ptr(), 0.0, st, stable=stable)
            for _ in range(5): run()
            torch.cuda.synchronize()
            best = 1e9
            for _ in range(3):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                for _ in range(50): run()
                b.record(); b.synchronize()
                best = min(best, a.elapsed_time(b) / 50)
            print(f"rows={rows} k={K} GSB_PDL={pdl} stable_query={stable}: {best:.4f} ms/query "
                  f"{rows*128/best/1e6:.0f} GB/s = {rows*128/best/1e6/peak:.3f} of the copy peak", flush=True)
    os.environ["GSB_PDL"] = "1"
    os.environ["GSB_DEBUG_TIMES"] = "1"
    db.search_rows(O.synth_template(0x5EED5EED, 32), 1000, 0.0)
    del os.environ["GSB_DEBUG_TIMES"]
    # host-buffer API: one blocking call after the other vs two in flight
    import time
    qn = O.synth_template(0x5EED5EED, 32)
    for _ in range(5): db.search_rows(qn, 1000, 0.0)
    t = time.perf_counter()
    for _ in range(100): db.search_rows(qn, 1000, 0.0)
    serial = (time.perf_counter() - t) / 100 * 1e3
    t = time.perf_counter()
    prev = db.search_rows_async(qn, 1000, 0.0)
    for _ in range(99):
        cur = db.search_rows_async(qn, 1000, 0.0); db.search_rows_wait(prev); prev = cur
    db.search_rows_wait(prev)
    piped = (time.perf_counter() - t) / 100 * 1e3
    print(f"rows={rows} host-buffer API: serial {serial:.4f} ms/query, two in flight {piped:.4f} ms/query", flush=True)
    db.close()
