"""Kernel-shape sweep on one GPU: times gsb_db_search_device for combinations of the env knobs.
usage: python tools/sweep.py [rows] [k]"""
import itertools, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import gpusimilarity_b200 as gsb
from oracle import oracle as O

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000_000
K = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
dev = torch.device("cuda", 0)
q = torch.from_numpy(O.synth_template(0x5EED5EED, 32).copy()).to(dev)
rec = torch.zeros(K + 2, dtype=torch.int64, device=dev)
stream = torch.cuda.current_stream()
results = []
for rowpop in (1, 0):
    os.environ["GSB_ROWPOP"] = str(rowpop)
    db = gsb.FingerprintDB.synthetic(rows, device=0, seed=0x5EED5EED, plant_period=250000)
    ref = None
    for warps, stages, unroll in itertools.product((16, 12, 8), (2, 3, 4), (1, 2)):
        os.environ["GSB_WARPS"], os.environ["GSB_STAGES"], os.environ["GSB_UNROLL"] = str(warps), str(stages), str(unroll)
        try:
            info = db.scan_info(K)
        except Exception:
            continue
        if info.block != warps * 32 or info.stages != stages:
            continue
        def run():
            db.search_device(stream.cuda_stream, q.data_ptr(), K, 0.0, rec.data_ptr(), rec.data_ptr() + 8 * (K + 1),
                             rec.data_ptr() + 8 * K)
        for _ in range(3):
            run()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 10
        a.record()
        for _ in range(n):
            run()
        b.record(); b.synchronize()
        ms = a.elapsed_time(b) / n
        keys = rec[:K].cpu().numpy()
        if ref is None:
            ref = keys
        same = bool(np.array_equal(ref, keys))
        print(f"rowpop={rowpop} warps={warps:2d} stages={stages} U={unroll} smem={info.smem_bytes:6d} cap={info.cand_capacity}: "
              f"{ms:8.3f} ms  {rows * 128 / ms / 1e6:8.1f} GB/s (alg)  {info.db_bytes_per_query / ms / 1e6:8.1f} GB/s (layout) same={same}",
              flush=True)
    db.close()
