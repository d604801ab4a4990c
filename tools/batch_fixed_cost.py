"""Fixed cost of a bit-sliced multi-query pass: small databases, where the per-launch part (list build,
warm-up tiles, final sorts, grid-wide merges) dominates.  usage: batch_fixed_cost.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import gpusimilarity_b200 as gsb
from gpusimilarity_b200._lib import check, lib
K = 100
dev = torch.device("cuda", 0)
st = torch.cuda.current_stream()
os.environ["GSB_BATCH_KERNEL"] = "3"
for rows in (150_000, 1_500_000, 15_000_000):
    db = gsb.FingerprintDB.synthetic(rows, device=0, seed=0x5EED5EED, plant_period=max(64, rows // 4000))
    for nq in (64, 1024):
        qs = np.stack([db.getFingerprint(int(r)) for r in np.linspace(0, rows - 1, nq).astype(np.int64)])
        d_q = torch.from_numpy(qs.copy()).to(dev)
        keys = torch.zeros(nq * K, dtype=torch.int64, device=dev)
        cnt = torch.zeros(nq, dtype=torch.int32, device=dev)
        surv = torch.zeros(nq, dtype=torch.int64, device=dev)
        def run():
            check(lib().gsb_db_search_batch_device(db._h, st.cuda_stream, d_q.data_ptr(), nq, K, 0.0, keys.data_ptr(),
                                                   cnt.data_ptr(), surv.data_ptr()))
        run(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5): run()
        b.record(); b.synchronize()
        print(f"rows={rows:9d} nq={nq:5d}: {a.elapsed_time(b) / 5:7.3f} ms per pass", flush=True)
    db.close()
