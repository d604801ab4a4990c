"""Multi-query kernels (bit-sliced vs POPC) for every row width they support: 100 M synthetic rows,
256 queries (database rows), top-100."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import gpusimilarity_b200 as gsb
from gpusimilarity_b200._lib import check, lib
rows, nq, K = 100_000_000, 256, 100
dev = torch.device("cuda", 0)
st = torch.cuda.current_stream()
for bits in (128, 256, 512, 1024):
    db = gsb.FingerprintDB.synthetic(rows, device=0, fp_bitcount=bits, seed=1, plant_period=100000)
    qs = np.stack([db.getFingerprint(int(r)) for r in np.linspace(0, rows - 1, nq).astype(np.int64)])
    d_q = torch.from_numpy(qs.copy()).to(dev)
    keys = torch.zeros(nq * K, dtype=torch.int64, device=dev)
    cnt = torch.zeros(nq, dtype=torch.int32, device=dev)
    surv = torch.zeros(nq, dtype=torch.int64, device=dev)
    out, res = {}, {}
    for name, mode in (("popc", "2"), ("bit-sliced", "3")):
        os.environ["GSB_BATCH_KERNEL"] = mode
        def run():
            check(lib().gsb_db_search_batch_device(db._h, st.cuda_stream, d_q.data_ptr(), nq, K, 0.0, keys.data_ptr(),
                                                   cnt.data_ptr(), surv.data_ptr()))
        run(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); run(); b.record(); b.synchronize()
        out[name] = a.elapsed_time(b)
        res[name] = keys.cpu().numpy().copy()
    print(f"bits={bits:5d} set bits/query {np.unpackbits(qs.view(np.uint8)).sum() / nq:5.1f}: popc {out['popc']:8.2f} ms   "
          f"bit-sliced {out['bit-sliced']:8.2f} ms ({rows * nq / out['bit-sliced'] / 1e6:7.1f} G row*query/s)   "
          f"identical: {bool(np.array_equal(res['popc'], res['bit-sliced']))}", flush=True)
    db.close()
