"""Where the bit-sliced multi-query kernel overtakes the POPC kernel: small batches, both kernels.
usage: crossover.py [rows]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import gpusimilarity_b200 as gsb
from gpusimilarity_b200._lib import check, lib
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
K = 100
dev = torch.device("cuda", 0)
db = gsb.FingerprintDB.synthetic(rows, device=0, seed=0x5EED5EED, plant_period=max(64, rows // 4000))
st = torch.cuda.current_stream()
for nq in (2, 4, 8, 16, 32, 64):
    qs = np.stack([db.getFingerprint(int(r)) for r in np.linspace(0, rows - 1, nq).astype(np.int64)])
    d_q = torch.from_numpy(qs.copy()).to(dev)
    keys = torch.zeros(nq * K, dtype=torch.int64, device=dev)
    cnt = torch.zeros(nq, dtype=torch.int32, device=dev)
    surv = torch.zeros(nq, dtype=torch.int64, device=dev)
    out = {}
    for name, mode in (("popc", "2"), ("bit-sliced", "3")):
        os.environ["GSB_BATCH_KERNEL"] = mode
        def run():
            check(lib().gsb_db_search_batch_device(db._h, st.cuda_stream, d_q.data_ptr(), nq, K, 0.0, keys.data_ptr(),
                                                   cnt.data_ptr(), surv.data_ptr()))
        run(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); run(); run(); b.record(); b.synchronize()
        out[name] = a.elapsed_time(b) / 2
    print(f"rows={rows} nq={nq:3d}: popc {out['popc']:8.2f} ms   bit-sliced {out['bit-sliced']:8.2f} ms")
