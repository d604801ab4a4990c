"""Minimal launch loop of the tensor-core multi-query kernel for ncu.  usage: prof_tensor.py [rows] [nq] [calls]
Every call is one scan_tensor_kernel launch per 128 queries; prints the CUDA-event time per call."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import gpusimilarity_b200 as gsb
from gpusimilarity_b200._lib import check, lib
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 20_000_000
nq = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
calls = int(sys.argv[3]) if len(sys.argv) > 3 else 2
K = 100
os.environ["GSB_BATCH_KERNEL"] = "4"
dev = torch.device("cuda", 0)
db = gsb.FingerprintDB.synthetic(rows, device=0, seed=0x5EED5EED, plant_period=max(64, rows // 4000))
qs = np.stack([db.getFingerprint(int(r)) for r in np.linspace(0, rows - 1, nq).astype(np.int64)])
d_q = torch.from_numpy(qs.copy()).to(dev)
keys = torch.zeros(nq * K, dtype=torch.int64, device=dev)
cnt = torch.zeros(nq, dtype=torch.int32, device=dev)
surv = torch.zeros(nq, dtype=torch.int64, device=dev)
st = torch.cuda.current_stream()
for _ in range(calls):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    check(lib().gsb_db_search_batch_device(db._h, st.cuda_stream, d_q.data_ptr(), nq, K, 0.0, keys.data_ptr(),
                                           cnt.data_ptr(), surv.data_ptr()))
    e1.record()
    torch.cuda.synchronize()
    print(f"rows={rows} nq={nq}: {e0.elapsed_time(e1):.3f} ms per call", flush=True)
print("done", int(cnt.sum().item()))
