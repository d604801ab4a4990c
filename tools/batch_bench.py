"""Throughput of the multi-query kernels (bit-sliced and POPC) vs looping single queries.
usage: batch_bench.py [rows] [nq] [k]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import gpusimilarity_b200 as gsb
from gpusimilarity_b200._lib import check, lib
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
nq = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
K = int(sys.argv[3]) if len(sys.argv) > 3 else 100
fast = len(sys.argv) > 4  # any 4th argument: bit-sliced kernel only
dev = torch.device("cuda", 0)
db = gsb.FingerprintDB.synthetic(rows, device=0, seed=0x5EED5EED, plant_period=max(64, rows // 4000))
qrows = np.linspace(0, rows - 1, nq).astype(np.int64)
qs = np.stack([db.getFingerprint(int(r)) for r in qrows])
print(f"mean set bits per query: {np.unpackbits(qs.view(np.uint8)).sum() / nq:.1f}")
d_q = torch.from_numpy(qs.copy()).to(dev)
keys = torch.zeros(nq * K, dtype=torch.int64, device=dev)
cnt = torch.zeros(nq, dtype=torch.int32, device=dev)
surv = torch.zeros(nq, dtype=torch.int64, device=dev)
st = torch.cuda.current_stream()


def run(group):
    for q0 in range(0, nq, group):
        n = min(group, nq - q0)
        check(lib().gsb_db_search_batch_device(db._h, st.cuda_stream, d_q[q0:].data_ptr(), n, K, 0.0,
                                               keys[q0 * K:].data_ptr(), cnt[q0:].data_ptr(), surv[q0:].data_ptr()))


results = {}
variants = [("bit-sliced", {"GSB_BATCH_KERNEL": "3"}, 1024)]
if not fast:
    variants += [(f"bit-sliced, {w} warps", {"GSB_BATCH_KERNEL": "3", "GSB_SLICED_WARPS": str(w)}, 1024) for w in (24, 16)]
    variants += [("popc", {"GSB_BATCH_KERNEL": "2"}, 256)]
for name, env, group in variants:
    for key in ("GSB_BATCH_KERNEL", "GSB_SLICED_WARPS"):
        os.environ.pop(key, None)
    os.environ.update(env)
    run(group); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); run(group); b.record(); b.synchronize()
    ms = a.elapsed_time(b)
    results[name] = keys.cpu().numpy().astype(np.uint64).copy()
    print(f"{name}: rows={rows} nq={nq} k={K}: {ms:.2f} ms -> {nq / ms * 1e3:.1f} q/s, "
          f"{rows * nq / ms / 1e6:.1f} G row*query/s")
if fast:
    sys.exit(0)
print("bit-sliced == popc:", bool(np.array_equal(results["bit-sliced"], results["popc"])))
rec = torch.zeros(K + 2, dtype=torch.int64, device=dev)
nloop = min(nq, 16)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); a.record()
for j in range(nloop):
    db.search_device(st.cuda_stream, d_q[j].data_ptr(), K, 0.0, rec.data_ptr(), rec.data_ptr() + 8 * (K + 1), rec.data_ptr() + 8 * K)
b.record(); b.synchronize()
ms1 = a.elapsed_time(b) / nloop
print(f"single-query kernel: {ms1:.3f} ms/query -> {1e3 / ms1:.1f} q/s")
db.search_device(st.cuda_stream, d_q[0].data_ptr(), K, 0.0, rec.data_ptr(), rec.data_ptr() + 8 * (K + 1), rec.data_ptr() + 8 * K)
torch.cuda.synchronize()
print("query 0 identical to the single-query kernel:",
      bool(np.array_equal(results["bit-sliced"][:K], rec[:K].cpu().numpy().astype(np.uint64))))
