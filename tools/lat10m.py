import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gpusimilarity_b200 as gsb
from oracle import oracle as O
dev = torch.device("cuda", 0)
q = torch.from_numpy(O.synth_template(0x5EED5EED, 32).copy()).to(dev)
st = torch.cuda.current_stream()
for rows in (10_000_000, 125_000_000):
    db = gsb.FingerprintDB.synthetic(rows, device=0, seed=0x5EED5EED, plant_period=max(64, rows // 4000))
    for warps, stages in ((16, 2), (12, 3), (12, 2), (8, 4)):
        os.environ["GSB_WARPS"], os.environ["GSB_STAGES"] = str(warps), str(stages)
        for K in (10, 1000):
            rec = torch.zeros(K + 2, dtype=torch.int64, device=dev)
            run = lambda: db.search_device(st.cuda_stream, q.data_ptr(), K, 0.0, rec.data_ptr(), rec.data_ptr() + 8 * (K + 1), rec.data_ptr() + 8 * K)
            for _ in range(5): run()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(50): run()
            b.record(); b.synchronize()
            ms = a.elapsed_time(b) / 50
            print(f"rows={rows} warps={warps} stages={stages} k={K}: {ms:.4f} ms {rows*128/ms/1e6:.0f} GB/s", flush=True)
    db.close()
