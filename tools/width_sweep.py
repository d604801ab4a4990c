"""Scan bandwidth for every supported fingerprint width (synthetic shards of ~12.8 GB)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gpusimilarity_b200 as gsb
from oracle import oracle as O
dev = torch.device("cuda", 0)
st = torch.cuda.current_stream()
K = 1000
rec = torch.zeros(K + 2, dtype=torch.int64, device=dev)
for bits in (128, 256, 512, 1024, 2048, 4096):
    words = bits // 32
    rows = int(12.8e9 // (bits // 8))
    db = gsb.FingerprintDB.synthetic(rows, device=0, fp_bitcount=bits, seed=1, plant_period=100000)
    q = torch.from_numpy(O.synth_template(1, words).copy()).to(dev)
    info = db.scan_info(K)
    run = lambda: db.search_device(st.cuda_stream, q.data_ptr(), K, 0.0, rec.data_ptr(), rec.data_ptr() + 8 * (K + 1), rec.data_ptr() + 8 * K)
    for _ in range(3): run()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10): run()
    b.record(); b.synchronize()
    ms = a.elapsed_time(b) / 10
    print(f"bits={bits:5d} rows={rows:11d} block={info.block} stages={info.stages} smem={info.smem_bytes}: {ms:7.3f} ms "
          f"{rows * (bits // 8) / ms / 1e6:7.0f} GB/s algorithmic, {rows / ms / 1e6:7.1f} G rows/s", flush=True)
    db.close()
