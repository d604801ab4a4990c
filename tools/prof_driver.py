"""Minimal driver for ncu: a few searches over a synthetic shard.  usage: prof_driver.py [rows] [reps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gpusimilarity_b200 as gsb
from oracle import oracle as O
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000_000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
K = 1000
dev = torch.device("cuda", 0)
db = gsb.FingerprintDB.synthetic(rows, device=0, seed=0x5EED5EED, plant_period=250000)
q = torch.from_numpy(O.synth_template(0x5EED5EED, 32).copy()).to(dev)
rec = torch.zeros(K + 2, dtype=torch.int64, device=dev)
st = torch.cuda.current_stream()
for _ in range(reps):
    db.search_device(st.cuda_stream, q.data_ptr(), K, 0.0, rec.data_ptr(), rec.data_ptr() + 8 * (K + 1), rec.data_ptr() + 8 * K)
torch.cuda.synchronize()
print("n =", int(rec[K + 1].item()) & 0xffffffff)
