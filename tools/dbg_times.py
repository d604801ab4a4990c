import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gpusimilarity_b200 as gsb
from oracle import oracle as O
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
q = O.synth_template(0x5EED5EED, 32)
db = gsb.FingerprintDB.synthetic(rows, device=0, seed=0x5EED5EED, plant_period=max(64, rows // 4000))
for k in (10, 1000):
    for _ in range(3):
        db.search_rows(q, k, 0.0)
    os.environ["GSB_DEBUG_TIMES"] = "1"
    print(f"--- rows={rows} k={k}", file=sys.stderr, flush=True)
    db.search_rows(q, k, 0.0)
    os.environ["GSB_DEBUG_TIMES"] = "0"
