"""First contact of the tensor-core multi-query kernel with a GPU: small cases against the oracle
and the bit-sliced kernel, with a per-query report of what differs; then timings against the
bit-sliced kernel for query densities 32 / 128 / 512 set bits.
usage: tensor_try.py [rows_for_timing]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import gpusimilarity_b200 as gsb
from oracle import oracle as O
from oracle import oracle_c as OC


def run(db, qs, k, cutoff, mode):
    os.environ["GSB_BATCH_KERNEL"] = str(mode)
    return db.search_batch_rows(qs, k, cutoff)


def compare(tag, got, want, limit=6):
    bad = 0
    for j, (g, w) in enumerate(zip(got, want)):
        same = (len(g[0]) == len(w[0]) and np.array_equal(g[0], w[0]) and
                np.array_equal(g[1].view(np.uint32), w[1].view(np.uint32)) and g[2] == w[2])
        if not same:
            bad += 1
            if bad <= limit:
                print(f"  {tag}: query {j} differs: n {len(g[0])} vs {len(w[0])}, approx {g[2]} vs {w[2]}")
                print("     got ", g[0][:6], g[1][:6])
                print("     want", w[0][:6], w[1][:6])
    print(f"{tag}: {len(got) - bad}/{len(got)} queries identical", flush=True)
    return bad == 0


def main():
    ok = True
    cases = ((128, 5, 10), (4096, 128, 100), (100_000, 37, 100), (300_001, 300, 50))
    if os.environ.get("TC_CASE"):
        cases = (cases[int(os.environ["TC_CASE"])],)
    for n_rows, nq, k in cases:
        rows_np = OC.c_synth_db(7 + n_rows, n_rows, 32, 53)
        db = gsb.FingerprintDB(1024, n_rows, "pass", [np.ascontiguousarray(rows_np, dtype=np.int32)])
        db.copyToGPU(1, None)
        rng = np.random.default_rng(n_rows)
        qs = np.stack([O.synth_template(7 + n_rows, 32), np.full(32, -1, np.int32),
                       rng.integers(-2**31, 2**31, 32).astype(np.int32), rows_np[0], rows_np[n_rows - 1]] +
                      [rows_np[i] for i in rng.integers(0, n_rows, nq - 5)])
        for cutoff in (0.0, 0.3):
            t0 = time.time()
            got = run(db, qs, k, cutoff, 4)
            dt = time.time() - t0
            want = run(db, qs, k, cutoff, 3)
            ok &= compare(f"rows={n_rows} nq={nq} k={k} cutoff={cutoff} ({dt*1e3:.1f} ms) tensor vs sliced", got, want)
            if n_rows <= 100_000:
                ora = [OC.c_search(q, rows_np, k, cutoff) for q in qs[:8]]
                ok &= compare("   tensor vs oracle (first 8)", got[:8], ora)
    print("PARITY", "OK" if ok else "FAILED", flush=True)
    if len(sys.argv) > 1:
        import torch
        n = int(sys.argv[1])
        db = gsb.FingerprintDB.synthetic(n, device=0, seed=0x5EED5EED, plant_period=250000)
        rng = np.random.default_rng(1)
        for bits in (32, 64, 128, 512):
            for nq in (128, 256, 1024):
                qs = np.zeros((nq, 32), np.int32)
                for j in range(nq):
                    pos = rng.choice(1024, bits, replace=False)
                    w = np.zeros(32, np.uint32)
                    np.bitwise_or.at(w, pos // 32, (np.uint32(1) << (pos % 32).astype(np.uint32)))
                    qs[j] = w.view(np.int32)
                line = f"rows={n} set_bits={bits} nq={nq}:"
                res = {}
                for mode, name in ((4, "tensor"), (3, "sliced")):
                    os.environ["GSB_BATCH_KERNEL"] = str(mode)
                    db.search_batch_rows_raw(qs, 100, 0.0)
                    t0 = time.time()
                    res[name] = db.search_batch_rows_raw(qs, 100, 0.0)
                    line += f" {name} {1e3 * (time.time() - t0):.1f} ms"
                same = all(np.array_equal(a, b) for a, b in zip(res["tensor"], res["sliced"]))
                print(line, "identical" if same else "DIFFERENT", flush=True)


if __name__ == "__main__":
    main()
