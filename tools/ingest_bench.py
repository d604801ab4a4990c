"""Load-time path: host rows -> HBM layout (gsb_db_upload), unfolded and folded, and the
whole .fsim ingest (gsb_fsim_open: read + parallel inflate straight into pinned chunks; adopt; upload).
usage: python tools/ingest_bench.py [rows] [fsim_rows]"""
import ctypes as C, os, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import gpusimilarity_b200 as gsb
from gpusimilarity_b200._lib import check, lib
from gpusimilarity_b200.fsim import write_fsim
from oracle import oracle_c as OC
n = int(sys.argv[1]) if len(sys.argv) > 1 else 50_000_000
n_fsim = int(sys.argv[2]) if len(sys.argv) > 2 else 4_000_000
rows = OC.c_synth_db(1, n, 32, 1000)
for pinned in ("1", "0"):
    os.environ["GSB_PINNED_HOST"] = pinned
    t = time.perf_counter(); db = gsb.FingerprintDB(1024, n, "k", [rows[i:i + (1 << 23)] for i in range(0, n, 1 << 23)])
    t_create = time.perf_counter() - t
    for fold in (1, 1, 2, 4):
        t = time.perf_counter(); db.copyToGPU(fold, devices=[0]); dt = time.perf_counter() - t
        print(f"rows={n} pinned_chunks={pinned} fold={fold}: create(copy) {t_create:.2f}s = {n*128/t_create/1e9:.1f} GB/s; "
              f"upload {dt*1e3:.1f} ms = {n*128/dt/1e9:.1f} GB/s of source rows", flush=True)
    db.close()
os.environ["GSB_PINNED_HOST"] = "1"
# .fsim end to end
with tempfile.TemporaryDirectory() as tmp:
    path = os.path.join(tmp, "synth.fsim")
    sub = rows[:n_fsim]
    t = time.perf_counter()
    write_fsim(path, sub, [b"C"] * n_fsim, [b"Z"] * n_fsim, chunk_bytes=64 << 20)
    print(f"wrote {path}: {os.path.getsize(path)/1e6:.0f} MB for {n_fsim*128/1e6:.0f} MB of rows in {time.perf_counter()-t:.1f}s", flush=True)
    for rep in range(2):
        f, h = C.c_void_p(), C.c_void_p()
        t = time.perf_counter(); check(lib().gsb_fsim_open(path.encode(), C.byref(f))); t_open = time.perf_counter() - t
        t = time.perf_counter(); check(lib().gsb_fsim_create_db(f, C.byref(h))); t_adopt = time.perf_counter() - t
        t = time.perf_counter(); check(lib().gsb_db_upload(h, None, 0, 1)); t_up = time.perf_counter() - t
        gb = n_fsim * 128 / 1e9
        print(f".fsim ingest rep {rep}: open+inflate {t_open:.2f}s ({gb/t_open:.2f} GB/s of rows, {os.cpu_count()} host threads), "
              f"adopt {t_adopt*1e3:.1f} ms, upload {t_up*1e3:.1f} ms ({gb/t_up:.1f} GB/s)", flush=True)
        lib().gsb_db_destroy(h); lib().gsb_fsim_close(f)
