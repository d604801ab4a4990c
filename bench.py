#!/usr/bin/env python
"""Benchmark of the hot path: single-query brute-force Tanimoto scan + top-1000 over a synthetic
1 B x 1024-bit fingerprint database (BASELINE.json configs[2]; sharded over the ranks for N > 1).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--rows R]

A step is one query over the whole database.  N > 1 is launched by torchrun (one rank per GPU):
every rank scans its contiguous shard and the per-shard candidates are exchanged and merged INSIDE
the scan launch over NVLink peer memory (NCCL all-gather + merge kernel with GSB_FUSED_EXCHANGE=0).
Prints ONE JSON line (rank 0).

  value      queries/s, device-timed (CUDA events, barrier + synchronize on both sides, max over
             ranks), query and database already resident in HBM
  e2e        the same through the host-buffer API: the query comes from host memory, rows and scores
             land in pinned host memory.  `value` keeps two queries in flight (gsb_db_search_async /
             ShardedSearcher.submit_host), `serial` is one blocking call after the other
  roofline   the scan kernel's algorithmic bytes (128 B per row) / its mean launch duration,
             against MEASURED_PEAKS.json's measured copy bandwidth
  cpu_baseline    the reference's own search_cpu (oracle/_ref, its sources compiled verbatim) on a
             bounded sample of the workload on this box's host cores; `scan_only` is the same without
             the reference's O(k N) bubble sort
  reference_cuda  (N = 1) the reference's own Thrust/CUDA path (oracle/_ref) on this GPU, 10 M and
             100 M rows in 2^23-row chunks, next to this engine on the same rows
  multi_query     BASELINE configs[4] (a batch of 1024 queries, top-100) on the same resident shards:
             device-timed, end to end with host buffers, its own roofline block, a dense-query sweep
  single_process  one process driving all N GPUs through gsb_db_search (the reference's own mode)
  verified   the timed results checked bit for bit against the oracle streaming the same rows
             through its scorer on the host (and, for N > 1, fused == NCCL == host merge)

--impl reference times the reference CPU path alone (rank 0; other ranks exit).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SEED = 0x5EED5EED
PLANT_PERIOD = 250_000        # ~4000 near-duplicates of the query in 1 B rows
K = 1000
CUTOFF = 0.0
ROW_BYTES = 128
METRIC = "queries/sec, 1Bx1024-bit DB, top-1000, single query"
CPU_SAMPLE_ROWS = 1 << 18
MQ_QUERIES, MQ_K = 1024, 100
REF_CHUNK = 1 << 23           # rows of a 1 GiB .fsim chunk (reference python/gpusim_createdb.py)


def log(*a):
    print("[bench]", *a, file=sys.stderr, flush=True)


def measured_peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, copy)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic_per_launch(rows):
    """dram bytes per launch: the committed ncu capture of this kernel, scaled by rows when the
    capture was taken at another size (the layout reads 130 B per row: 128 + the popcount trailer)."""
    try:
        with open(os.path.join(ROOT, "profiles", "scan_kernel_ncu.json")) as fh:
            prof = json.load(fh)
        per_row = (float(prof["dram_bytes_read"]) + float(prof["dram_bytes_write"])) / float(prof["rows"])
        return per_row * rows, f"ncu capture at {int(prof['rows'])} rows ({per_row:.2f} B/row), scaled by rows"
    except Exception:
        return None, None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons while the timed region runs (B200_PROFILING.md)."""
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "20",
                 "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
            t_end = time.time() + 3.0          # nvidia-smi needs a moment before its first sample
            while not self.lines and time.time() < t_end:
                time.sleep(0.01)
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.06)
        self.proc.terminate()
        sm, smax, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        # samples inside the load window; a window shorter than the sampling period falls back to
        # the nearest samples around it (the GPU is under the same load during warm-up)
        rows = [l for (t, l) in self.lines if t0 <= t <= t1 + 0.03]
        if len(rows) < 3:
            near = sorted(self.lines, key=lambda tl: abs(tl[0] - (t0 + t1) / 2))[:5]
            rows = [l for (_, l) in near]
        for line in rows:
            f = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(f[1])); smax.append(float(f[2])); power.append(float(f[3]))
                for name, v in zip(names, f[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------ CPU arm
def cpu_reference_run(steps, warmup, rows_total):
    """The reference's FingerprintDB::search_cpu (TanimotoFunctorCPU under blockingMap on every
    host thread + its partial bubble sort), on a 2^18-row sample; linear in rows, so the
    whole-database time is the sample time x rows_total / sample rows."""
    import numpy as np
    from oracle import oracle as O, oracle_c as OC
    cores = os.cpu_count() or 1
    sample = OC.c_synth_db(SEED, CPU_SAMPLE_ROWS, 32, PLANT_PERIOD // 64)
    query = O.synth_template(SEED, 32)
    if OC.ref_available():
        kind = "reference"
        ref = OC.RefDB([sample], 1024)

        def step():
            return ref.search(query, K, CUTOFF, cpu=True)

        def score_only():
            return OC.ref_score_cpu(query, sample, cores)
    else:
        kind = "port"

        def step():
            return OC.c_search(query, sample, K, CUTOFF, n_threads=cores)

        def score_only():
            return OC.c_score(query, sample, cores)
    for _ in range(warmup):
        step()
    times = []
    for _ in range(steps):
        t = time.perf_counter(); step(); times.append(time.perf_counter() - t)
    score_only()
    t_score = []
    for _ in range(5):
        t = time.perf_counter(); score_only(); t_score.append(time.perf_counter() - t)
    t_score = min(t_score)
    per_query = statistics.mean(times) * rows_total / CPU_SAMPLE_ROWS
    scan_s = t_score * rows_total / CPU_SAMPLE_ROWS
    return {
        "value": 1.0 / per_query, "unit": "queries/s", "cores": cores, "kind": kind,
        "sample": (f"search_cpu on {CPU_SAMPLE_ROWS} synthetic rows, k={K}: {statistics.mean(times) * 1e3:.1f} ms "
                   f"(score-only {CPU_SAMPLE_ROWS / t_score / 1e6:.1f} M rows/s = "
                   f"{CPU_SAMPLE_ROWS * ROW_BYTES / t_score / 1e9:.2f} GB/s); scaled linearly to "
                   f"{rows_total} rows (EXTRAPOLATED: 99.9 % of it is the reference's O(k N) bubble sort)"),
        # the scan alone (TanimotoFunctorCPU over every row, no sort): what the GPU scan replaces
        "scan_only": {"value": 1.0 / scan_s, "unit": "queries/s", "ms_per_step": scan_s * 1e3,
                      "gbs": CPU_SAMPLE_ROWS * ROW_BYTES / t_score / 1e9,
                      "note": "score-only pass on the same sample, scaled linearly in rows; no top-k"},
        "ms_per_step": per_query * 1e3,
    }


def run_reference(args, rank):
    if rank != 0:
        return
    res = cpu_reference_run(args.steps, args.warmup, args.rows)
    line = {
        "impl": "reference", "metric": METRIC, "value": res["value"], "unit": "queries/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": res["ms_per_step"],
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u32 popcount + f32 divide",
        "data": "synthetic",
        "config": {"workload": f"{args.rows} x 1024-bit synthetic fingerprints, single query, top-{K}, "
                               "reference CPU path (search_cpu) on host cores"},
        "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample", "scan_only")},
        "e2e": {"value": res["value"], "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------ reference CUDA leg
def reference_cuda_leg(gsb, O, OC, np, sizes):
    """The reference's own Thrust/CUDA search (fingerprintdb_cuda.cu compiled verbatim for sm_100a,
    oracle/_ref) on this GPU, wall clock per FingerprintDB::search call with host buffers, next to
    gsb_db_search on the same chunks; score vectors must be bit-identical."""
    out = []
    if not OC.ref_available():
        return [{"unavailable": "oracle/_ref/libgpusim_ref.so not built"}]
    try:
        avail_gb = int([l for l in open("/proc/meminfo") if l.startswith("MemAvailable")][0].split()[1]) / 1e6
    except Exception:
        avail_gb = 0.0
    for n in sizes:
        if n * ROW_BYTES * 4.5 / 1e9 > avail_gb:
            out.append({"rows": n, "skipped": f"needs ~{n * ROW_BYTES * 4.5 / 1e9:.0f} GB of host memory, "
                                              f"{avail_gb:.0f} GB available"})
            continue
        rows = OC.c_synth_db(SEED, n, 32, max(64, n // 4000))
        q = O.synth_template(SEED, 32)
        chunks = [rows[i:i + REF_CHUNK] for i in range(0, n, REF_CHUNK)]
        ref = OC.RefDB(chunks, 1024)
        ref.copy_to_gpu(1)
        ref.search(q, K, CUTOFF)
        t_ref = []
        for _ in range(4):
            t = time.perf_counter(); r_rows, r_scores, r_approx = ref.search(q, K, CUTOFF)
            t_ref.append(time.perf_counter() - t)
        ref.close()
        del ref
        t = time.perf_counter()
        db = gsb.FingerprintDB(1024, n, "pass", chunks)      # copies the chunks (as the reference ctor does), into pinned memory
        t_create = time.perf_counter() - t
        t = time.perf_counter()
        db.copyToGPU(1, devices=[0])                          # DMA out of the pinned chunks + ingest kernel (layout, popcounts)
        t_upload = time.perf_counter() - t
        for _ in range(3):
            db.search_rows(q, K, CUTOFF)
        t_own = []
        for _ in range(20):
            t = time.perf_counter(); g_rows, g_scores, g_approx = db.search_rows(q, K, CUTOFF)
            t_own.append(time.perf_counter() - t)
        same = bool(np.array_equal(np.sort(r_scores)[::-1].view(np.uint32), g_scores.view(np.uint32))) \
            and r_approx == g_approx
        ref_ms, own_ms = statistics.median(t_ref) * 1e3, statistics.median(t_own) * 1e3
        out.append({"rows": n, "chunks": len(chunks), "k": K, "reference_cuda_ms": ref_ms,
                    "reference_cuda_ms_min_max": [min(t_ref) * 1e3, max(t_ref) * 1e3],
                    "reference_cuda_gbs": n * ROW_BYTES / ref_ms / 1e6,
                    "b200_ms": own_ms, "b200_gbs": n * ROW_BYTES / own_ms / 1e6, "speedup": ref_ms / own_ms,
                    "scores_identical": same,
                    "ingest": {"source_gb": n * ROW_BYTES / 1e9, "create_s": t_create, "upload_s": t_upload,
                               "upload_gbs": n * ROW_BYTES / 1e9 / t_upload,
                               "note": "gsb_db_upload: H2D straight out of the pinned host chunks + ingest_rows_kernel; "
                                       "PCIe 5 x16 carries ~55-63 GB/s per direction"}})
        db.close()
        del db, rows, chunks
        log("reference_cuda", out[-1])
    return out


# ------------------------------------------------------------------------------------ small shards
def small_shard_leg(gsb, O, np, torch, ShardedSearcher, local_rank, sizes, peak_gbs):
    """What BASELINE configs[1] (10 M rows) and one 8-way shard of configs[2] (125 M rows) cost per
    query on one GPU, device-timed with the query resident: `stream_ms` = back to back on one stream
    (programmatic dependent launch: the next scan starts while the last CTA still selects),
    `latency_ms` = one query at a time with a synchronize in between.  The fixed cost per launch
    (ramp + grid-wide select) is what the fraction of the copy peak loses on small shards."""
    dev = torch.device("cuda", local_rank)
    stream = torch.cuda.current_stream()
    d_query = torch.from_numpy(O.synth_template(SEED, 32).copy()).to(dev)
    out = []
    for rows in sizes:
        db = gsb.FingerprintDB.synthetic(rows, device=local_rank, seed=SEED, plant_period=PLANT_PERIOD)
        s = ShardedSearcher(db, K, local_rank, None, 1)
        for _ in range(10):
            s.search_local(d_query.data_ptr(), CUTOFF, stream)
        torch.cuda.synchronize()
        reps = 100
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for _ in range(reps):
            s.search_local(d_query.data_ptr(), CUTOFF, stream)
        b.record(stream); b.synchronize()
        stream_ms = a.elapsed_time(b) / reps
        lat = []
        for _ in range(30):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream); s.search_local(d_query.data_ptr(), CUTOFF, stream); b.record(stream); b.synchronize()
            lat.append(a.elapsed_time(b))
        gbs = rows * ROW_BYTES / (stream_ms * 1e-3) / 1e9
        out.append({"rows": rows, "k": K, "stream_ms": stream_ms, "latency_ms": statistics.median(lat),
                    "stream_gbs": gbs, "stream_frac_of_copy_peak": gbs / peak_gbs,
                    "latency_gbs": rows * ROW_BYTES / (statistics.median(lat) * 1e-3) / 1e9})
        del s
        db.close()
        log("small_shard", out[-1])
    return out


# ------------------------------------------------------------------------------------ multi-query leg
def sm_clock_hz(clocks):
    mhz = (clocks or {}).get("sm_mhz") or (clocks or {}).get("sm_max_mhz") or 1965.0
    return float(mhz) * 1e6


def multi_query_leg(gsb, O, np, torch, db, dist, world, rank, local_rank, rows_total, stream, clocks):
    """BASELINE configs[4]: MQ_QUERIES queries (copies of database rows), top-MQ_K, against the resident
    shards: one pass of the bit-sliced multi-query kernel per rank, all-gather of the per-rank
    records, merge kernel.  Device-timed (CUDA events, max over ranks) and end to end with host
    buffers; plus a dense-query sweep (queries with 32 / 128 / 512 set bits)."""
    from gpusimilarity_b200.dist import ShardedBatchSearcher
    dev = torch.device("cuda", local_rank)
    nq, k = MQ_QUERIES, MQ_K
    seed_rows = (np.arange(nq, dtype=np.uint64) * np.uint64(rows_total // nq + 1)) % np.uint64(rows_total)
    qs = np.ascontiguousarray(O.synth_rows(SEED, seed_rows, 32, PLANT_PERIOD))
    d_q = torch.from_numpy(qs).to(dev)
    s = ShardedBatchSearcher(db, k, local_rank, dist, world)
    group = s.max_queries(nq, 0.0)
    groups = [(q0, min(group, nq - q0)) for q0 in range(0, nq, group)]

    def run_device(dq):
        for q0, n in groups:
            s.search_device(dq[q0:].data_ptr(), n, 0.0, stream)

    def timed(fn, reps=2):
        best = None
        for _ in range(reps):
            if dist is not None:
                dist.barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream); fn(); b.record(stream); b.synchronize()
            ms = a.elapsed_time(b)
            if dist is not None:
                t = torch.tensor([ms], dtype=torch.float64, device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms = float(t.item())
            best = ms if best is None else min(best, ms)
        return best

    launches0 = gsb.launch_count()
    run_device(d_q)
    torch.cuda.synchronize()
    launches = gsb.launch_count() - launches0
    ms = timed(lambda: run_device(d_q))

    # ---- end to end: host queries in (H2D), results back in pinned host memory (D2H)
    def e2e_once():
        if dist is None:
            return db.search_batch_rows_raw(qs, k, 0.0)
        return s.search_host(qs, 0.0, stream)

    e2e_once()
    t_e2e = []
    for _ in range(3):
        if dist is not None:
            dist.barrier()
        t = time.perf_counter(); res = e2e_once(); t_e2e.append(time.perf_counter() - t)
    e2e_ms = min(t_e2e) * 1e3
    if dist is not None:
        t = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
    rows_h, scores_h, n_h, approx_h = res

    # ---- roofline of the counting loop: shared-memory (LDS) wavefronts.  Algorithmic work = ONE
    # 32-bit shared-memory word per lane, i.e. one wavefront per warp, for every set bit of a query
    # and every 1024-row tile; the SM serves 1 wavefront per clock (tools/ubench.cu, profiles/r01_sweep.md).
    set_bits = int(np.bitwise_count(qs.view(np.uint32)).sum())
    tiles_per_gpu = (rows_total / world) / 1024.0
    wavefronts = tiles_per_gpu * set_bits
    peak = 148 * sm_clock_hz(clocks)
    roof = {"bound": "lds", "kernel": "scan_sliced_kernel<32,32>", "unit": "wavefronts/s",
            "achieved": wavefronts / (ms * 1e-3), "peak": peak, "frac": wavefronts / (ms * 1e-3) / peak,
            "peak_source": "1 shared-memory wavefront / clk / SM x 148 SMs x SM clock under load (tools/ubench.cu)",
            "algorithmic_units_per_launch": wavefronts,
            "unit_definition": "1 wavefront per set bit of a query and 1024-row tile (the tile word of 32 lanes); "
                               "list-entry fetches, filter and select are overhead, not algorithmic",
            "query_set_bits_total": set_bits, "traffic": None}

    # ---- dense queries: the bit-sliced kernel pays per set bit of the queries, the tensor-core kernel
    # (tcgen05.mma kind::i8, 128 queries per launch) does not; both are timed on the same batches and
    # must return identical keys.  GSB_BATCH_KERNEL forces the kernel for the device entry point (the
    # host-buffer API picks by the batch's set-bit total).
    dense = []
    rng = np.random.default_rng(7)
    keys_of = {}
    mac_peak = 4.5e15 / 2   # dense i8 MAC/s (B200_PROFILING.md: 4.5 POPS = 2 ops per MAC)
    for bits in (32, 128, 512):
        dq_np = np.zeros((256, 32), dtype=np.uint32)
        for j in range(256):
            pos = rng.choice(1024, size=bits, replace=False)
            np.bitwise_or.at(dq_np[j], pos >> 5, np.uint32(1) << (pos & 31).astype(np.uint32))
        d_dq = torch.from_numpy(dq_np.view(np.int32)).to(dev)
        entry = {"set_bits_per_query": bits, "queries": 256}
        for mode, name in (("3", "sliced"), ("4", "tensor")):
            os.environ["GSB_BATCH_KERNEL"] = mode

            def run_dense():
                s.search_device(d_dq.data_ptr(), 256, 0.0, stream)
            try:
                run_dense()
                torch.cuda.synchronize()
                dms = timed(run_dense, reps=1)
                keys_of[name] = (s.out_rows[:256 * k].clone(), s.out_scores[:256 * k].clone(), s.out_n[:256].clone())
                entry[name + "_ms"] = dms
                entry[name + "_row_query_per_s"] = rows_total * 256 / (dms * 1e-3)
                if name == "sliced":
                    entry["sliced_lds_frac"] = (tiles_per_gpu * 256 * bits) / (dms * 1e-3) / peak
                else:
                    entry["tensor_mac_frac"] = (rows_total / world) * 256 * 1024 / (dms * 1e-3) / mac_peak
            except Exception as e:
                entry[name + "_error"] = repr(e)
        os.environ.pop("GSB_BATCH_KERNEL", None)
        if "sliced" in keys_of and "tensor" in keys_of:
            entry["identical_results"] = all(bool(torch.equal(a, b)) for a, b in zip(keys_of["sliced"], keys_of["tensor"]))
        keys_of.clear()
        dense.append(entry)
        log("dense", entry)
    return {"workload": f"{rows_total} rows, {nq} queries per batch, top-{k}, {world} shard(s)",
            "batch_ms": ms, "queries_per_s": nq / (ms * 1e-3),
            "row_query_per_s": rows_total * nq / (ms * 1e-3), "queries_per_pass": group,
            "gpu_launches_per_batch": int(launches),
            "e2e": {"value": nq / (e2e_ms * 1e-3), "unit": "queries/s", "batch_ms": e2e_ms,
                    "h2d_bytes_per_step": int(qs.nbytes) * world, "d2h_bytes_per_step": nq * k * 8 + nq * 12,
                    "api": "gsb_db_search_batch (host buffers)" if dist is None else
                           "ShardedBatchSearcher.search_host (pinned host buffers, NCCL all-gather, merge kernel)"},
            "roofline": roof, "dense_sweep": dense}, (qs, rows_h, scores_h, n_h, approx_h)


# ------------------------------------------------------------------------------------ verification
def verify_against_streamed_oracle(OC, np, rows_total, single, batch, budget_s, max_queries):
    """Bit-exact check of what was timed: the oracle regenerates the synthetic rows on the fly on the
    host cores and scores them (oracle_stream_search_multi).  `single` = (query, rows, scores, approx),
    `batch` = (queries, rows [nq][k], scores, n, approx) or None.  The number of batch queries checked
    is fitted to `budget_s` from a timed 4 M-row probe and reported."""
    q1, rows1, scores1, approx1 = single
    qs = [q1]
    picks = []
    if batch is not None:
        bq, brows, bscores, bn, bapprox = batch
        probe_n = 4_000_000
        t = time.perf_counter(); OC.c_stream_search_multi(bq[:9], SEED, PLANT_PERIOD, probe_n, K, CUTOFF)
        t9 = time.perf_counter() - t
        t = time.perf_counter(); OC.c_stream_search_multi(bq[:1], SEED, PLANT_PERIOD, probe_n, K, CUTOFF)
        t1 = time.perf_counter() - t
        per_q = max(1e-9, (t9 - t1) / 8) * rows_total / probe_n
        base = t1 * rows_total / probe_n
        fit = int(max(4, min(max_queries, (budget_s - base) / per_q)))
        picks = [int(i) for i in np.linspace(0, bq.shape[0] - 1, fit)]
        qs += [bq[i] for i in picks]
    t = time.perf_counter()
    want = OC.c_stream_search_multi(np.stack(qs), SEED, PLANT_PERIOD, rows_total, K, CUTOFF)
    took = time.perf_counter() - t
    w_rows, w_scores, w_approx = want[0]
    ok_single = (len(rows1) == len(w_rows) and np.array_equal(rows1, w_rows)
                 and np.array_equal(scores1.view(np.uint32), w_scores.view(np.uint32))
                 and (approx1 is None or approx1 == w_approx))
    ok_batch, lists = None, 0
    if batch is not None:
        ok_batch = True
        for j, i in enumerate(picks):
            w_rows, w_scores, w_approx = want[1 + j]
            n = int(bn[i])
            kk = min(MQ_K, len(w_rows))
            ok = (n == kk and np.array_equal(brows[i, :n], w_rows[:kk])
                  and np.array_equal(np.ascontiguousarray(bscores[i, :n]).view(np.uint32), w_scores[:kk].view(np.uint32))
                  and int(bapprox[i]) == w_approx)
            ok_batch = ok_batch and bool(ok)
            lists += 1
    return {"single_query_top1000_bit_exact": bool(ok_single), "multi_query_lists_bit_exact": ok_batch,
            "multi_query_lists_checked": lists, "oracle": "oracle_stream_search_multi, full database, host cores",
            "oracle_seconds": round(took, 1)}


def cross_check_exchange(gsb, np, torch, db, dist, world, rank, local_rank, searcher, d_query, stream):
    """N > 1, outside the timed regions: the fused in-kernel exchange, the NCCL all-gather + merge
    kernel and a host merge of the gathered per-shard records must give the same answer."""
    from gpusimilarity_b200.dist import RECORD_EXTRA, ShardedSearcher, unpack_key
    dev = torch.device("cuda", local_rank)
    searcher.search_device(d_query.data_ptr(), CUTOFF, stream)
    torch.cuda.synchronize()
    n_a = int(searcher.out_n.item())
    a_rows = searcher.out_rows[:n_a].cpu().numpy().astype(np.int64) & 0xFFFFFFFF
    a_scores = searcher.out_scores[:n_a].cpu().numpy()
    a_approx = searcher.approx_count()
    nccl = ShardedSearcher(db, K, local_rank, dist, world, rank=rank, fused=False)
    nccl.search_device(d_query.data_ptr(), CUTOFF, stream)
    torch.cuda.synchronize()
    n_b = int(nccl.out_n.item())
    b_rows = nccl.out_rows[:n_b].cpu().numpy().astype(np.int64) & 0xFFFFFFFF
    b_scores = nccl.out_scores[:n_b].cpu().numpy()
    rec = nccl.gathered.cpu().numpy().view(np.uint64).reshape(world, K + RECORD_EXTRA)
    keys = np.concatenate([rec[r, :int(rec[r, K + 1]) & 0xFFFFFFFF] for r in range(world)])
    keys = np.sort(keys)[::-1][:K]
    c_rows, c_scores = unpack_key(keys)
    c_approx = int(rec[:, K].sum())
    same = (np.array_equal(a_rows, b_rows) and np.array_equal(a_rows, c_rows)
            and np.array_equal(a_scores.view(np.uint32), b_scores.view(np.uint32))
            and np.array_equal(a_scores.view(np.uint32), c_scores.view(np.uint32)) and a_approx == c_approx)
    t = torch.tensor([1 if same else 0], dtype=torch.int32, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    return {"fused": bool(searcher.fused), "fused_eq_nccl_eq_host_merge_on_every_rank": bool(t.item() == 1)}


def single_process_leg(gsb, O, np, torch, world, rows_total, steps):
    """One process driving all `world` GPUs through gsb_db_search (the reference's own mode and what
    the C++ FingerprintDB adapter reaches): a sharded database, one launch per shard and query, host
    merge of the per-shard records."""
    db = gsb.FingerprintDB.synthetic_sharded(rows_total, list(range(world)), seed=SEED, plant_period=PLANT_PERIOD)
    q = O.synth_template(SEED, 32)
    for _ in range(3):
        res = db.search_rows(q, K, CUTOFF)
    t = time.perf_counter()
    for _ in range(steps):
        res = db.search_rows(q, K, CUTOFF)
    serial_ms = (time.perf_counter() - t) / steps * 1e3
    t = time.perf_counter()
    prev = db.search_rows_async(q, K, CUTOFF)
    for _ in range(steps - 1):
        cur = db.search_rows_async(q, K, CUTOFF)
        res = db.search_rows_wait(prev)
        prev = cur
    res = db.search_rows_wait(prev)
    piped_ms = (time.perf_counter() - t) / steps * 1e3
    db.close()
    return {"shards": world, "api": "gsb_db_search / gsb_db_search_async over one sharded gsb_db, host merge",
            "ms_per_step": piped_ms, "value": 1e3 / piped_ms, "unit": "queries/s",
            "serial_ms_per_step": serial_ms, "serial_value": 1e3 / serial_ms}, res


# ------------------------------------------------------------------------------------ GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--rows", type=int, default=1_000_000_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-multi-query", action="store_true")
    ap.add_argument("--no-reference-cuda", action="store_true")
    ap.add_argument("--no-single-process", action="store_true")
    ap.add_argument("--no-small-shard", action="store_true")
    ap.add_argument("--no-full-verify", action="store_true")
    ap.add_argument("--verify-budget-s", type=float, default=75.0)
    ap.add_argument("--verify-queries", type=int, default=32)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return run_reference(args, rank)

    import numpy as np
    import torch
    import gpusimilarity_b200 as gsb
    from gpusimilarity_b200.dist import ShardedSearcher, shard_range
    from oracle import oracle as O, oracle_c as OC   # checker + reported baselines only, never the timed path

    if not torch.cuda.is_available() or gsb.get_gpu_count() == 0:
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- reported baseline, before the big database takes the HBM: the reference's own CUDA path
    reference_cuda = None
    if world == 1 and not args.no_reference_cuda:
        try:
            sizes = [10_000_000, 100_000_000] if args.rows >= 100_000_000 else [min(args.rows, 10_000_000)]
            reference_cuda = reference_cuda_leg(gsb, O, OC, np, sizes)
        except Exception as e:  # reporting only: never hide the headline number
            reference_cuda = [{"error": repr(e)}]

    small_shard = None
    if world == 1 and not args.no_small_shard and args.rows >= 100_000_000:
        try:
            small_shard = small_shard_leg(gsb, O, np, torch, ShardedSearcher, local_rank, [10_000_000, 125_000_000],
                                          measured_peak_gbs()[0])
        except Exception as e:
            small_shard = [{"error": repr(e)}]

    # ---- the database: contiguous, equal shards of the synthetic rows, generated in HBM
    row_base, n_rows = shard_range(args.rows, rank, world)
    t_gen = time.perf_counter()
    db = gsb.FingerprintDB.synthetic(n_rows, device=local_rank, seed=SEED, plant_period=PLANT_PERIOD,
                                     row_base=row_base)
    torch.cuda.synchronize()
    t_gen = time.perf_counter() - t_gen
    info = db.scan_info(K)

    query_np = O.synth_template(SEED, 32)
    d_query = torch.from_numpy(query_np.copy()).to(dev)
    stream = torch.cuda.current_stream()
    fused = os.environ.get("GSB_FUSED_EXCHANGE", "1") != "0"
    searcher = ShardedSearcher(db, K, local_rank, dist, world, rank=rank, fused=fused)

    def device_step():
        """scan (1 launch; sharded: + exchange + merge inside it); query and results stay in HBM."""
        if dist is None:
            searcher.search_local(d_query.data_ptr(), CUTOFF, stream)
        else:
            searcher.search_device(d_query.data_ptr(), CUTOFF, stream)

    # host-buffer API: N = 1 the C ABI's own calls, N > 1 the per-rank searcher over the same C ABI
    def e2e_serial():
        if dist is None:
            return db.search_rows(query_np, K, CUTOFF)
        return searcher.search_host(query_np, CUTOFF, stream)

    def e2e_submit():
        if dist is None:
            return db.search_rows_async(query_np, K, CUTOFF)
        return searcher.submit_host(query_np, CUTOFF, stream)

    def e2e_wait(h):
        if dist is None:
            return db.search_rows_wait(h)
        return searcher.wait_host(h)

    # ---- warm-up
    for _ in range(args.warmup):
        device_step()
        e2e_serial()
    barrier()

    # ---- timed region 1: device-resident steps (the `value`)
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.15)
    launches0 = gsb.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_wall0 = time.time()
    ev0.record(stream)
    for i in range(args.steps):
        device_step()
    ev1.record(stream)
    barrier()
    launches = gsb.launch_count() - launches0
    total_ms = ev0.elapsed_time(ev1)

    # ---- the scan kernel alone (roofline): one event pair per launch, same stream
    kern_ms = []
    for _ in range(args.steps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        searcher.search_local(d_query.data_ptr(), CUTOFF, stream)
        b.record(stream)
        b.synchronize()
        kern_ms.append(a.elapsed_time(b))
    barrier()

    # ---- timed region 2: end to end through host buffers, two queries in flight
    t = time.perf_counter()
    prev = e2e_submit()
    for _ in range(args.steps - 1):
        cur = e2e_submit()
        res = e2e_wait(prev)
        prev = cur
    res = e2e_wait(prev)
    e2e_ms = (time.perf_counter() - t) * 1e3
    barrier()
    # ---- the same, one blocking call after the other
    t = time.perf_counter()
    for _ in range(args.steps):
        res_serial = e2e_serial()
    e2e_serial_ms = (time.perf_counter() - t) * 1e3
    barrier()
    t_wall2 = time.time()
    clocks = sampler.stop(t_wall0, t_wall2)   # all timed regions keep the GPU under the same load

    # max over ranks
    if dist is not None:
        t = torch.tensor([total_ms, e2e_ms, e2e_serial_ms, statistics.mean(kern_ms)], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms, e2e_ms, e2e_serial_ms, kern_mean = (float(x) for x in t.tolist())
    else:
        kern_mean = statistics.mean(kern_ms)

    rows, scores, approx = res
    same_paths = (np.array_equal(rows, res_serial[0])
                  and np.array_equal(scores.view(np.uint32), res_serial[1].view(np.uint32)))

    # ---- outside the timed regions: BASELINE configs[4] on the same resident shards
    multi_query, mq_results = None, None
    if not args.no_multi_query:
        try:
            multi_query, mq_results = multi_query_leg(gsb, O, np, torch, db, dist, world, rank, local_rank, args.rows,
                                                      stream, clocks)
        except Exception as e:  # reporting only: never hide the headline number
            multi_query = {"error": repr(e)}

    exchange_check = None
    if dist is not None:
        try:
            exchange_check = cross_check_exchange(gsb, np, torch, db, dist, world, rank, local_rank, searcher, d_query,
                                                  stream)
        except Exception as e:
            exchange_check = {"error": repr(e)}

    # ---- one process driving all N GPUs (after the per-rank shards are released)
    single_process = None
    if not args.no_single_process:
        if dist is None:
            single_process = {"shards": 1, "same_as": "e2e (N = 1: gsb_db_search on one shard)",
                              "value": args.steps / (e2e_ms * 1e-3), "unit": "queries/s",
                              "serial_value": args.steps / (e2e_serial_ms * 1e-3)}
        else:
            del searcher
            db.close()
            torch.cuda.empty_cache()
            barrier()
            # The other ranks wait on the rendezvous store, NOT in an NCCL barrier: a collective
            # kernel spinning on their GPUs would hold SMs the persistent scan grid needs.
            import datetime
            store = dist.distributed_c10d._get_default_store()
            if rank == 0:
                try:
                    single_process, sp_res = single_process_leg(gsb, O, np, torch, world, args.rows, args.steps)
                    single_process["same_result_as_sharded_run"] = bool(
                        np.array_equal(sp_res[0], rows) and np.array_equal(sp_res[1].view(np.uint32), scores.view(np.uint32)))
                except Exception as e:
                    single_process = {"error": repr(e)}
                store.set("gsb_single_process_done", "1")
            else:
                store.wait(["gsb_single_process_done"], datetime.timedelta(seconds=1500))

    if rank == 0:
        # ---- full verification of what was timed: the oracle streams the same rows on the host
        verification = {"e2e_async_eq_serial": bool(same_paths)}
        ok = bool(same_paths) and len(rows) == min(K, args.rows)
        if not args.no_full_verify:
            try:
                batch = None
                if mq_results is not None:
                    batch = mq_results
                v = verify_against_streamed_oracle(OC, np, args.rows, (query_np, rows, scores, approx), batch,
                                                   args.verify_budget_s, args.verify_queries)
                verification.update(v)
                ok = ok and v["single_query_top1000_bit_exact"] and (v["multi_query_lists_bit_exact"] is not False)
            except Exception as e:
                verification["error"] = repr(e)
                ok = False
        else:
            probe = np.concatenate([rows[:8], rows[-8:]]).astype(np.uint64)
            fps = O.synth_rows(SEED, probe, 32, PLANT_PERIOD)
            want = O.tanimoto_scores_gpu(query_np, fps, CUTOFF)
            ok = ok and bool(np.array_equal(want.view(np.uint32),
                                            np.concatenate([scores[:8], scores[-8:]]).view(np.uint32)))
            verification["probes_only"] = True
        if exchange_check is not None:
            verification["exchange"] = exchange_check
            ok = ok and bool(exchange_check.get("fused_eq_nccl_eq_host_merge_on_every_rank"))

        peak, peak_src = measured_peak_gbs()
        alg_bytes = n_rows * ROW_BYTES                       # per launch, this rank's shard
        achieved = alg_bytes / (kern_mean * 1e-3) / 1e9
        traffic, traffic_src = ncu_traffic_per_launch(n_rows)
        line = {
            "metric": METRIC, "value": args.steps / (total_ms * 1e-3), "unit": "queries/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "u32 popcount + f32 divide", "data": "synthetic",
            "config": {
                "workload": f"{args.rows} x 1024-bit synthetic fingerprints ({args.rows * ROW_BYTES / 1e9:.1f} GB), "
                            f"single query, top-{K}, cutoff {CUTOFF}, {world} shard(s)",
                "rows_per_gpu": n_rows, "k": K, "l2": "database shard is far larger than the 126 MB L2",
                "grid": info.grid, "block": info.block, "tma_stages_per_warp": info.stages, "tma_bytes_per_copy": info.tile_bytes,
                "layout_bytes_per_query": info.db_bytes_per_query,
                "smem_bytes": info.smem_bytes, "synthetic_gen_s": round(t_gen, 2),
                "launch": "programmatic dependent launch: the scan of query i+1 starts while the last CTA of "
                          "query i finishes its select / exchange / merge" if os.environ.get("GSB_PDL", "1") != "0"
                          else "cooperative launch, no overlap between queries",
                "parallelism": f"row-sharded x{world}" + (
                    "" if world == 1 else (", fused in-kernel exchange of per-shard top-k over NVLink peer memory"
                                           if fused else ", NCCL all-gather of per-shard top-k + merge kernel")),
            },
            "e2e": {"value": args.steps / (e2e_ms * 1e-3), "unit": "queries/s",
                    "h2d_bytes_per_step": 128 * world, "d2h_bytes_per_step": (K * 8 + 24) * world,
                    "ms_per_step": e2e_ms / args.steps,
                    "mode": "host-buffer API with two queries in flight (gsb_db_search_async / _wait); query as "
                            "launch parameter, results stored by the kernel into mapped pinned host memory",
                    "serial": {"value": args.steps / (e2e_serial_ms * 1e-3), "ms_per_step": e2e_serial_ms / args.steps,
                               "mode": "one blocking gsb_db_search after the other"}},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": "scan_topk_kernel<W=32>", "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": kern_mean,
                         "traffic": traffic, "traffic_source": traffic_src},
            "verified": bool(ok), "verification": verification,
        }
        if multi_query is not None:
            line["multi_query"] = multi_query
        if single_process is not None:
            line["single_process"] = single_process
        if small_shard is not None:
            line["small_shard"] = small_shard
        if reference_cuda is not None:
            line["reference_cuda"] = reference_cuda
            ing = [r["ingest"] for r in reference_cuda if isinstance(r, dict) and "ingest" in r]
            if ing:
                line["ingest"] = ing[-1]
        if world == 1 and not args.no_cpu_baseline:
            try:
                res_cpu = cpu_reference_run(3, 1, args.rows)
                line["cpu_baseline"] = {k: res_cpu[k] for k in ("value", "unit", "cores", "kind", "sample", "scan_only")}
            except Exception as e:  # the baseline is reporting only; never hide the GPU number
                line["cpu_baseline"] = {"value": None, "unit": "queries/s", "cores": os.cpu_count(),
                                        "kind": "unavailable", "sample": repr(e)}
        print(json.dumps(line))
    if dist is not None:
        import datetime
        store = dist.distributed_c10d._get_default_store()
        if rank == 0:
            store.set("gsb_bench_done", "1")
        else:   # rank 0 may spend a minute in the host-side verification: wait on the store, not in NCCL
            store.wait(["gsb_bench_done"], datetime.timedelta(seconds=1500))
        dist.barrier()
        dist.destroy_process_group()


def _main_with_clean_stdout():
    """stdout carries exactly one JSON line: libraries that print to fd 1 (NCCL's version banner,
    for one) are sent to stderr for the duration of the run."""
    import io
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    captured = io.StringIO()
    real_stdout, sys.stdout = sys.stdout, captured
    try:
        main()
    finally:
        sys.stdout = real_stdout
        sys.stdout.flush()
        os.dup2(saved, 1)
        os.close(saved)
    lines = [l for l in captured.getvalue().splitlines() if l.strip()]
    json_lines = [l for l in lines if l.lstrip().startswith("{")]
    for l in lines:
        if l not in json_lines:
            print(l, file=sys.stderr)
    if json_lines:
        print(json_lines[-1], flush=True)


if __name__ == "__main__":
    _main_with_clean_stdout()
