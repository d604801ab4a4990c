#!/usr/bin/env python
"""Benchmark of the hot path: single-query brute-force Tanimoto scan + top-1000 over a synthetic
1 B x 1024-bit fingerprint database (BASELINE.json configs[2]; sharded over the ranks for N > 1).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--rows R]

A step is one query over the whole database.  N > 1 is launched by torchrun (one rank per GPU):
every rank scans its contiguous shard, the per-shard candidates are all-gathered over NCCL and
merged by gsb_merge_device.  Prints ONE JSON line (rank 0).

  value      queries/s, device-timed (CUDA events, barrier + synchronize on both sides, max over
             ranks), query and database already resident in HBM
  e2e        the same through the public host-buffer API: pinned host query in, rows/scores out
  roofline   the scan kernel's algorithmic bytes (128 B per row) / its mean launch duration,
             against MEASURED_PEAKS.json's measured copy bandwidth
  cpu_baseline  the reference's own search_cpu (oracle/_ref, its sources compiled verbatim) on a
             bounded sample of the workload on this box's host cores
  multi_query   extra, outside the timed regions above: BASELINE configs[4] (a batch of 1024 queries,
             top-100) on the same resident shards, device-timed (--no-multi-query skips it)

--impl reference times that reference CPU path alone (rank 0; other ranks exit).
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


def measured_peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, copy)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic_per_launch(rows):
    """dram bytes per launch from the committed ncu capture, if it was taken at this size."""
    try:
        with open(os.path.join(ROOT, "profiles", "scan_kernel_ncu.json")) as fh:
            prof = json.load(fh)
        if int(prof.get("rows", -1)) == int(rows):
            return float(prof["dram_bytes_read"]) + float(prof["dram_bytes_write"])
    except Exception:
        pass
    return None


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
    t = time.perf_counter(); score_only(); t_score = time.perf_counter() - t
    per_query = statistics.mean(times) * rows_total / CPU_SAMPLE_ROWS
    return {
        "value": 1.0 / per_query, "unit": "queries/s", "cores": cores, "kind": kind,
        "sample": (f"search_cpu on {CPU_SAMPLE_ROWS} synthetic rows, k={K}: {statistics.mean(times) * 1e3:.1f} ms "
                   f"(score-only {CPU_SAMPLE_ROWS / t_score / 1e6:.1f} M rows/s = "
                   f"{CPU_SAMPLE_ROWS * ROW_BYTES / t_score / 1e9:.2f} GB/s); scaled linearly to "
                   f"{rows_total} rows"),
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
        "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": res["value"], "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------ GPU arm
def multi_query_pass(gsb, O, np, torch, db, dist, world, rank, local_rank, rows_total, stream, n_queries=1024, k=100):
    """BASELINE configs[4]: n_queries queries (copies of database rows), top-k, against the resident
    shards: one pass of the bit-sliced multi-query kernel per rank, all-gather of the per-rank
    records, merge kernel.  Device-timed (CUDA events, max over ranks), second of two runs."""
    from gpusimilarity_b200.dist import ShardedBatchSearcher
    dev = torch.device("cuda", local_rank)
    seed_rows = (np.arange(n_queries, dtype=np.uint64) * np.uint64(rows_total // n_queries + 1)) % np.uint64(rows_total)
    qs = O.synth_rows(SEED, seed_rows, 32, PLANT_PERIOD)
    d_q = torch.from_numpy(np.ascontiguousarray(qs)).to(dev)
    s = ShardedBatchSearcher(db, k, local_rank, dist, world)
    group = s.max_queries(n_queries, 0.0)
    groups = [(q0, min(group, n_queries - q0)) for q0 in range(0, n_queries, group)]
    launches0 = gsb.launch_count()

    def run_all():
        for q0, nq in groups:
            s.search_device(d_q[q0:].data_ptr(), nq, 0.0, stream)

    run_all()
    torch.cuda.synchronize()
    launches = gsb.launch_count() - launches0
    if dist is not None:
        dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    run_all()
    b.record(stream)
    b.synchronize()
    ms = a.elapsed_time(b)
    if dist is not None:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    q0, nq = groups[-1]   # the result buffers hold the last group: every query finds its own row first
    top1 = s.out_rows.view(-1, k)[:nq, 0].cpu().numpy().astype(np.int64) & 0xffffffff
    ok = bool((top1 == seed_rows[q0:q0 + nq].astype(np.int64)).all()) and \
        bool((s.out_scores.view(-1, k)[:nq, 0] == 1.0).all().item())
    return {"workload": f"{rows_total} rows, {n_queries} queries per batch, top-{k}, {world} shard(s)",
            "batch_ms": ms, "queries_per_s": n_queries / (ms * 1e-3),
            "row_query_per_s": rows_total * n_queries / (ms * 1e-3), "queries_per_pass": group,
            "gpu_launches_per_batch": int(launches), "verified": ok}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--rows", type=int, default=1_000_000_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-multi-query", action="store_true")
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
    from oracle import oracle as O   # checker only: verifies the timed results afterwards

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

    # ---- the database: contiguous, equal shards of the synthetic rows, generated in HBM
    row_base, n_rows = shard_range(args.rows, rank, world)
    t_gen = time.perf_counter()
    db = gsb.FingerprintDB.synthetic(n_rows, device=local_rank, seed=SEED, plant_period=PLANT_PERIOD,
                                     row_base=row_base)
    torch.cuda.synchronize()
    t_gen = time.perf_counter() - t_gen
    info = db.scan_info(K)

    query_np = O.synth_template(SEED, 32)
    q_pinned = torch.from_numpy(query_np.copy()).pin_memory()
    d_query = q_pinned.to(dev)
    stream = torch.cuda.current_stream()
    fused = os.environ.get("GSB_FUSED_EXCHANGE", "1") != "0"
    searcher = ShardedSearcher(db, K, local_rank, dist, world, rank=rank, fused=fused)

    def device_step(q_ptr):
        """scan (1 launch) [+ all-gather + merge (1 launch) when sharded]; results stay in HBM."""
        if dist is None:
            searcher.search_local(q_ptr, CUTOFF, stream)
        else:
            searcher.search_device(q_ptr, CUTOFF, stream)

    def e2e_step():
        """Public host-buffer call: pinned query in, rows + scores back on the host."""
        if dist is None:
            return db.search_rows(query_np, K, CUTOFF)
        rows, scores = searcher.search_host(d_query, q_pinned, CUTOFF, stream)
        return rows, scores, None

    # ---- warm-up
    for _ in range(args.warmup):
        device_step(d_query.data_ptr())
        e2e_step()
    barrier()

    # ---- timed region 1: device-resident steps (the `value`)
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.15)
    launches0 = gsb.launch_count()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    barrier()
    t_wall0 = time.time()
    ev[0].record(stream)
    for i in range(args.steps):
        device_step(d_query.data_ptr())
        ev[i + 1].record(stream)
    barrier()
    t_wall1 = time.time()
    launches = gsb.launch_count() - launches0
    total_ms = ev[0].elapsed_time(ev[-1])

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

    # ---- timed region 2: end to end through host buffers
    t_e2e = []
    barrier()
    for _ in range(args.steps):
        t = time.perf_counter()
        res = e2e_step()
        t_e2e.append(time.perf_counter() - t)
    barrier()
    t_wall2 = time.time()
    e2e_ms = sum(t_e2e) * 1e3
    clocks = sampler.stop(t_wall0, t_wall2)   # all three timed regions keep the GPU under the same load

    # max over ranks
    if dist is not None:
        t = torch.tensor([total_ms, e2e_ms, statistics.mean(kern_ms)], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms, e2e_ms, kern_mean = (float(x) for x in t.tolist())
    else:
        kern_mean = statistics.mean(kern_ms)

    # ---- light verification of what was timed (not timed): returned scores re-score exactly
    rows, scores, _ = res
    ok = len(rows) == min(K, args.rows) and bool(np.all(np.diff(scores) <= 0)) and len(set(rows.tolist())) == len(rows)
    if rank == 0 and ok:
        probe = np.concatenate([rows[:8], rows[-8:]]).astype(np.uint64)
        fps = O.synth_rows(SEED, probe, 32, PLANT_PERIOD)
        want = O.tanimoto_scores_gpu(query_np, fps, CUTOFF)
        ok = bool(np.array_equal(want.view(np.uint32), np.concatenate([scores[:8], scores[-8:]]).view(np.uint32)))

    # ---- extra, outside every timed region above: BASELINE configs[4] on the same resident shards
    # (a batch of 1024 queries, top-100, one pass over the database; not part of `value`)
    multi_query = None
    if not args.no_multi_query:
        try:
            multi_query = multi_query_pass(gsb, O, np, torch, db, dist, world, rank, local_rank, args.rows, stream)
        except Exception as e:  # reporting only: never hide the headline number
            multi_query = {"error": repr(e)}

    if rank == 0:
        peak, peak_src = measured_peak_gbs()
        alg_bytes = n_rows * ROW_BYTES                       # per launch, this rank's shard
        achieved = alg_bytes / (kern_mean * 1e-3) / 1e9
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
                "parallelism": f"row-sharded x{world}" + (
                    "" if world == 1 else (", fused in-kernel exchange of per-shard top-k over NVLink peer memory"
                                           if searcher.fused else ", NCCL all-gather of per-shard top-k + merge kernel")),
            },
            "e2e": {"value": args.steps / (e2e_ms * 1e-3), "unit": "queries/s",
                    "h2d_bytes_per_step": 128 * world, "d2h_bytes_per_step": (K + 2) * 8 if world == 1 else K * 8 + 4,
                    "ms_per_step": e2e_ms / args.steps},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": "scan_topk_kernel<W=32>", "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": kern_mean,
                         "traffic": ncu_traffic_per_launch(n_rows)},
            "verified": ok,
        }
        if multi_query is not None:
            line["multi_query"] = multi_query
        if world == 1 and not args.no_cpu_baseline:
            try:
                res_cpu = cpu_reference_run(3, 1, args.rows)
                line["cpu_baseline"] = {k: res_cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}
            except Exception as e:  # the baseline is reporting only; never hide the GPU number
                line["cpu_baseline"] = {"value": None, "unit": "queries/s", "cores": os.cpu_count(),
                                        "kind": "unavailable", "sample": repr(e)}
        print(json.dumps(line))
    if dist is not None:
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
