#!/usr/bin/env python
"""Markdown summary of one `ncu --set full` capture for profiles/.

    python tools/ncu_summary.py <name>.raw.csv <name>.source.csv "title" "command" > profiles/<name>.md

The inputs are `ncu -i <rep> --page raw --csv` and `--page source --csv` of the same report
(tools/gpu_r2f.sh exports them on the GPU box; the .ncu-rep itself is too big to travel back).
"""
import csv
import sys

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
    "smsp__warps_eligible.avg.per_cycle_active", "sm__cycles_active.min", "sm__cycles_active.avg",
    "sm__cycles_active.max", "sm__cycles_elapsed.max", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.per_second",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_subpipe_imma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum", "smsp__inst_executed_pipe_tmem.sum",
]
STALL = "smsp__average_warps_issue_stalled_"


def main():
    raw, src, title, cmd = sys.argv[1:5]
    rows = list(csv.reader(open(raw)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    col = {h: i for i, h in enumerate(hdr)}
    print(f"# {title}\n")
    print(f"Kernel: `{vals[col['Kernel Name']]}`, grid {vals[col['Grid Size']]}, block {vals[col['Block Size']]}.")
    print(f"Command: `{cmd}`")
    print("Per-launch numbers under the profiler are cold-cache and serialised; the reported timings are bench.py's "
          "CUDA-event ones.\n")
    print("| metric | value | unit |\n|---|---|---|")
    for m in METRICS:
        if m in col and vals[col[m]] != "":
            print(f"| {m} | {vals[col[m]]} | {units[col[m]]} |")
    stalls = []
    for h, i in col.items():
        if h.startswith(STALL) and h.endswith("_per_warp_active.pct") is False and h.endswith(".ratio") and "not_issued" not in h:
            try:
                stalls.append((float(vals[i]), h[len(STALL):].replace("_per_issue_active.ratio", "")))
            except ValueError:
                pass
    if stalls:
        print("\nWarp stall reasons (warps per issue-active cycle):\n")
        for v, n in sorted(stalls, reverse=True)[:9]:
            print(f"- {n}: {v:.2f}")
    srows = list(csv.reader(open(src)))
    h_i = next(i for i, r in enumerate(srows) if r and r[0] == "Address")
    sh = {h: i for i, h in enumerate(srows[h_i])}
    body = [r for r in srows[h_i + 1:] if len(r) > sh["# Samples"]]
    total = sum(int(r[sh["# Samples"]] or 0) for r in body) or 1
    top = sorted(body, key=lambda r: -int(r[sh["# Samples"]] or 0))[:10]
    print("\nTop sampled SASS instructions (source page):\n")
    for r in top:
        print(f"- {100.0 * int(r[sh['# Samples']] or 0) / total:.1f}% `{r[sh['Source']].strip()}`")
    mn = {}
    for r in body:
        op = r[sh["Source"]].strip().split()
        if not op:
            continue
        name = op[1] if op[0].startswith("@") and len(op) > 1 else op[0]
        mn[name.split(".")[0]] = mn.get(name.split(".")[0], 0) + int(r[sh["Instructions Executed"]] or 0)
    tot = sum(mn.values()) or 1
    print("\nExecuted warp-instructions by SASS mnemonic (source page):\n")
    for name, n in sorted(mn.items(), key=lambda kv: -kv[1])[:14]:
        print(f"- {name}: {n} ({100.0 * n / tot:.1f}%)")


if __name__ == "__main__":
    main()
