#!/usr/bin/env python
"""Per-source-line view of an `ncu --page source --csv` export: joins the SASS rows (by offset from
the kernel's first instruction) with the line table of the same kernel from
`nvdisasm -g -c <cubin>` (cuobjdump -xelf all libgpusim_b200.so gives the cubin).

    python tools/ncu_lines.py <name>.source.csv <nvdisasm listing> <mangled kernel name> [top]
"""
import csv, re, sys
from collections import defaultdict


def line_table(listing, kernel):
    table, cur, inside = {}, None, False
    for ln in open(listing, errors="replace"):
        if ln.startswith("//--------------------- .text."):
            inside = kernel in ln
            continue
        if not inside:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            table[int(m.group(1), 16)] = (cur, m.group(2).strip())
    return table


def main():
    src, listing, kernel = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    table = line_table(listing, kernel)
    rows = list(csv.reader(open(src)))
    h_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    col = {h: i for i, h in enumerate(rows[h_i])}
    body = [r for r in rows[h_i + 1:] if len(r) > col["# Samples"]]
    base = int(body[0][col["Address"]], 16)
    per = defaultdict(lambda: [0, 0, 0])
    tot_s = tot_i = 0
    for r in body:
        off = int(r[col["Address"]], 16) - base
        key = table.get(off, (None, ""))[0] or ("?", 0)
        s, n = int(r[col["# Samples"]] or 0), int(r[col["Instructions Executed"]] or 0)
        per[key][0] += s
        per[key][1] += n
        tot_s += s
        tot_i += n
    print(f"{tot_s} samples, {tot_i} warp-instructions executed")
    print("samples%  instr%   file:line")
    for key, (s, n, _) in sorted(per.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"{100.0 * s / max(tot_s, 1):7.2f} {100.0 * n / max(tot_i, 1):7.2f}   {key[0]}:{key[1]}")


if __name__ == "__main__":
    main()
