"""Summarise `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass` by CUDA source line:
warp instructions per warp and column, share of the warp samples, top stall reasons.
    python tools/ncu_lines_summary.py pair_source.csv WARPS COLUMNS 'title' > profiles/..._lines.md"""
import csv
import sys
from collections import defaultdict


def main(path, warps, columns, title):
    rows = list(csv.reader(open(path, newline="")))
    hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
    hdr = rows[hdr_i]
    col = {name: i for i, name in enumerate(hdr)}
    main_file = next(r[1] for r in rows if r and r[0] == "File Path").rsplit("/", 1)[-1]
    src_i = 1                                     # first "Source" column = the CUDA source text
    stall_cols = [(n, i) for n, i in col.items() if n.startswith("stall_") and "Not Issued" not in n]
    per = defaultdict(lambda: {"inst": 0.0, "samples": 0.0, "src": "", "stalls": defaultdict(float)})
    cur_line, cur_src, cur_file = None, "", ""
    for r in rows:
        if r and r[0] == "File Path":
            cur_file, cur_line = r[1].rsplit("/", 1)[-1], None
            continue
        if len(r) < len(hdr) or r[0] == "Line No":
            continue
        if not r[0].strip():
            continue                              # SASS rows under a source line: the line's row already has the totals
        cur_line, cur_src = int(r[0]), r[src_i].strip()
        d = per[cur_line if cur_file == main_file else f"{cur_file}:{cur_line}"]
        d["src"] = cur_src
        try:
            d["inst"] += float(r[col["Instructions Executed"]] or 0)
            d["samples"] += float(r[col["# Samples"]] or 0)
        except ValueError:
            continue
        for n, i in stall_cols:
            try:
                d["stalls"][n[6:]] += float(r[i] or 0)
            except ValueError:
                pass
    tot_s = sum(d["samples"] for d in per.values()) or 1.0
    tot_i = sum(d["inst"] for d in per.values())
    print(f"# {title}\n")
    print(f"Per CUDA source line (inlined helpers are attributed to the helper's line). `inst/warp-col` = warp "
          f"instructions of the line / ({warps} warps x {columns} columns); samples = share of all warp samples. "
          f"An instruction of an inlined helper counts on the helper's line and on the line that called it. Lines below 0.7 % of the "
          f"samples and below 8 inst/warp-col are omitted.\n")
    print("| line | inst/warp-col | samples | top stalls | source |")
    print("|---|---|---|---|---|")
    for line, d in sorted(per.items(), key=lambda kv: -kv[1]["samples"]):
        ipc = d["inst"] / warps / columns
        share = d["samples"] / tot_s
        if share < 0.007 and ipc < 8:
            continue
        st = sorted(d["stalls"].items(), key=lambda kv: -kv[1])[:3]
        ssum = sum(d["stalls"].values()) or 1.0
        sts = ", ".join(f"{n} {100 * v / ssum:.0f}%" for n, v in st if v > 0)
        print(f"| {line} | {ipc:.1f} | {100 * share:.1f}% | {sts} | `{d['src'][:90]}` |")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), sys.argv[4])
