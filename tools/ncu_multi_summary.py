"""Summarise the raw page of an `ncu --set full` capture with SEVERAL launches (ncu -i x.ncu-rep --page raw --csv)
into one markdown table, a column per launch: python tools/ncu_multi_summary.py raw.csv out.md "title" """
import csv
import re
import sys

src, out, title = sys.argv[1], sys.argv[2], sys.argv[3]
lines = open(src).read().splitlines()
start = [i for i, l in enumerate(lines) if l.startswith('"ID"')][0]
rows = list(csv.reader(lines[start:]))
hdr, units, launches = rows[0], rows[1], rows[2:]
want = [
    "Block Size", "Grid Size", "gpu__time_duration.sum", "launch__registers_per_thread",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__sass_inst_executed_op_local_ld.sum",
    "smsp__sass_inst_executed_op_local_st.sum",
]


def short(name):
    m = re.search(r"extend_kernel<([^>]*)>", name)
    return "extend_kernel<" + re.sub(r"\((?:int|bool)\)", "", m.group(1)) + ">" if m else name.split("(")[0]


def fmt(v):
    try:
        x = float(v.replace(",", ""))
        return f"{x:.4g}"
    except ValueError:
        return v


idx = {h: i for i, h in enumerate(hdr)}
names = [short(r[idx["Kernel Name"]]) for r in launches]
with open(out, "w") as f:
    f.write(f"# {title}\n\n`ncu --set full --clock-control none` of every `extend_kernel` launch (≈40 replay passes each; times are "
            "serialised and cold-cache).\n\n")
    f.write("| metric | " + " | ".join(f"`{n}`" for n in names) + " | unit |\n|---|" + "---|" * (len(names) + 1) + "\n")
    for k in want:
        if k in idx:
            f.write(f"| {k} | " + " | ".join(fmt(r[idx[k]]) for r in launches) + f" | {units[idx[k]]} |\n")
    f.write("\n## Warp stall reasons (warps per issue-active cycle)\n\n| reason | " + " | ".join(f"`{n}`" for n in names) + " |\n|---|" +
            "---|" * len(names) + "\n")
    st = [h for h in hdr if "average_warps_issue_stalled" in h and h.endswith("per_issue_active.ratio")]
    st.sort(key=lambda h: -max(float(r[idx[h]] or 0) for r in launches))
    for h in st[:10]:
        nm = h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")
        f.write(f"| {nm} | " + " | ".join(f"{float(r[idx[h]] or 0):.3f}" for r in launches) + " |\n")
print(open(out).read())
