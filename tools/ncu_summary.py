"""Summarise an .ncu-rep (one --set full capture) into markdown for profiles/."""
import csv
import subprocess
import sys

rep, out, title = sys.argv[1], sys.argv[2], sys.argv[3]
# a .csv argument is the output of `ncu -i X.ncu-rep --page raw --csv` made on the GPU box
raw = open(rep).read() if rep.endswith(".csv") else \
    subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
lines = raw.splitlines()
start = [i for i, l in enumerate(lines) if l.startswith('"ID"')][0]
rows = list(csv.reader(lines[start:]))
hdr, units, vals = rows[0], rows[1], rows[2]
d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
want = [
    "Kernel Name", "Block Size", "Grid Size", "gpu__time_duration.sum", "sm__cycles_elapsed.max",
    "launch__registers_per_thread", "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.per_cycle_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
]
with open(out, "w") as f:
    f.write(f"# {title}\n\n`ncu --set full --clock-control none --import-source on` (one launch; ~40 replay passes).\n\n")
    f.write("| metric | value | unit |\n|---|---|---|\n")
    for k in want:
        if k in d:
            f.write(f"| {k} | {d[k][0]} | {d[k][1]} |\n")
    f.write("\n## Warp stall reasons (warps per issue-active cycle)\n\n| reason | value |\n|---|---|\n")
    st = [(k, float(v[0])) for k, v in d.items() if "average_warps_issue_stalled" in k and k.endswith("per_issue_active.ratio")]
    for k, v in sorted(st, key=lambda x: -x[1]):
        f.write(f"| {k.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')} | {v:.3f} |\n")
print(open(out).read())
