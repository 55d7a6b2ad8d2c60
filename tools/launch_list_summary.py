"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into markdown for profiles/:
python tools/launch_list_summary.py launches.csv out.md "title" """
import collections
import csv
import re
import sys

src, out, title = sys.argv[1], sys.argv[2], sys.argv[3]
lines = open(src).read().splitlines()
start = [i for i, l in enumerate(lines) if l.startswith('"ID"')][0]
rows = list(csv.reader(lines[start:]))
hdr = rows[0]
iname, iblock, igrid, ival, imetric = (hdr.index(x) for x in ("Kernel Name", "Block Size", "Grid Size", "Metric Value", "Metric Name"))
agg = collections.OrderedDict()
for r in rows[1:]:
    if len(r) <= ival or r[imetric] != "gpu__time_duration.sum":
        continue
    name = r[iname]
    m = re.search(r"(extend_kernel<[^>]*>|\w+_kernel\b|\w+)", name.replace("ramcu::", "").replace("void ", ""))
    short = m.group(1) if m else name
    if "at::native" in name or "elementwise" in name or "vectorized" in name:
        short = "torch fill / copy kernels (L2 flush, bench plumbing)"
    a = agg.setdefault(short, dict(n=0, ns=0.0, shapes=collections.Counter()))
    a["n"] += 1
    a["ns"] += float(r[ival].replace(",", ""))
    a["shapes"][f"{r[igrid]} x {r[iblock]}"] += 1
total = sum(a["ns"] for a in agg.values())
with open(out, "w") as f:
    f.write(f"# {title}\n\n`ncu --metrics gpu__time_duration.sum --clock-control none` (per-launch times under ncu are serialised and "
            "cold-cache, so the SHARE of each kernel is what is comparable with the CUDA-event timing of bench.py).\n\n")
    f.write("| kernel | launches | total ms | share | avg ms | grid x block |\n|---|---|---|---|---|---|\n")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["ns"]):
        shapes = "; ".join(s for s, _ in a["shapes"].most_common(3)) + (" ..." if len(a["shapes"]) > 3 else "")
        f.write(f"| `{k}` | {a['n']} | {a['ns'] / 1e6:.3f} | {100 * a['ns'] / total:.1f} % | {a['ns'] / a['n'] / 1e6:.3f} | {shapes} |\n")
    f.write(f"\nTotal {total / 1e6:.1f} ms over {sum(a['n'] for a in agg.values())} launches.\n")
print(open(out).read())
