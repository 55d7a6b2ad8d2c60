#!/usr/bin/env python
"""tools/sass_report.py -- `ptxas -v` and SASS opcode evidence of the shipped band tilings (CPU only).

Compiles every kernel translation unit of csrc/ with `-Xptxas -v` (objects go to a scratch directory; the
product objects are left alone), disassembles them with `cuobjdump -sass` and writes
profiles/r2_ptxas_sass.md:
  * registers / spill bytes / shared memory of every extend_kernel instantiation (ptxas -v),
  * the opcode histogram of the instantiations named on the command line (default: the config-4 tiling, a two-pass tiling and the
    fewest-lane tilings a crowded genome-scale batch runs), whole kernel and the hottest loop bodies are not separated: the histogram is static.
"""
from __future__ import annotations

import collections
import os
import re
import subprocess
import sys
import tempfile
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "repeatafterme_b200", "csrc")
sys.path.insert(0, ROOT)
from repeatafterme_b200.build import ARCH, KERNEL_UNITS, NVCC  # noqa: E402

HIST = ["81, 1, 81, 256, 1, false", "27, 3, 81, 256, 2, false", "29, 1, 29, 384, 1, false", "41, 1, 41, 384, 1, false",
        "41, 3, 121, 384, 1, false", "41, 4, 161, 384, 1, false", "41, 5, 201, 384, 1, false"]


def compile_unit(unit, tmp):
    obj = os.path.join(tmp, unit + ".o")
    cmd = [NVCC, *ARCH, "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC", "-Xptxas", "-v", "-c",
           os.path.join(CSRC, unit + ".cu"), "-o", obj]
    log = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True, check=True)
    return obj, log.stderr


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.split("\n")
    return dict(zip(names, out))


def ptxas_rows(text):
    rows, cur = [], None
    for line in text.splitlines():
        m = re.search(r"Compiling entry function '(\S+)'", line)
        if m:
            cur = {"name": m.group(1), "stack": 0, "st": 0, "ld": 0}
            rows.append(cur)
            continue
        if cur is None:
            continue
        m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
        if m:
            cur["stack"], cur["st"], cur["ld"] = map(int, m.groups())
        m = re.search(r"Used (\d+) registers, used (\d+) barriers(?:, (\d+) bytes cumulative stack size)?, (\d+) bytes smem", line)
        if m:
            cur["regs"], cur["bar"], cur["smem"] = int(m.group(1)), int(m.group(2)), int(m.group(4))
    return rows


def sass_hist(obj):
    txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True, check=True).stdout
    hists, cur = {}, None
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = collections.Counter()
            hists[m.group(1)] = cur
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)((?:\.[A-Z0-9_]+)*)", line)
        if m and cur is not None:
            op = m.group(1)
            mods = m.group(2)
            if op in ("LDL", "STL", "LDS", "STS", "LDG", "STG", "ATOMS", "ATOMG", "RED", "SHFL", "BAR", "VIADDMNMX", "VIMNMX3",
                      "VIMNMX", "REDUX", "IMAD", "IADD3", "PRMT", "LOP3", "SHF", "ISETP", "SEL", "BRA", "MOV", "IADD", "IMNMX"):
                pass
            cur[op] += 1
            cur["_total"] += 1
            if op in ("LDS", "LDG", "LDL", "STL") and mods:
                w = [x for x in mods.split(".") if x in ("128", "64", "U8", "U16")]
                if w:
                    cur[op + "." + w[0]] += 1
    return hists


def main():
    out_md = os.path.join(ROOT, "profiles", "r2_ptxas_sass.md")
    with tempfile.TemporaryDirectory() as tmp:
        with ThreadPoolExecutor(max_workers=os.cpu_count() or 1) as ex:
            res = list(ex.map(lambda u: compile_unit(u, tmp), KERNEL_UNITS))
        rows, hists = [], {}
        for (obj, err), unit in zip(res, KERNEL_UNITS):
            for r in ptxas_rows(err):
                r["unit"] = unit
                rows.append(r)
            hists.update(sass_hist(obj))
    dm = demangle([r["name"] for r in rows] + list(hists))
    lines = ["# Round 2: `ptxas -v` and SASS opcode histogram of the shipped `extend_kernel` tilings", "",
             "Made by `python tools/sass_report.py` (nvcc 12.9 cross-compile, `-gencode arch=compute_100a,code=sm_100a -O3 "
             "-lineinfo -Xptxas -v`, then `cuobjdump -sass`). Template arguments: `<C cells per lane, G lanes per instance, "
             "W static band width (0 = run time), MAXT largest CTA (caps registers at 65536/MAXT), NP passes per thread>`.", "",
             "## ptxas -v", "",
             "| unit | kernel | registers | stack B | spill st B | spill ld B | smem B |", "|---|---|---|---|---|---|---|"]
    for r in rows:
        name = dm[r["name"]]
        m = re.search(r"extend_kernel<([^>]*)>", name)
        short = f"extend_kernel<{m.group(1)}>" if m else name.split("(")[0]
        lines.append(f"| {r['unit']} | `{short}` | {r.get('regs', '?')} | {r['stack']} | {r['st']} | {r['ld']} | {r.get('smem', '?')} |")
    lines += ["", "## SASS opcode histogram (static instruction counts of the whole kernel)", ""]
    want = sys.argv[1:] or HIST
    cols = []
    for w in want:
        for mangled, h in hists.items():
            if f"extend_kernel<{w}>" in dm.get(mangled, ""):
                cols.append((w, h))
                break
    ops = collections.Counter()
    for _, h in cols:
        ops.update(h)
    order = [o for o, _ in ops.most_common() if o != "_total"]
    lines.append("| opcode | " + " | ".join(f"`<{w}>`" for w, _ in cols) + " |")
    lines.append("|---|" + "---|" * len(cols))
    lines.append("| all instructions | " + " | ".join(str(h["_total"]) for _, h in cols) + " |")
    for o in order:
        if max(h[o] for _, h in cols) < 8 and o not in ("LDL", "STL", "BAR", "RED", "ATOMG", "ATOMS", "REDUX"):
            continue
        lines.append(f"| {o} | " + " | ".join(str(h[o]) for _, h in cols) + " |")
    lines += ["", "`VIADDMNMX` = fused add + max (DPX), `VIMNMX3` = three-input max, `VIMNMX` = two-input max, `PRMT` = byte "
              "extraction from the base window, `LDS.128` = the four candidate scores of a cell's base, `SHFL`/`REDUX` = the "
              "lane-group scans and warp sums, `ATOMS`/`ATOMG`/`RED` = the column handshake, `LDL`/`STL` = spills. No tensor-core "
              "or TMA opcodes: the path is integer max-plus over registers (DESIGN.md 3.1).", ""]
    with open(out_md, "w") as f:
        f.write("\n".join(lines))
    print(out_md)


if __name__ == "__main__":
    main()
