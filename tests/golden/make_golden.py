"""Generate the committed golden fixtures from the REFERENCE itself (run in the build container,
where /root/reference and oracle/_ref exist):  python tests/golden/make_golden.py

  unit_vectors.json     data of the reference's own unit test test_compute_nw_row
                        (bnw_extend.c:1384-1651): libraries, cores, forced consensus, the
                        known-answer row 6 (bnw_extend.c:1566-1567) and every row the reference
                        object code produces for it
  flat_*.npz            flat inputs (post-loader sequence bytes + cores + parameters) and the
                        outputs of the reference's extend_alignment() object code, both directions
  cli_*/                inputs (sequences + BED rows) and the byte-exact outputs of the reference
                        RAMExtend binary (-cons / -outtsv / -outfa / filtered stdout)
  ce10_fam{1,2,3}.json  rows of test/ce10-fam*.stk turned into BED-6 by the rule of
                        util/extend-stk.pl:312-359, with the ungapped sequence of every row
"""
import json
import os
import re
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from repeatafterme_b200 import synth  # noqa: E402
from repeatafterme_b200.abi import CORE_DTYPE, make_params  # noqa: E402
from tests import support  # noqa: E402
from tests.cases import (FLAT_CASES, CLI_CASES, UNIT, build_flat_case, build_cli_case,  # noqa: E402
                         filter_stdout, parse_stk, REF_TEST_DIR)

REF_BIN = os.path.join(ROOT, "oracle", "_ref", "RAMExtend")


def main():
    only = sys.argv[1:]   # e.g. "cli_overlap": regenerate just that CLI case
    if only:
        global FLAT_CASES, CLI_CASES
        FLAT_CASES = [c for c in FLAT_CASES if f"flat_{c}" in only]
        CLI_CASES = [c for c in CLI_CASES if f"cli_{c}" in only]
    ref = support.Reference.try_load()
    assert ref is not None, "build oracle/_ref first (make -C oracle ref)"

    # ---- unit vectors ---------------------------------------------------------------------
    u = dict(UNIT)
    sc = ref.scoring(u["matrix"])
    B, W = u["bandwidth"], 2 * u["bandwidth"] + 1
    rows = {}
    for side in ("left", "right"):
        direction = 0 if side == "left" else 1
        seq = np.array(u[f"p_{side}"], dtype=np.uint8)
        out = []
        for n in range(3):
            core = np.zeros(1, dtype=CORE_DTYPE)
            l, r, lo, hi, orient = u[f"{side}_cores"][n]
            core["left_seq_pos"], core["right_seq_pos"] = l, r
            core["lower_seq_bound"], core["upper_seq_bound"], core["orient"] = lo, hi, orient
            core["left_extendable"] = core["right_extendable"] = 1
            ps = np.array([0 if k == B else abs(k - B) * sc.gapextn + sc.gapopen for k in range(W)], dtype=np.int32)
            pg = ps.copy()
            per = []
            for row, cb in enumerate(u["cons_base"]):
                best, idx, cs, cg = ref.nw_row(direction, row, cb, core, ps, pg, seq, sc, B)
                per.append(dict(best=int(best), idx=int(idx), sub=cs.tolist(), gap=cg.tolist()))
                ps, pg = cs, cg
            out.append(per)
        rows[side] = out
    u["reference_rows"] = rows
    assert rows["right"][0][6]["gap"] == u["final_right_gap"]
    assert rows["right"][0][6]["sub"] == u["final_right_sub"]
    if not only:
        json.dump(u, open(os.path.join(HERE, "unit_vectors.json"), "w"))

    # ---- flat cases -----------------------------------------------------------------------
    for name in FLAT_CASES:
        case = build_flat_case(name, ref)
        sc = ref.scoring(case["matrix"], case.get("rs", (1, -1, -5)))
        if "gap_override" in case:
            sc.gapopen, sc.gapextn = case["gap_override"]
        p = make_params(**case["params"])
        outs = {}
        for d in (1, 0):
            o = ref.extend(d, case["cores"], case["sequence"], sc, p)
            outs[f"bp{d}"] = o["extended_bp"]
            outs[f"master{d}"] = o["master"]
            outs[f"ext_len{d}"] = o["ext_len"]
            outs[f"score_add{d}"] = o["score_add"]
        np.savez_compressed(os.path.join(HERE, f"flat_{name}.npz"), sequence=case["sequence"],
                            cores=case["cores"], meta=json.dumps(dict(
                                matrix=case["matrix"], rs=case.get("rs", (1, -1, -5)),
                                gap_override=case.get("gap_override"), params=case["params"])), **outs)
        print("flat", name, outs["bp1"], outs["bp0"])

    # ---- CLI cases -------------------------------------------------------------------------
    for name in CLI_CASES:
        case = build_cli_case(name)
        d = os.path.join(HERE, f"cli_{name}")
        os.makedirs(d, exist_ok=True)
        np.savez_compressed(os.path.join(d, "input.npz"),
                            seq=np.concatenate(case["seqs"]).astype(np.uint8),
                            lengths=np.array([len(x) for x in case["seqs"]], dtype=np.int64),
                            meta=json.dumps(dict(names=case["names"], bed=case["bed"], flags=case["flags"])))
        with tempfile.TemporaryDirectory() as tmp:
            tb, rg = os.path.join(tmp, "in.2bit"), os.path.join(tmp, "in.tsv")
            synth.write_twobit(tb, case["names"], case["seqs"])
            synth.write_ranges(rg, case["bed"])
            cmd = [REF_BIN, "-twobit", tb, "-ranges", rg, "-outtsv", os.path.join(tmp, "out.tsv"),
                   "-outfa", os.path.join(tmp, "out.fa"), "-cons", os.path.join(tmp, "cons.fa")] + case["flags"]
            r = subprocess.run(cmd, capture_output=True, text=True)
            assert r.returncode == 0, r.stdout + r.stderr
            open(os.path.join(d, "stdout.txt"), "w").write(filter_stdout(r.stdout, tmp))
            for f in ("out.tsv", "out.fa", "cons.fa"):
                src = os.path.join(tmp, f)
                open(os.path.join(d, f), "w").write(open(src).read() if os.path.exists(src) else "")
            # -outmat (ram_extend.c:1122-1132): the file is large, so its sha256, its line count and its
            # first and last 40 lines are the golden
            import hashlib
            mat = os.path.join(tmp, "mat.txt")
            r2 = subprocess.run(cmd + ["-outmat", mat], capture_output=True, text=True)
            assert r2.returncode == 0, r2.stdout + r2.stderr
            data = open(mat).read()
            lines = data.splitlines()
            json.dump(dict(sha256=hashlib.sha256(data.encode()).hexdigest(), lines=len(lines), head=lines[:40],
                           tail=lines[-40:]), open(os.path.join(d, "outmat.json"), "w"), indent=0)
        m = re.findall(r"Extended (?:right|left ): (\d+) bp", r.stdout)
        print("cli", name, m)

    # ---- ce10 families (configs 2-3) -------------------------------------------------------
    for fam in (() if only else (1, 2, 3)):
        rows = parse_stk(os.path.join(REF_TEST_DIR, f"ce10-fam{fam}.stk"))
        json.dump(rows, open(os.path.join(HERE, f"ce10_fam{fam}.json"), "w"))
        print("ce10 fam", fam, len(rows))


if __name__ == "__main__":
    main()
