"""Golden outputs of the reference binary's verbose dumps (-vvvv = VERBOSE 10, -vvvvv = 12, -vvvvvv = 20;
ram_extend.c:253-259): for each case and level the filtered stdout's sha256, line count and first / last
lines go to tests/golden/cli_<case>/verbose.json (the full text runs to tens of MB).

    python tests/golden/make_golden_verbose.py        # needs oracle/_ref/RAMExtend (make -C oracle)
"""
import hashlib
import json
import lzma
import os
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from repeatafterme_b200 import synth          # noqa: E402
from tests import golden_io as gio            # noqa: E402
from tests.cases import filter_stdout         # noqa: E402

REF_BIN = os.path.join(ROOT, "oracle", "_ref", "RAMExtend")
CASES = {"config1_default": ["-vvvv", "-vvvvv", "-vvvvvv"], "satellite": ["-vvvv", "-vvvvvv"],
         "overlap": ["-vvvv"], "tandem": ["-vvvvv"]}
# complete texts (xz) for the CPU test that checks the kernel's per-candidate outputs under the SIMT emulator
FULL_TEXT = {("config1_default", "-vvvvv"), ("satellite", "-vvvv")}


def main():
    for name, levels in CASES.items():
        case = gio.load_cli(name)
        out = {}
        with tempfile.TemporaryDirectory() as tmp:
            tb, rg = os.path.join(tmp, "in.2bit"), os.path.join(tmp, "in.tsv")
            synth.write_twobit(tb, case["names"], case["seqs"])
            synth.write_ranges(rg, case["bed"])
            for lv in levels:
                cmd = [REF_BIN, "-twobit", tb, "-ranges", rg, "-cons", os.path.join(tmp, "cons.fa")] + case["flags"] + [lv]
                r = subprocess.run(cmd, capture_output=True, text=True)
                assert r.returncode == 0, r.stdout[-2000:] + r.stderr
                text = filter_stdout(r.stdout, tmp)
                lines = text.splitlines()
                out[lv] = dict(sha256=hashlib.sha256(text.encode()).hexdigest(), lines=len(lines), head=lines[:60],
                               tail=lines[-60:])
                print(name, lv, len(lines), "lines", len(text) >> 10, "KiB")
                if (name, lv) in FULL_TEXT:
                    with lzma.open(os.path.join(HERE, f"cli_{name}", f"verbose{lv}.txt.xz"), "wb", preset=9) as f:
                        f.write(text.encode())
        json.dump(out, open(os.path.join(HERE, f"cli_{name}", "verbose.json"), "w"), indent=0)


if __name__ == "__main__":
    main()
