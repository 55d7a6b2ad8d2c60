"""Wall time of the product CLI (reference main() + our extend_alignment) with the pair launch on/off."""
import os, subprocess, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from repeatafterme_b200 import synth
exe = os.path.join(ROOT, "repeatafterme_b200", "bin", "RAMExtend")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
fam = synth.make_family(21, n, 2000, 0.15)
with tempfile.TemporaryDirectory() as tmp:
    tb, rg = os.path.join(tmp, "in.2bit"), os.path.join(tmp, "in.tsv")
    synth.write_twobit(tb, fam.names, fam.seqs)
    synth.write_ranges(rg, fam.bed)
    outs = {}
    for pair in ("1", "0") * 4:
        cmd = [exe, "-twobit", tb, "-ranges", rg, "-cons", os.path.join(tmp, f"cons{pair}.fa"), "-L", "20000",
               "-bandwidth", "40", "-matrix", "14p43g", "-minimprovement", "30"]
        t0 = time.perf_counter()
        r = subprocess.run(cmd, capture_output=True, text=True, env=dict(os.environ, RAMCU_PAIR=pair))
        dt = time.perf_counter() - t0
        ext = [l for l in r.stdout.splitlines() if l.startswith("Extended")]
        outs[pair] = open(os.path.join(tmp, f"cons{pair}.fa")).read()
        print(f"N={n} RAMCU_PAIR={pair}: {dt:.2f} s wall, rc={r.returncode}, {ext}", flush=True)
    print("consensus identical:", outs["1"] == outs["0"])
