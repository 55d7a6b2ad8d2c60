"""Wall clock of the product CLI (bin/RAMExtend) on BASELINE config 4 written as .2bit + ranges files,
with the breakdown its boundary prints (RAMX_TIMING=1), next to the reference binary's host-side
overhead on the same files (its DP switched off with -stopafter 1 -minimprovement 100000000, as in
BASELINE.md section 2) -- development aid and the source of bench.py's `cli` leg."""
import json, os, subprocess, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from repeatafterme_b200 import synth

FLAGS = ["-bandwidth", "40", "-matrix", "14p43g", "-minimprovement", "30", "-L", "20000"]


def run(binary, tb, rg, tmp, extra=(), env=None):
    cmd = [binary, "-twobit", tb, "-ranges", rg, "-outtsv", os.path.join(tmp, "o.tsv"), "-outfa", os.path.join(tmp, "o.fa"),
           "-cons", os.path.join(tmp, "c.fa")] + FLAGS + list(extra)
    t0 = time.perf_counter()
    r = subprocess.run(cmd, capture_output=True, text=True, env=env)
    return time.perf_counter() - t0, r


def main(n=10000, flank=5000, reps=3):
    fam = synth.make_family(4, n, flank, 0.15)
    print('generated', file=sys.stderr, flush=True)
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        tb, rg = os.path.join(tmp, "f.2bit"), os.path.join(tmp, "f.tsv")
        synth.write_family_files(fam, tb, rg)
        env = dict(os.environ, RAMX_TIMING="1")
        walls = []
        for _ in range(reps):
            print('product run', file=sys.stderr, flush=True)
            w, r = run(os.path.join(ROOT, "repeatafterme_b200", "bin", "RAMExtend"), tb, rg, tmp, env=env)
            assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
            walls.append(w)
        out["product_wall_s"] = [round(w, 3) for w in walls]
        out["product_breakdown"] = [l for l in r.stderr.splitlines() if l.startswith("[ramx timing]")]
        out["product_extended"] = [l for l in r.stdout.splitlines() if l.startswith("Extended")]
        cons = open(os.path.join(tmp, "c.fa")).read()
        out["consensus_bytes"] = len(cons)
        print(json.dumps(out), flush=True)     # the product's numbers first: the reference run below may take minutes
        ref_bin = os.path.join(ROOT, "oracle", "_ref", "RAMExtend")
        if os.path.exists(ref_bin) and os.environ.get('CLI_BENCH_REFERENCE'):
            print('reference overhead run', file=sys.stderr, flush=True)
            w, r2 = run(ref_bin, tb, rg, tmp, extra=["-stopafter", "1", "-minimprovement", "100000000"])
            out["reference_host_overhead_wall_s"] = round(w, 3)
            print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main(*[int(x) for x in sys.argv[1:]])
