"""RAMExtendBatch against one reference RAMExtend process per family (`xargs -P $(nproc)`, what
util/extend-stk.pl:365-371 amounts to) on the same files: K cheapest families of a seeded config-5
subset written as .2bit + ranges, outputs compared byte for byte, families/s of both (development aid and
the source of the numbers in DESIGN.md)."""
import json, os, subprocess, sys, tempfile, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from repeatafterme_b200 import synth


def main(k=24, pool=200, ref_timeout=600, procs=1):
    specs = synth.config5_specs()
    ids = np.random.default_rng(1).permutation(len(specs))[:pool].tolist()
    ids.sort(key=lambda i: specs[i]["n"] * (2 * specs[i]["params"]["bandwidth"] + 1) * specs[i]["flank"])
    ids = ids[:k]
    tmp = tempfile.mkdtemp(prefix="rambatch_")
    lines, ref_cmds = [], []
    cu_est = 0.0
    for i in ids:
        fam = synth.family_from_spec(specs[i])
        p = fam.params
        tb, rg = os.path.join(tmp, f"f{i}.2bit"), os.path.join(tmp, f"f{i}.tsv")
        synth.write_family_files(fam, tb, rg)
        flags = ["-L", str(p["L"]), "-bandwidth", str(p["bandwidth"]), "-matrix", p["matrix"], "-minimprovement",
                 str(p["minimprovement"]), "-vvv"]
        def args(tag):
            return ["-twobit", tb, "-ranges", rg, "-outtsv", os.path.join(tmp, f"{tag}{i}.tsv"), "-outfa",
                    os.path.join(tmp, f"{tag}{i}.fa"), "-cons", os.path.join(tmp, f"{tag}{i}.cons")] + flags
        lines.append(" ".join([f"stdout={os.path.join(tmp, f'b{i}.out')}"] + args("b")))
        ref_cmds.append(" ".join([os.path.join(ROOT, "oracle", "_ref", "RAMExtend")] + args("r")) +
                        " > " + os.path.join(tmp, f"r{i}.out"))
        cu_est += specs[i]["n"] * (2 * p["bandwidth"] + 1) * (specs[i]["flank"] + 100) * 8.0
    man = os.path.join(tmp, "manifest.txt")
    open(man, "w").write("\n".join(lines) + "\n")
    out = {"families": k, "instances": int(sum(specs[i]["n"] for i in ids)), "cell_updates_estimate": cu_est}
    t0 = time.perf_counter()
    r = subprocess.run([os.path.join(ROOT, "repeatafterme_b200", "bin", "RAMExtendBatch"), "-procs", str(procs), man],
                       capture_output=True, text=True)
    out["batch_wall_s"] = round(time.perf_counter() - t0, 3)
    assert r.returncode == 0, r.stderr[-2000:]
    out["procs"] = procs
    out["batch_report"] = r.stderr.strip().splitlines()[-procs:]
    out["batch_families_per_s"] = round(k / out["batch_wall_s"], 2)
    if os.path.exists(os.path.join(ROOT, "oracle", "_ref", "RAMExtend")) and ref_timeout > 0:
        cmdfile = os.path.join(tmp, "ref_cmds.txt")
        open(cmdfile, "w").write("\n".join(ref_cmds) + "\n")
        t0 = time.perf_counter()
        try:
            subprocess.run(f"xargs -P {os.cpu_count()} -I CMD sh -c CMD < {cmdfile}", shell=True, timeout=ref_timeout)
            out["reference_wall_s"] = round(time.perf_counter() - t0, 3)
            out["reference_processes_at_once"] = os.cpu_count()
            out["reference_families_per_s"] = round(k / out["reference_wall_s"], 3)
            same = 0
            for i in ids:
                ok = True
                for ext in ("tsv", "fa", "cons"):
                    a, b = os.path.join(tmp, f"b{i}.{ext}"), os.path.join(tmp, f"r{i}.{ext}")
                    ok = ok and (open(a).read() if os.path.exists(a) else "") == (open(b).read() if os.path.exists(b) else "")
                strip = lambda t: "\n".join(l for l in t.splitlines() if not (l.startswith("RAMExtend Version") or
                                            "Program duration" in l or "_FILE" in l))
                ok = ok and strip(open(os.path.join(tmp, f"b{i}.out")).read().replace(f"b{i}.", "X.")) == \
                    strip(open(os.path.join(tmp, f"r{i}.out")).read().replace(f"r{i}.", "X."))
                same += ok
            out["families_byte_identical"] = same
        except subprocess.TimeoutExpired:
            out["reference_wall_s"] = f"> {ref_timeout} (stopped)"
    print(json.dumps(out))


if __name__ == "__main__":
    main(*[int(x) for x in sys.argv[1:]])
