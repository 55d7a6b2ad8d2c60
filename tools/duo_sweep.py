"""Both passes of a family over sizes and bandwidths (development aid): what ramcu_job_create_pair decides on its own,
both passes in the same threads (RAMCU_PAIR_ALWAYS=1), and two single-pass units sharing the launch (RAMCU_NO_PAIR=1)
for every forced lanes-per-instance. Outputs are compared with the automatic run's. usage: duo_sweep.py [quick]"""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from repeatafterme_b200 import RamCu, synth
from repeatafterme_b200.abi import make_params, RamCuError

rc = RamCu(0)
GS = {40: (1, 2, 3, 4, 6, 8), 20: (1, 2, 3, 4), 100: (5, 6, 8, 16), 14: (1, 2, 3), 60: (3, 4, 5, 8), 80: (4, 5, 6, 8)}
NS = (50, 700, 2500, 4000, 6000, 8000, 10000)
BS = ((40, "14p43g"), (20, "20p43g"), (100, "25p43g"), (60, "18p43g"), (80, "25p43g"))
if len(sys.argv) > 1 and sys.argv[1] == "quick":
    NS, BS = (700, 6000, 10000), ((40, "14p43g"), (100, "25p43g"))


def run(lib, fam, sc, p, env):
    for k in ("RAMCU_NO_PAIR", "RAMCU_PAIR_ALWAYS"):
        os.environ.pop(k, None)
    if env:
        os.environ[env] = "1"
    best, out = 1e9, None
    for rep in range(3):
        jobs = list(rc.job_create_pair(lib, fam.cores, sc, p))
        rc.batch_run(jobs)
        out = [rc.job_fetch(j, fam.cores.size, p.L) for j in jobs]
        best = min(best, out[0][0].kernel_ms)
        for j in jobs:
            rc.job_free(j)
    os.environ.pop(env, None) if env else None
    return best, out


for B, matrix in BS:
    sc = rc.scoring(matrix)
    for n in NS:
        fam = synth.make_family(11, n, 1200, 0.15)
        lib = rc.upload_library(fam.sequence)
        p = make_params(bandwidth=B, minimprovement=30, L=20000)
        auto_ms, ref = run(lib, fam, sc, p, None)
        row = dict(B=B, N=n, auto_ms=round(auto_ms, 3), auto_G=ref[0][0].lanes_per_instance)
        pair_ms, o = run(lib, fam, sc, p, "RAMCU_PAIR_ALWAYS")
        row["pair_ms"], row["pair_G"] = round(pair_ms, 3), o[0][0].lanes_per_instance
        same = all(np.array_equal(x[1], y[1]) and np.array_equal(x[2], y[2]) and np.array_equal(x[3], y[3]) for x, y in zip(ref, o))
        duo = {}
        for G in GS[B]:
            try:
                pg = make_params(bandwidth=B, minimprovement=30, L=20000, lanes_per_instance=G)
                ms, o = run(lib, fam, sc, pg, "RAMCU_NO_PAIR")
                duo[G] = round(ms, 3)
                same = same and all(np.array_equal(x[1], y[1]) and np.array_equal(x[2], y[2]) and np.array_equal(x[3], y[3])
                                    for x, y in zip(ref, o))
            except RamCuError as e:
                duo[G] = None
        row["duo_ms"] = duo
        row["same"] = bool(same)
        best = min([v for v in duo.values() if v] + [pair_ms])
        row["auto_vs_best"] = round(auto_ms / best, 3)
        print(json.dumps(row), flush=True)
        rc.free_library(lib)
