"""Right and left pass of one family in ONE launch (ramcu_batch_run of two jobs) vs two launches --
development aid: does co-residency of two independent column loops on an SM hide the grid exchange?"""
import json
import os
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from repeatafterme_b200 import RamCu, synth  # noqa: E402
from repeatafterme_b200.abi import make_params  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
flank = int(sys.argv[2]) if len(sys.argv) > 2 else 5000
fam = synth.make_family(4, n, flank, 0.15)
rc = RamCu(0)
sc = rc.scoring("14p43g")
lib = rc.upload_library(fam.sequence)
ref = {}
for d in (1, 0):
    p = make_params(bandwidth=40, minimprovement=30, L=20000)
    res, master, ext_len, score_add = rc.extend(lib, d, fam.cores, sc, p)
    ref[d] = (res.rows_processed, master.copy(), ext_len.copy(), score_add.copy())
    print(json.dumps(dict(mode="solo", dir=d, G=res.lanes_per_instance, rows=res.rows_processed, ms=round(res.kernel_ms, 3))), flush=True)
for G, maxt in [tuple(int(x) for x in a.split(':')) for a in (sys.argv[3] if len(sys.argv) > 3 else '3:512').split(',')]:
    try:
        p = make_params(bandwidth=40, minimprovement=30, L=20000, lanes_per_instance=G, cta_thread_limit=maxt)
        for rep in range(2):
            jobs = [rc.job_create(lib, d, fam.cores, sc, p) for d in (1, 0)]
            rc.batch_run(jobs)
            out = [rc.job_fetch(j, fam.cores.size, 20000) for j in jobs]
            ok = all(o[0].rows_processed == ref[d][0] and np.array_equal(o[1], ref[d][1]) and
                     np.array_equal(o[2], ref[d][2]) and np.array_equal(o[3], ref[d][3]) for o, d in zip(out, (1, 0)))
            cu = sum(o[0].rows_processed * o[0].n_extendable * 81 * 4 for o in out)
            ms = out[0][0].kernel_ms
            print(json.dumps(dict(mode="pair", G=G, maxt=maxt, ms=round(ms, 3), gcups=round(cu / ms / 1e6, 1), same=bool(ok))), flush=True)
            for j in jobs:
                rc.job_free(j)
    except Exception as e:  # noqa: BLE001
        print(f"G={G} maxt={maxt}: n/a ({e})", flush=True)
