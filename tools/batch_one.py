"""One ramcu_batch_run over a seeded subset of the config-5 family mix (development aid for ncu captures of the
batch launches): python tools/batch_one.py [families] [reps]"""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from repeatafterme_b200 import RamCu, synth
from repeatafterme_b200.abi import make_params

k = int(sys.argv[1]) if len(sys.argv) > 1 else 96
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
specs = synth.config5_specs()
ids = sorted(np.random.default_rng(0).permutation(len(specs))[:k].tolist())
fams = synth.generate_packed_families([specs[i] for i in ids])
rc = RamCu(0)
total = sum(f.n for f in fams)
libs, jobs, prs = [], [], []
for f in fams:
    pr = dict(f.params)
    sc = rc.scoring(pr.pop("matrix"))
    pr["concurrent_instances"] = total - f.n
    prm = make_params(**pr)
    h = rc.library_from_twobit(f.blob, f.windows, f.n_start, f.n_len, f.total)
    libs.append(h); prs.append(prm)
    jobs += list(rc.job_create_pair(h, f.cores, sc, prm))
for rep in range(reps + 1):
    t0 = time.perf_counter()
    rc.batch_run(jobs)
    outs = [rc.job_fetch(j, fams[i // 2].n, prs[i // 2].L) for i, j in enumerate(jobs)]
    dt = time.perf_counter() - t0
    cu = sum(o[0].rows_processed * o[0].n_extendable * (2 * prs[i // 2].bandwidth + 1) * 4.0 for i, o in enumerate(outs))
    print(json.dumps(dict(families=k, instances=total, wall_s=round(dt, 4), kernel_ms=round(outs[0][0].kernel_ms, 2),
                          gcups_device=round(cu / outs[0][0].kernel_ms / 1e6, 1))))
