"""Batch throughput vs lanes-per-instance policy (development aid)."""
import os, sys, time, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from repeatafterme_b200 import RamCu, synth
from repeatafterme_b200.abi import make_params, RamCuError
nf = int(sys.argv[1]) if len(sys.argv) > 1 else 24
rc = RamCu(0)
specs = synth.config5_specs(5, 2000)
pick = np.random.default_rng(0).choice(len(specs), size=nf, replace=False)
fams = [synth.family_from_spec(specs[i]) for i in sorted(pick)]
libs = [rc.upload_library(f.sequence) for f in fams]
total = sum(f.n for f in fams)
def run(policy, maxt=0):
    jobs = []
    for f, h in zip(fams, libs):
        p = dict(f.params); m = p.pop("matrix")
        if policy == "auto":
            p["concurrent_instances"] = total - f.n
            j = rc.job_create(h, 1, f.cores, rc.scoring(m), make_params(cta_thread_limit=maxt, **p))
        elif policy == "solo":
            j = rc.job_create(h, 1, f.cores, rc.scoring(m), make_params(**p))
        else:
            j = None
            for G in policy:
                try:
                    j = rc.job_create(h, 1, f.cores, rc.scoring(m), make_params(lanes_per_instance=G, cta_thread_limit=maxt, **p)); break
                except RamCuError:
                    continue
            if j is None:
                j = rc.job_create(h, 1, f.cores, rc.scoring(m), make_params(**p))
        jobs.append(j)
    best = 1e9
    for rep in range(3):
        torch.cuda.synchronize(); t = time.perf_counter(); rc.batch_run(jobs); torch.cuda.synchronize(); best = min(best, time.perf_counter() - t)
    outs = [rc.job_fetch(j, f.n, 20000) for j, f in zip(jobs, fams)]
    cu = sum(o[0].rows_processed * o[0].n_extendable * (2 * f.params["bandwidth"] + 1) * 4.0 for o, f in zip(outs, fams))
    gs = [o[0].lanes_per_instance for o in outs]
    for j in jobs: rc.job_free(j)
    print(json.dumps(dict(policy=str(policy), maxt=maxt, ms=round(best * 1e3, 1), gcups=round(cu / best / 1e9, 1), G=gs)), flush=True)
quick = len(sys.argv) > 2
for pol in (("auto", (2, 3, 4, 5, 8)) if quick else ("auto", "solo", (2, 3, 4, 5, 8), (3, 4, 5, 2, 8), (4, 5, 3, 6, 8, 2))):
    for maxt in ((0,) if quick else (0, 512)):
        run(pol, maxt)
