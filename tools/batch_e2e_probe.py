"""Where the end-to-end time of a family batch goes (development aid)."""
import os, sys, time, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from repeatafterme_b200 import RamCu, synth, batch
from repeatafterme_b200.abi import make_params
nf = int(sys.argv[1]) if len(sys.argv) > 1 else 12
rc = RamCu(0)
specs = synth.config5_specs(5, 2000)
pick = np.random.default_rng(0).choice(len(specs), size=nf, replace=False)
fams = [synth.family_from_spec(specs[i]) for i in sorted(pick)]
total = sum(f.n for f in fams)
batch.extend_families(rc, fams[:2])
for rep in range(2):
    T = {}
    def tick(k, t0):
        torch.cuda.synchronize(); T[k] = T.get(k, 0) + time.perf_counter() - t0
    t_all = time.perf_counter()
    t0 = time.perf_counter(); libs = [rc.upload_library(f.sequence) for f in fams]; tick("upload", t0)
    jobs = []
    t0 = time.perf_counter()
    for d in (1, 0):
        for f, h in zip(fams, libs):
            p = dict(f.params); m = p.pop("matrix"); p["concurrent_instances"] = total - f.n
            jobs.append((f, rc.job_create(h, d, f.cores, rc.scoring(m), make_params(**p))))
    tick("create", t0)
    t0 = time.perf_counter(); rc.batch_run([j for _, j in jobs]); tick("run", t0)
    t0 = time.perf_counter(); outs = [rc.job_fetch(j, f.n, 20000) for f, j in jobs]; tick("fetch", t0)
    t0 = time.perf_counter()
    for _, j in jobs: rc.job_free(j)
    for h in libs: rc.free_library(h)
    tick("free", t0)
    T["all"] = time.perf_counter() - t_all
    cu = sum(o[0].rows_processed * o[0].n_extendable * (2 * f.params["bandwidth"] + 1) * 4.0 for o, (f, _) in zip(outs, jobs))
    print(json.dumps(dict(n=nf, inst=total, mbases=round(sum(f.sequence.size for f in fams) / 1e6, 1),
                          **{k: round(v * 1e3, 1) for k, v in T.items()}, gcups_run=round(cu / T["run"] / 1e9, 1),
                          gcups_all=round(cu / T["all"] / 1e9, 1), fam_per_s=round(nf / T["all"], 1))), flush=True)
