"""Timing breakdown of the family-batch path (development aid)."""
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
t0 = time.time()
fams = [synth.family_from_spec(specs[i]) for i in sorted(pick)]
print("gen %.1fs" % (time.time() - t0), [(f.n, f.params["bandwidth"], f.sequence.size // max(f.n,1)) for f in fams])
def sync(): torch.cuda.synchronize()
for rep in range(2):
    t = time.perf_counter(); libs = [rc.upload_library(f.sequence) for f in fams]; sync(); t_up = time.perf_counter() - t
    total = sum(f.n for f in fams)
    for direction in (1, 0):
        t = time.perf_counter()
        jobs = []
        for f, h in zip(fams, libs):
            p = dict(f.params); m = p.pop("matrix"); p["concurrent_instances"] = total - f.n
            jobs.append(rc.job_create(h, direction, f.cores, rc.scoring(m), make_params(**p)))
        sync(); t_cr = time.perf_counter() - t
        t = time.perf_counter(); rc.batch_run(jobs); sync(); t_run = time.perf_counter() - t
        t = time.perf_counter()
        outs = [rc.job_fetch(j, f.n, 20000) for j, f in zip(jobs, fams)]
        t_f = time.perf_counter() - t
        cu = sum(o[0].rows_processed * o[0].n_extendable * (2 * f.params["bandwidth"] + 1) * 4.0 for o, f in zip(outs, fams))
        print(f"rep{rep} dir{direction}: create {t_cr*1e3:.1f} ms, run {t_run*1e3:.1f} ms ({cu/t_run/1e9:.1f} GCUPS), fetch {t_f*1e3:.1f} ms; G=", [o[0].lanes_per_instance for o in outs])
        # solo timings
        if rep == 1 and direction == 1:
            for j, f in zip(jobs, fams):
                t = time.perf_counter(); rc.job_run(j); sync(); dt = time.perf_counter() - t
                r = rc.job_fetch(j, f.n, 20000)[0]
                print(f"   solo N={f.n} B={f.params['bandwidth']} rows={r.rows_processed} G={r.lanes_per_instance}: {dt*1e3:.2f} ms, {1e3*dt/max(r.rows_processed,1):.2f} us/row")
        for j in jobs: rc.job_free(j)
    t = time.perf_counter()
    for h in libs: rc.free_library(h)
    t_fr = time.perf_counter() - t
    print(f"rep{rep}: upload {t_up*1e3:.1f} ms, free {t_fr*1e3:.1f} ms")
