"""Crowded batch: lanes per instance of the wide bands (development aid). The automatic choice takes the fewest
lanes (41-cell lanes at 168 registers for bandwidth 60 / 80 / 100, which spill); this forces one or two lanes more
for those bands and compares the device time of the same right-pass batch."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from repeatafterme_b200 import RamCu, synth
from repeatafterme_b200.abi import make_params

nf = int(sys.argv[1]) if len(sys.argv) > 1 else 96
rc = RamCu(0)
specs = synth.config5_specs(5, 2000)
pick = np.random.default_rng(0).choice(len(specs), size=nf, replace=False)
fams = synth.generate_packed_families([specs[i] for i in sorted(pick)])
libs = [rc.library_from_twobit(f.blob, f.windows, f.n_start, f.n_len, f.total) for f in fams]
total = sum(f.n for f in fams)
POL = {"auto": {}, "plus1": {60: 4, 80: 5, 100: 6}, "plus2": {60: 5, 80: 6, 100: 7}, "b100_only_6": {100: 6}, "b80_100_8": {80: 8, 100: 8}}
for name, pol in POL.items():
    jobs = []
    for f, h in zip(fams, libs):
        p = dict(f.params); m = p.pop("matrix")
        p["concurrent_instances"] = total - f.n
        p["lanes_per_instance"] = pol.get(p["bandwidth"], 0)
        jobs += list(rc.job_create_pair(h, f.cores, rc.scoring(m), make_params(**p)))
    best = 1e9
    for rep in range(3):
        torch.cuda.synchronize(); t = time.perf_counter(); rc.batch_run(jobs); torch.cuda.synchronize(); best = min(best, time.perf_counter() - t)
    outs = [rc.job_fetch(j, fams[i // 2].n, 20000) for i, j in enumerate(jobs)]
    cu = sum(o[0].rows_processed * o[0].n_extendable * (2 * fams[i // 2].params["bandwidth"] + 1) * 4.0 for i, o in enumerate(outs))
    gs = sorted({(fams[i // 2].params["bandwidth"], o[0].lanes_per_instance) for i, o in enumerate(outs)})
    for j in jobs:
        rc.job_free(j)
    print(json.dumps(dict(policy=name, ms=round(best * 1e3, 1), gcups=round(cu / best / 1e9, 1), lanes=gs)), flush=True)
