"""Run one synthetic family through the CUDA path (development aid for ncu captures)."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from repeatafterme_b200 import RamCu, synth  # noqa: E402
from repeatafterme_b200.abi import make_params  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=10000)
ap.add_argument("--flank", type=int, default=5000)
ap.add_argument("--bandwidth", type=int, default=40)
ap.add_argument("--matrix", default="14p43g")
ap.add_argument("--L", type=int, default=20000)
ap.add_argument("--G", type=int, default=0)
ap.add_argument("--reps", type=int, default=2)
ap.add_argument("--dirs", default="1,0")
a = ap.parse_args()
fam = synth.make_family(4, a.n, a.flank, 0.15)
rc = RamCu(0)
sc = rc.scoring(a.matrix)
lib = rc.upload_library(fam.sequence)
p = make_params(bandwidth=a.bandwidth, minimprovement=30, L=a.L, lanes_per_instance=a.G)
for rep in range(a.reps):
    for d in [int(x) for x in a.dirs.split(",")]:
        res, master, ext_len, score_add = rc.extend(lib, d, fam.cores, sc, p)
        cu = res.rows_processed * res.n_extendable * (2 * a.bandwidth + 1) * 4
        print(json.dumps(dict(dir=d, G=res.lanes_per_instance, rows=res.rows_processed, bp=res.extended_bp,
                              ms=round(res.kernel_ms, 3), gcups=round(cu / max(res.kernel_ms, 1e-6) / 1e6, 1))))
