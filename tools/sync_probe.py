"""Per-column time vs number of CTAs (how much the grid exchange costs) -- development aid."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from repeatafterme_b200 import RamCu, synth
from repeatafterme_b200.abi import make_params
rc = RamCu(0)
sc = rc.scoring("14p43g")
G = int(sys.argv[1]) if len(sys.argv) > 1 else 3
for n in (10, 20, 40, 60, 70, 140, 280, 700, 2100, 4200, 10000):
    fam = synth.make_family(4, n, 1500, 0.15)
    lib = rc.upload_library(fam.sequence)
    p = make_params(bandwidth=40, minimprovement=30, L=20000, lanes_per_instance=G)
    best = None
    for rep in range(3):
        res, *_ = rc.extend(lib, 1, fam.cores, sc, p)
        best = res if best is None or res.kernel_ms < best.kernel_ms else best
    print(json.dumps(dict(N=n, G=best.lanes_per_instance, rows=best.rows_processed, ms=round(best.kernel_ms, 3),
                          us_per_row=round(1e3 * best.kernel_ms / best.rows_processed, 3),
                          gcups=round(best.rows_processed * best.n_extendable * 324 / best.kernel_ms / 1e6, 1))), flush=True)
    rc.free_library(lib)
