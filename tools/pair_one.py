"""Both passes of the BASELINE config-4 family in one launch (development aid for ncu captures)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from repeatafterme_b200 import RamCu, synth  # noqa: E402
from repeatafterme_b200.abi import make_params  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
flank = int(sys.argv[2]) if len(sys.argv) > 2 else 5000
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
lanes = int(sys.argv[4]) if len(sys.argv) > 4 else 0   # force lanes per instance (0 = auto)
fam = synth.make_family(4, n, flank, 0.15)
rc = RamCu(0)
sc = rc.scoring("14p43g")
lib = rc.upload_library(fam.sequence)
p = make_params(bandwidth=40, minimprovement=30, L=20000, lanes_per_instance=lanes)
for rep in range(reps):
    out = rc.extend_pair(lib, fam.cores, sc, p)
    cu = sum(o[0].rows_processed * o[0].n_extendable * 81 * 4 for o in out)
    ms = out[0][0].kernel_ms
    print(json.dumps(dict(G=out[0][0].lanes_per_instance, passes_per_thread=out[0][0].passes_per_thread, rows=[o[0].rows_processed for o in out],
                          bp=[o[0].extended_bp for o in out], ms=round(ms, 3), gcups=round(cu / ms / 1e6, 1))))
