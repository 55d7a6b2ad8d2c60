import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from repeatafterme_b200 import RamCu, synth
from repeatafterme_b200.abi import make_params
n = int(sys.argv[1]); G = int(sys.argv[2]); flank = int(sys.argv[3]) if len(sys.argv) > 3 else 1500
rc = RamCu(0); sc = rc.scoring("14p43g")
fam = synth.make_family(4, n, flank, 0.15)
lib = rc.upload_library(fam.sequence)
p = make_params(bandwidth=40, minimprovement=30, L=20000, lanes_per_instance=G)
for rep in range(2):
    res, *_ = rc.extend(lib, 1, fam.cores, sc, p)
    print(json.dumps(dict(N=n, G=res.lanes_per_instance, rows=res.rows_processed, ms=round(res.kernel_ms, 3), us_per_row=round(1e3 * res.kernel_ms / res.rows_processed, 3))))
