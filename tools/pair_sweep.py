"""Pair launch vs two launches over family sizes (development aid): the CLI's right call uses the pair."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from repeatafterme_b200 import RamCu, synth
from repeatafterme_b200.abi import make_params
rc = RamCu(0)
for B, matrix in ((40, "14p43g"), (20, "20p43g"), (100, "25p43g")):
    sc = rc.scoring(matrix)
    for n in (50, 200, 700, 1500, 2500, 4000, 6000, 8000, 10000):
        fam = synth.make_family(11, n, 1200, 0.15)
        lib = rc.upload_library(fam.sequence)
        p = make_params(bandwidth=B, minimprovement=30, L=20000)
        best_seq, best_pair = 1e9, 1e9
        try:
            for rep in range(3):
                a = rc.extend(lib, 1, fam.cores, sc, p)
                b = rc.extend(lib, 0, fam.cores, sc, p)
                best_seq = min(best_seq, a[0].kernel_ms + b[0].kernel_ms)
                o = rc.extend_pair(lib, fam.cores, sc, p)
                best_pair = min(best_pair, o[0][0].kernel_ms)
                same = all(np.array_equal(x[1], y[1]) and np.array_equal(x[2], y[2]) for x, y in zip((a, b), o))
            print(json.dumps(dict(B=B, N=n, G_seq=a[0].lanes_per_instance, G_pair=o[0][0].lanes_per_instance,
                                  seq_ms=round(best_seq, 3), pair_ms=round(best_pair, 3),
                                  speedup=round(best_seq / best_pair, 2), same=bool(same))), flush=True)
        except Exception as e:
            print(f"B={B} N={n}: {e}", flush=True)
        rc.free_library(lib)
