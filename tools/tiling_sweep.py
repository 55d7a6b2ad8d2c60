"""Config 4 (N=10000, 5 kb flanks, B=40) over band tilings and register budgets: two launches vs the
pair launch, every result checked against the reference golden (development aid)."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from repeatafterme_b200 import RamCu, synth
from repeatafterme_b200.abi import make_params, RamCuError
from tests import golden_io as gio
from tests.cases import full_case_family

rc = RamCu(0)
case = gio.load_full("config4")["seed4"]
fam = full_case_family("config4", "seed4")
sc = rc.scoring("14p43g")
lib = rc.upload_library(fam.sequence)
CU = sum(case["want"][d]["rows_processed"] for d in (1, 0)) * 10000 * 81 * 4.0


def same(d, o):
    w = case["want"][d]
    return bool(o[0].extended_bp == w["extended_bp"] and np.array_equal(o[1][: w["rows_processed"]], w["master"][: w["rows_processed"]])
                and np.array_equal(o[2], w["ext_len"]) and np.array_equal(o[3], w["score_add"]))


only = [int(x) for x in sys.argv[1:]] or [1, 2, 3, 4, 5, 6, 8]
for G in only:
    for maxt in (0, 512):
        p = make_params(bandwidth=40, minimprovement=30, L=20000, lanes_per_instance=G, cta_thread_limit=maxt)
        try:
            best_seq, best_pair, ok = 1e9, 1e9, True
            for rep in range(3):
                a = rc.extend(lib, 1, fam.cores, sc, p)
                b = rc.extend(lib, 0, fam.cores, sc, p)
                best_seq = min(best_seq, a[0].kernel_ms + b[0].kernel_ms)
                o = rc.extend_pair(lib, fam.cores, sc, p)
                best_pair = min(best_pair, o[0][0].kernel_ms)
                ok = ok and same(1, a) and same(0, b) and same(1, o[0]) and same(0, o[1])
            print(json.dumps(dict(G=G, maxt=maxt, G_used=a[0].lanes_per_instance, G_pair=o[0][0].lanes_per_instance,
                                  seq_ms=round(best_seq, 3), pair_ms=round(best_pair, 3),
                                  seq_gcups=round(CU / best_seq / 1e6), pair_gcups=round(CU / best_pair / 1e6),
                                  equals_reference=ok)), flush=True)
        except RamCuError as e:
            print(f"G={G} maxt={maxt}: {str(e)[:100]}", flush=True)
rc.free_library(lib)
