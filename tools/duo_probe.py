"""Config 4 (N=10000, 5 kb flanks, B=40): the two passes as (a) two launches, (b) two single-pass units sharing one
launch (RAMCU_NO_PAIR=1: each pass keeps its own layout), (c) both passes in the same threads; per band tiling,
every result checked against the reference golden (development aid).  usage: duo_probe.py [G ...]"""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from repeatafterme_b200 import RamCu
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


def run_pair(p):
    best, ok, G = 1e9, True, 0
    for rep in range(3):
        jobs = list(rc.job_create_pair(lib, fam.cores, sc, p))
        rc.batch_run(jobs)
        outs = [rc.job_fetch(j, fam.cores.size, p.L) for j in jobs]
        best = min(best, outs[0][0].kernel_ms)
        ok = ok and same(1, outs[0]) and same(0, outs[1])
        G = outs[0][0].lanes_per_instance
        for j in jobs:
            rc.job_free(j)
    return best, ok, G


for G in [int(x) for x in sys.argv[1:]] or [1, 2, 3, 4]:
    p = make_params(bandwidth=40, minimprovement=30, L=20000, lanes_per_instance=G)
    try:
        seq = 1e9
        for rep in range(2):
            a = rc.extend(lib, 1, fam.cores, sc, p)
            b = rc.extend(lib, 0, fam.cores, sc, p)
            seq = min(seq, a[0].kernel_ms + b[0].kernel_ms)
        ok0 = same(1, a) and same(0, b)
        os.environ["RAMCU_NO_PAIR"] = "1"
        duo, ok1, g1 = run_pair(p)
        del os.environ["RAMCU_NO_PAIR"]
        pair, ok2, g2 = run_pair(p)
        print(json.dumps(dict(G=G, seq_ms=round(seq, 3), duo_ms=round(duo, 3), duo_G=g1, pair_ms=round(pair, 3), pair_G=g2,
                              duo_gcups=round(CU / duo / 1e6), pair_gcups=round(CU / pair / 1e6),
                              equals_reference=bool(ok0 and ok1 and ok2))), flush=True)
    except RamCuError as e:
        print(f"G={G}: {str(e)[:160]}", flush=True)
rc.free_library(lib)
