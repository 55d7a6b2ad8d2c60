"""GPU probe: smoke, parity per band tiling, and raw kernel timings (development aid)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from repeatafterme_b200 import RamCu, synth  # noqa: E402
from repeatafterme_b200.abi import make_params, RamCuError  # noqa: E402
from tests import support  # noqa: E402


def main():
    import __graft_entry__ as ge
    ge.smoke()
    rc = RamCu(0)
    print(rc.device_info())
    ora = support.Oracle()
    # parity per tiling
    for B, matrix in ((40, "14p43g"), (14, "20p43g"), (20, "25p43g"), (60, "18p43g"), (100, "14p43g"), (33, "14p43g")):
        fam = synth.make_family(11 + B, 300, 400, 0.15, one_sided_frac=0.1, n_frac=0.005)
        osc = ora.scoring(matrix)
        sc = osc.to_product()
        lib = rc.upload_library(fam.sequence)
        for G in (0, 1, 2, 3, 4, 5, 6, 7, 8, 16, 32):
            p = make_params(bandwidth=B, minimprovement=27, L=20000, lanes_per_instance=G)
            for d in (1, 0):
                try:
                    res, master, ext_len, score_add = rc.extend(lib, d, fam.cores, sc, p)
                except RamCuError as ex:
                    if d == 1:
                        print(f"B={B} G={G}: n/a ({str(ex)[:60]})")
                    continue
                want = ora.extend(d, fam.cores, fam.sequence, osc, p)
                got = dict(res.as_dict(), master=master, ext_len=ext_len, score_add=score_add)
                try:
                    support.compare_extend(want, got, f"B={B} G={G} dir={d}")
                    ok = "OK"
                except AssertionError as ex:
                    ok = f"MISMATCH {ex}"
                print(f"B={B} G={G}->{res.lanes_per_instance} dir={d}: bp={res.extended_bp} rows={res.rows_processed} "
                      f"{res.kernel_ms:.2f} ms {ok}", flush=True)
        rc.free_library(lib)
    # timings
    for (n, flank) in ((1000, 2000), (10000, 5000)):
        t0 = time.time()
        fam = synth.make_family(4, n, flank, 0.15)
        print(f"generated N={n} flank={flank} in {time.time()-t0:.1f}s, {fam.sequence.size/1e6:.1f} Mbases")
        sc = rc.scoring("14p43g")
        t0 = time.time()
        lib = rc.upload_library(fam.sequence)
        print(f"upload+pack {1e3*(time.time()-t0):.1f} ms")
        for G in (0, 1, 2, 3, 4, 5, 6, 8):
            p = make_params(bandwidth=40, minimprovement=30, L=20000, lanes_per_instance=G)
            for rep in range(2):
                tot_ms, cu = 0.0, 0.0
                for d in (1, 0):
                    try:
                        res, master, ext_len, score_add = rc.extend(lib, d, fam.cores, sc, p)
                    except RamCuError as ex:
                        print(f"N={n} G={G}: n/a ({str(ex)[:80]})")
                        break
                    tot_ms += res.kernel_ms
                    cu += res.rows_processed * res.n_extendable * 81 * 4
                else:
                    print(json.dumps(dict(N=n, flank=flank, G=res.lanes_per_instance, ms=round(tot_ms, 3),
                                          gcups=round(cu / tot_ms / 1e6, 1), bp=res.extended_bp)), flush=True)
        rc.free_library(lib)


if __name__ == "__main__":
    main()
