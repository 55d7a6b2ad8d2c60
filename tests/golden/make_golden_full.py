"""Reference outputs at the BENCHMARKED sizes (run in the build container, where /root/reference
and oracle/_ref exist):  python tests/golden/make_golden_full.py [config4] [config5] [wide]

  full_config4.npz          BASELINE configs[3] exactly as bench.py generates it
                            (synth.make_family(4, 10000, 5000, 0.15), bandwidth 40, 14p43g): the
                            outputs of the reference's extend_alignment() object code for both
                            directions (~800 s of one core per run)
  full_config5_sample.npz   synth.config5_parity_sample(): 22 families of the configs[4] mix at FULL
                            size (every bandwidth, four size strata, six families with N >= 5000)
  full_wide_bands.npz       bandwidth 500 (util/davidExtendConsRAM.pl:248), 511 and 100 with a
                            family that spans more than 40 CTAs

Inputs are NOT stored: they are regenerated from the seeds; the fixture holds the sha256 of the
sequence bytes and of the core records so that a drifted generator is caught, plus the outputs.
"""
import hashlib
import json
import multiprocessing as mp
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from repeatafterme_b200 import synth  # noqa: E402
from repeatafterme_b200.abi import make_params  # noqa: E402
from tests import support  # noqa: E402
from tests.cases import full_case_family, FULL_CASES  # noqa: E402


def _run(task):
    group, key, direction = task
    fam = full_case_family(group, key)
    ref = support.Reference.try_load()
    p = dict(fam.params)
    matrix = p.pop("matrix")
    sc = ref.scoring(matrix)
    pr = make_params(**p)
    t0 = time.time()
    o = ref.extend(direction, fam.cores, fam.sequence, sc, pr)
    rows = o["rows_processed"]
    print(f"[{group}/{key} dir={direction}] n={fam.n} B={pr.bandwidth} bp={o['extended_bp']} rows={rows} "
          f"{time.time() - t0:.0f}s", flush=True)
    return dict(group=group, key=key, direction=direction, bp=int(o["extended_bp"]), rows=int(rows),
                master=o["master"][:rows].copy(), ext_len=o["ext_len"], score_add=o["score_add"],
                seq_sha=hashlib.sha256(fam.sequence.tobytes()).hexdigest(),
                cores_sha=hashlib.sha256(fam.cores.tobytes()).hexdigest(), n=fam.n,
                params=dict(fam.params), seconds=o["seconds"])


def main():
    groups = [g for g in sys.argv[1:] if g in FULL_CASES] or list(FULL_CASES)
    tasks = []
    for g in groups:
        for key, cost in FULL_CASES[g]():
            for d in (1, 0):
                tasks.append((cost, (g, key, d)))
    tasks.sort(key=lambda t: -t[0])
    procs = int(os.environ.get("GOLDEN_PROCS", max(1, (os.cpu_count() or 2) - 1)))
    with mp.get_context("spawn").Pool(procs) as pool:
        outs = pool.map(_run, [t for _, t in tasks], chunksize=1)
    for g in groups:
        arrays, meta = {}, {}
        for o in outs:
            if o["group"] != g:
                continue
            k, d = str(o["key"]), o["direction"]
            arrays[f"{k}_master{d}"] = o["master"]
            arrays[f"{k}_ext_len{d}"] = o["ext_len"]
            arrays[f"{k}_score_add{d}"] = o["score_add"]
            m = meta.setdefault(k, dict(n=o["n"], params=o["params"], seq_sha=o["seq_sha"],
                                        cores_sha=o["cores_sha"]))
            m[f"bp{d}"], m[f"rows{d}"], m[f"ref_seconds{d}"] = o["bp"], o["rows"], round(o["seconds"], 1)
        np.savez_compressed(os.path.join(HERE, f"full_{g}.npz"), meta=json.dumps(meta), **arrays)
        print("wrote", g, len(meta), "families")


if __name__ == "__main__":
    main()
