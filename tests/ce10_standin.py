"""Stand-in genome for BASELINE configs 2-3 (test/ce10.2bit is absent from the reference,
.MISSING_LARGE_BLOBS): chrI..chrX of ce10 size, seeded random background, every row of the three
seed alignments (tests/golden/ce10_fam*.json, made from test/ce10-fam{1,2,3}.stk by the rule of
util/extend-stk.pl:312-359) planted at its stated coordinates and strand, plus diverged copies of a
per-family ancestral flank on every extendable side so that the extension is non-trivial.
Deterministic: the reference binary (golden, build container) and the product binary (GPU box)
see byte-identical .2bit / .tsv files."""
from __future__ import annotations

import json
import os

import numpy as np

from repeatafterme_b200 import synth

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CE10_SIZES = {"chrI": 15072423, "chrII": 15279345, "chrIII": 13783700, "chrIV": 17493793,
              "chrV": 20924149, "chrX": 17718866}
CODE = {"A": 0, "C": 1, "G": 2, "T": 3}
FLANK = 700
STK_FLAGS = ["-L", "20000", "-bandwidth", "40", "-matrix", "14p43g", "-minimprovement", "30", "-vvv"]


def load_rows(fam: int):
    return json.load(open(os.path.join(GOLDEN, f"ce10_fam{fam}.json")))


def build(out_dir: str, seed: int = 2024):
    """writes <out_dir>/ce10_standin.2bit and fam{1,2,3}.tsv; returns their paths"""
    rng = np.random.default_rng(seed)
    names = list(CE10_SIZES)
    chroms = {n: rng.integers(0, 4, CE10_SIZES[n], dtype=np.uint8) for n in names}
    tsvs = {}
    for fam in (1, 2, 3):
        rows = load_rows(fam)
        left_anc = rng.integers(0, 4, FLANK, dtype=np.uint8)
        right_anc = rng.integers(0, 4, FLANK, dtype=np.uint8)
        bed = []
        for row in rows:
            g = chroms[row["chrom"]]
            seq = np.array([CODE[c] for c in row["seq"]], dtype=np.uint8)
            s, e = row["start"], row["end"]
            assert e - s == seq.size
            minus = row["orient"] == "-"
            g[s:e] = synth.revcomp(seq) if minus else seq
            lf = synth.mutate(rng, left_anc, 0.12) if row["left"] else None
            rf = synth.mutate(rng, right_anc, 0.12) if row["right"] else None
            # consensus-left flank sits before the core on '+', after it (reverse-complemented) on '-'
            before, after = (rf, lf) if minus else (lf, rf)
            if before is not None:
                b = synth.revcomp(before) if minus else before
                n = min(b.size, s)
                g[s - n:s] = b[b.size - n:]
            if after is not None:
                a = synth.revcomp(after) if minus else after
                n = min(a.size, g.size - e)
                g[e:e + n] = a[:n]
            bed.append((row["chrom"], s, e, row["left"], row["right"], row["orient"]))
        tsvs[fam] = os.path.join(out_dir, f"fam{fam}.tsv")
        synth.write_ranges(tsvs[fam], bed)
    twobit = os.path.join(out_dir, "ce10_standin.2bit")
    synth.write_twobit(twobit, names, [chroms[n] for n in names])
    return twobit, tsvs
