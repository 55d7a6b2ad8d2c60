"""Definitions of the golden cases (inputs only). Used by tests/golden/make_golden.py in the
build container; the tests themselves read the committed fixtures and never need /root/reference.
"""
from __future__ import annotations

import os
import re
import tempfile

import numpy as np

from repeatafterme_b200 import synth
from repeatafterme_b200.abi import CORE_DTYPE

REF_TEST_DIR = "/root/reference/test"

# ---------------------------------------------------------------------------------------------
# The reference's own unit-test vectors for compute_nw_row (facts of bnw_extend.c:1482-1567):
# two reciprocal 48-base libraries, three cores each (fwd, rev, fwd), bandwidth 5, 20p43g,
# 19 forced consensus bases, and the known-answer gap/sub rows at consensus row 6 of core 0.
# cores: (leftSeqPos, rightSeqPos, lower bound, upper bound as PASSED to compute_nw_row, orient)
# ---------------------------------------------------------------------------------------------
UNIT = dict(
    matrix="20p43g", bandwidth=5,
    p_left=[2, 1, 0, 0, 3, 2, 1, 3, 2, 0, 0, 3, 1, 3, 0, 2,
            1, 3, 0, 2, 0, 3, 3, 1, 0, 2, 1, 0, 3, 3, 2, 99,
            2, 1, 3, 2, 0, 1, 1, 3, 1, 0, 0, 3, 2, 3, 0, 2],
    p_right=[2, 0, 3, 1, 3, 0, 0, 2, 3, 1, 2, 3, 0, 0, 1, 2,
             99, 2, 3, 3, 0, 1, 2, 0, 1, 3, 3, 0, 2, 0, 3, 1,
             2, 0, 3, 2, 3, 0, 0, 1, 3, 1, 1, 0, 2, 3, 1, 2],
    cons_base=[1, 3, 0, 0, 1, 3, 1, 1, 0, 2, 3, 1, 2, 3, 0, 0, 1, 2, 2],
    left_cores=[(12, 14, 0, 15, 0), (16, 18, 16, 31, 1), (44, 46, 32, 47, 0)],
    right_cores=[(1, 3, 0, 15, 0), (29, 31, 16, 31, 1), (33, 35, 32, 47, 0)],
    final_right_gap=[-35, -21, -26, -17, -22, -7, -12, -17, -22, -27, -32],
    final_right_sub=[-31, -45, -31, -25, 26, -32, -25, -37, -47, -27, -57],
)

STK = dict(bandwidth=40, minimprovement=30, cappenalty=-90, L=20000, when_to_stop=100)
DEF = dict(bandwidth=14, minimprovement=27, cappenalty=-90, L=10000, when_to_stop=100)

FLAT_CASES = ["config1_default", "config1_stk", "config1_b100", "synth_small", "l_limited", "ends",
              "satellite", "repeatscout", "wide25", "gapoverride", "ragged", "identical", "n_heavy",
              "b20", "b60"]
CLI_CASES = ["config1_default", "config1_stk", "satellite", "synth", "tandem", "overlap"]


def satellite_inputs(seed=123456789):
    """test/genTest.pl:127-165 'satellite' pattern restated (numpy RNG instead of perl's srand)."""
    rng = np.random.default_rng(seed)
    flank, sat_len, sat_num, core_pos, core_len = 1000, 1000, 5, 250, 10
    sat = rng.integers(0, 4, sat_len, dtype=np.uint8)
    seq = np.concatenate([rng.integers(0, 4, flank, dtype=np.uint8)] + [sat] * sat_num +
                         [rng.integers(0, 4, flank, dtype=np.uint8)])
    bed, pos = [], flank + core_pos
    for _ in range(sat_num):
        bed.append(("seq1", pos - 1, pos + core_len, 1, 1, "+"))
        pos += sat_len
    return ["seq1"], [seq], bed


def tandem_inputs(seed=77):
    """Cores on both strands of two sequences, close enough for CORE_BOUNDARY / midpoint windows
    (sequence.c:625-743) and for OVERLAP AVOIDANCE between the passes (ram_extend.c:445-499)."""
    rng = np.random.default_rng(seed)
    unit = rng.integers(0, 4, 400, dtype=np.uint8)
    names, seqs, bed = [], [], []
    for s in range(3):
        parts, pos = [rng.integers(0, 4, 300, dtype=np.uint8)], 300
        for k in range(4):
            copy = synth.mutate(rng, unit, 0.08)
            rev = (k + s) % 2 == 1
            if rev:
                copy = synth.revcomp(copy)
            cs = pos + copy.size // 2 - 6
            bed.append((f"chr{s}", cs, cs + 12, 1, 1, "-" if rev else "+"))
            parts.append(copy)
            pos += copy.size
            gap = rng.integers(0, 4, int(rng.integers(5, 120)), dtype=np.uint8)
            parts.append(gap)
            pos += gap.size
        parts.append(rng.integers(0, 4, 200, dtype=np.uint8))
        names.append(f"chr{s}")
        seqs.append(np.concatenate(parts))
    return names, seqs, bed


def overlap_inputs(seed=2):
    """Right extensions that run into the left window of the next core on the same sequence, so that
    OVERLAP AVOIDANCE (ram_extend.c:445-499) narrows bounds between the passes: core A (right-
    extendable only) is followed by core B (left-extendable only); the loader gives both the whole
    gap (sequence.c:625-743), A's extension crosses it, and B's lower bound moves before the left
    pass. Six more sequences carry an unaffected B, so the left pass still extends."""
    rng = np.random.default_rng(seed)
    unit = rng.integers(0, 4, 500, dtype=np.uint8)
    names, seqs, bed = [], [], []
    for s in range(12):
        b = synth.mutate(rng, unit, 0.05)
        pre = rng.integers(0, 4, 250, dtype=np.uint8)
        post = rng.integers(0, 4, 250, dtype=np.uint8)
        if s < 6:
            a = synth.mutate(rng, unit, 0.05)
            gap = rng.integers(0, 4, 30, dtype=np.uint8)
            seq = np.concatenate([pre, a, gap, b, post])
            ca = pre.size + 100
            bed.append((f"c{s}", ca, ca + 12, 0, 1, "+"))
            cb = pre.size + a.size + gap.size + 300
        else:
            seq = np.concatenate([pre, b, post])
            cb = pre.size + 300
        bed.append((f"c{s}", cb, cb + 12, 1, 0, "+"))
        names.append(f"c{s}")
        seqs.append(seq)
    return names, seqs, bed


def ragged_family(seed, n, flank, pad=0):
    """Family whose instances run out of flank at very different rows (truncated copies)."""
    fam = synth.make_family(seed, n, flank, 0.12, pad=pad, one_sided_frac=0.15)
    rng = np.random.default_rng(seed + 1)
    cores = fam.cores.copy()
    for c in cores:
        span = c["upper_seq_bound"] - c["lower_seq_bound"]
        lo_cut = int(rng.integers(0, max(1, int(0.4 * span))))
        hi_cut = int(rng.integers(0, max(1, int(0.4 * span))))
        core_lo = min(c["left_seq_pos"], c["right_seq_pos"])
        core_hi = max(c["left_seq_pos"], c["right_seq_pos"])
        c["lower_seq_bound"] = min(c["lower_seq_bound"] + lo_cut, core_lo)
        c["upper_seq_bound"] = max(c["upper_seq_bound"] - hi_cut, core_hi)
    fam.cores = cores
    return fam


def build_flat_case(name, ref):
    """-> dict(sequence, cores, matrix, params[, rs, gap_override])"""
    if name.startswith("config1"):
        tb, rg = os.path.join(REF_TEST_DIR, "extension-test2.2bit"), os.path.join(REF_TEST_DIR, "extension-test2.tsv")
        params = dict(DEF) if name == "config1_default" else dict(STK)
        matrix = "20p43g" if name == "config1_default" else "14p43g"
        if name == "config1_b100":
            params.update(bandwidth=100, L=500)
        seq, cores, _, _ = ref.load(tb, rg, params["L"] + params["bandwidth"])
        return dict(sequence=seq, cores=cores, matrix=matrix, params=params)
    if name == "satellite":
        names, seqs, bed = satellite_inputs()
        with tempfile.TemporaryDirectory() as tmp:
            tb, rg = os.path.join(tmp, "s.2bit"), os.path.join(tmp, "s.tsv")
            synth.write_twobit(tb, names, seqs)
            synth.write_ranges(rg, bed)
            seq, cores, _, _ = ref.load(tb, rg, DEF["L"] + DEF["bandwidth"])
        return dict(sequence=seq, cores=cores, matrix="20p43g", params=dict(DEF))
    if name == "synth_small":
        f = synth.make_family(21, 40, 300, 0.15, one_sided_frac=0.2, n_frac=0.01)
        return dict(sequence=f.sequence, cores=f.cores, matrix="14p43g", params=dict(STK))
    if name == "l_limited":
        f = synth.make_family(22, 30, 200, 0.10)
        return dict(sequence=f.sequence, cores=f.cores, matrix="18p43g",
                    params=dict(bandwidth=5, minimprovement=27, cappenalty=-90, L=60, when_to_stop=100))
    if name == "ends":
        f = synth.make_family(23, 50, 150, 0.15, pad=0, one_sided_frac=0.3)
        return dict(sequence=f.sequence, cores=f.cores, matrix="14p43g", params=dict(STK))
    if name == "repeatscout":
        f = synth.make_family(24, 30, 250, 0.10, one_sided_frac=0.1)
        return dict(sequence=f.sequence, cores=f.cores, matrix="repeatscout", rs=(1, -1, -5),
                    params=dict(bandwidth=14, minimprovement=3, cappenalty=-20, L=10000, when_to_stop=100))
    if name == "wide25":
        f = synth.make_family(25, 24, 250, 0.24, one_sided_frac=0.1)
        return dict(sequence=f.sequence, cores=f.cores, matrix="25p43g",
                    params=dict(bandwidth=100, minimprovement=24, cappenalty=-90, L=20000, when_to_stop=100))
    if name == "gapoverride":
        f = synth.make_family(26, 30, 250, 0.15)
        return dict(sequence=f.sequence, cores=f.cores, matrix="20p43g", gap_override=(-20, -3),
                    params=dict(DEF))
    if name == "ragged":
        f = ragged_family(27, 60, 400)
        return dict(sequence=f.sequence, cores=f.cores, matrix="14p43g", params=dict(STK))
    if name == "identical":
        f = synth.make_family(28, 9, 120, 0.0, pad=30)
        return dict(sequence=f.sequence, cores=f.cores, matrix="20p43g", params=dict(DEF))
    if name == "n_heavy":
        f = synth.make_family(29, 30, 250, 0.10, n_frac=0.3)
        return dict(sequence=f.sequence, cores=f.cores, matrix="14p43g", params=dict(STK))
    if name == "b20":
        f = synth.make_family(30, 40, 300, 0.2, one_sided_frac=0.1)
        return dict(sequence=f.sequence, cores=f.cores, matrix="20p43g",
                    params=dict(bandwidth=20, minimprovement=27, cappenalty=-90, L=20000, when_to_stop=100))
    if name == "b60":
        f = synth.make_family(31, 40, 300, 0.18, one_sided_frac=0.1)
        return dict(sequence=f.sequence, cores=f.cores, matrix="18p43g",
                    params=dict(bandwidth=60, minimprovement=27, cappenalty=-90, L=20000, when_to_stop=100))
    raise KeyError(name)


STK_FLAGS = ["-L", "20000", "-bandwidth", "40", "-matrix", "14p43g", "-minimprovement", "30", "-vvv"]


def build_cli_case(name):
    """-> dict(names, seqs, bed, flags)"""
    if name.startswith("config1"):
        names, seqs = synth.read_twobit(os.path.join(REF_TEST_DIR, "extension-test2.2bit"))
        bed = []
        for line in open(os.path.join(REF_TEST_DIR, "extension-test2.tsv")):
            f = line.rstrip("\n").split("\t")
            bed.append((f[0], int(f[1]), int(f[2]), int(f[3]), int(f[4]), f[5]))
        return dict(names=names, seqs=seqs, bed=bed, flags=[] if name == "config1_default" else list(STK_FLAGS))
    if name == "satellite":
        names, seqs, bed = satellite_inputs()
        return dict(names=names, seqs=seqs, bed=bed, flags=[])
    if name == "synth":
        f = synth.make_family(41, 60, 400, 0.15, one_sided_frac=0.2, n_frac=0.01)
        return dict(names=f.names, seqs=f.seqs, bed=f.bed, flags=list(STK_FLAGS))
    if name == "tandem":
        names, seqs, bed = tandem_inputs()
        return dict(names=names, seqs=seqs, bed=bed, flags=["-bandwidth", "20", "-matrix", "18p43g", "-vvv"])
    if name == "overlap":
        names, seqs, bed = overlap_inputs()
        return dict(names=names, seqs=seqs, bed=bed, flags=["-bandwidth", "20", "-matrix", "18p43g", "-vvv"])
    raise KeyError(name)


def filter_stdout(text: str, tmpdir: str = "") -> str:
    """Drop the run-dependent lines of RAMExtend's stdout (banner with build number/date,
    'Program duration', file paths) -- SURVEY.md Appendix E."""
    out = []
    for line in text.splitlines():
        if line.startswith("RAMExtend Version") or "Program duration" in line:
            continue
        if "SEQUENCE_FILE" in line or "RANGES_FILE" in line:
            continue
        if tmpdir:
            line = line.replace(tmpdir, "<TMP>")
        out.append(line)
    return "\n".join(out) + "\n"


def parse_stk(path):
    """Stockholm seed alignment -> BED-6 rows by the rule of util/extend-stk.pl:312-359
    (start-1, end, left/right flags = at most 10 leading/trailing '.' columns, orient) plus the
    ungapped sequence of every row (stored in consensus orientation for '-' rows)."""
    rows = []
    for line in open(path):
        if line.startswith("#") or line.startswith("//") or not line.strip():
            continue
        name, aln = line.split()[:2]
        m = re.match(r"^(\S+):(\d+)-(\d+)_([+-])$", name)
        if not m:
            continue
        left = 1 if re.match(r"^[.]{0,10}[^.]", aln) else 0
        right = 1 if re.search(r"[^.][.]{0,10}$", aln) else 0
        rows.append(dict(chrom=m.group(1), start=int(m.group(2)) - 1, end=int(m.group(3)),
                         left=left, right=right, orient=m.group(4), seq=aln.replace(".", "").upper()))
    return rows


# ---------------------------------------------------------------------------------------------
# Full-size cases (tests/golden/make_golden_full.py): inputs come from seeds, the fixtures hold only
# the reference's outputs and the sha256 of the regenerated inputs.
# ---------------------------------------------------------------------------------------------
CONFIG4 = dict(bandwidth=40, matrix="14p43g", minimprovement=30, cappenalty=-90, L=20000, when_to_stop=100)

WIDE_CASES = {
    # bandwidth 500 is what util/davidExtendConsRAM.pl:248 passes; 511 the widest register-tiled band
    "b500": dict(seed=51, n=150, flank=700, div=0.15, params=dict(CONFIG4, bandwidth=500)),
    "b511": dict(seed=52, n=60, flank=400, div=0.20, one_sided=0.2,
                 params=dict(CONFIG4, bandwidth=511, matrix="20p43g", minimprovement=27)),
    # a family that spans far more than 40 CTAs at bandwidth 100 (RED exchange path)
    "b100_many_ctas": dict(seed=53, n=3000, flank=800, div=0.12, one_sided=0.1, params=dict(CONFIG4, bandwidth=100)),
}


def _cu(n, bandwidth, flank):
    return n * (2 * bandwidth + 1) * (flank + 100) * 8.0


def _config4_keys():
    return [("seed4", _cu(10000, 40, 5000))]


def _config5_keys():
    specs = synth.config5_specs()
    return [(i, _cu(specs[i]["n"], specs[i]["params"]["bandwidth"], specs[i]["flank"]))
            for i in synth.config5_parity_sample(specs)]


def _wide_keys():
    return [(k, _cu(v["n"], v["params"]["bandwidth"], v["flank"])) for k, v in WIDE_CASES.items()]


FULL_CASES = {"config4": _config4_keys, "config5_sample": _config5_keys, "wide_bands": _wide_keys}


def full_case_family(group, key):
    """the synth.Family (with .params) of one full-size case"""
    if group == "config4":
        fam = synth.make_family(4, 10000, 5000, 0.15)      # exactly bench.py's rank-0 family
        fam.params = dict(CONFIG4)
        return fam
    if group == "config5_sample":
        return synth.family_from_spec(synth.config5_specs()[int(key)])
    if group == "wide_bands":
        c = WIDE_CASES[str(key)]
        return synth.make_family(c["seed"], c["n"], c["flank"], c["div"], one_sided_frac=c.get("one_sided", 0.0),
                                 params=c["params"])
    raise KeyError(group)
