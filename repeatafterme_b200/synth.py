"""Synthetic repeat families and a minimal .2bit writer.

Everything here is *input generation* for tests and benchmarks: seeded numpy, no DP.
A family is produced directly in the flat form the extension path consumes -- the concatenated
per-core windows that ``loadSequenceSubsetMinimal()`` (reference sequence.c:505-923) would build
(``sequence`` bytes: A0 C1 G2 T3 N99; one ``CORE_DTYPE`` record per core) -- and can also be
written as ``.2bit`` + BED-6 ranges for the unchanged RAMExtend CLI.
"""
from __future__ import annotations

import struct
from dataclasses import dataclass, field

import numpy as np

from .abi import CORE_DTYPE

COMPL = np.array([3, 2, 1, 0], dtype=np.uint8)


@dataclass
class Family:
    """One MSA family in flat (post-loader) form + what is needed to write it as 2bit/tsv."""

    sequence: np.ndarray            # uint8 codes 0..3 / 99, concatenation of the per-core windows
    cores: np.ndarray               # CORE_DTYPE[n], in the loader's sorted order
    names: list = field(default_factory=list)       # 2bit sequence name per core
    seqs: list = field(default_factory=list)        # full 2bit sequence (uint8 codes) per core
    bed: list = field(default_factory=list)         # (name, start, end, left_flag, right_flag, strand)
    params: dict = field(default_factory=dict)      # bandwidth / matrix / minimprovement / L ...
    spans: list = field(default_factory=list)       # (window start, window end) of every core inside its 2bit sequence

    @property
    def n(self) -> int:
        return int(self.cores.size)


def mutate(rng: np.random.Generator, anc: np.ndarray, p: float) -> np.ndarray:
    """Per-base mutation with probability p: 80% substitution (uniform other base), 10% insertion
    of 1-3 random bases after the base, 10% deletion."""
    n = anc.size
    u = rng.random(n)
    kind = rng.random(n)
    mut = u < p
    sub = mut & (kind < 0.8)
    ins = mut & (kind >= 0.8) & (kind < 0.9)
    dele = mut & (kind >= 0.9)
    out = anc.copy()
    out[sub] = (anc[sub] + rng.integers(1, 4, size=int(sub.sum()))) % 4
    reps = np.ones(n, dtype=np.int64)
    reps[dele] = 0
    nins = rng.integers(1, 4, size=n)
    reps[ins] += nins[ins]
    res = np.repeat(out, reps)
    # positions of inserted bases: every repeat after the first of an `ins` base
    first = np.cumsum(reps) - reps
    is_first = np.zeros(res.size, dtype=bool)
    is_first[first[reps > 0]] = True
    extra = ~is_first
    res[extra] = rng.integers(0, 4, size=int(extra.sum()))
    return res.astype(np.uint8)


def revcomp(x: np.ndarray) -> np.ndarray:
    x = np.asarray(x, dtype=np.uint8)[::-1]
    return np.where(x > 3, 99, COMPL[np.minimum(x, 3)]).astype(np.uint8)


def make_family(seed: int, n: int, flank: int, divergence: float = 0.15, core_len: int = 20,
                pad: int = 500, one_sided_frac: float = 0.0, n_frac: float = 0.0,
                params: dict | None = None) -> Family:
    """Config-4 style family: a random ancestor of ``flank + core_len + flank`` bases; every
    instance = independently mutated left flank + exact core + mutated right flank, 50% reverse
    complemented, each in its own 2bit sequence ``s%07d`` with ``pad`` random bases either side.
    ``one_sided_frac`` of the cores are extendable on one side only (config 5);
    ``n_frac`` of the flank bases are turned into N."""
    rng = np.random.default_rng(seed)
    anc = rng.integers(0, 4, size=2 * flank + core_len, dtype=np.uint8)
    left_anc, core, right_anc = anc[:flank], anc[flank:flank + core_len], anc[flank + core_len:]
    cores = np.zeros(n, dtype=CORE_DTYPE)
    windows, names, seqs, bed, spans = [], [], [], [], []
    off = 0
    for i in range(n):
        lf = mutate(rng, left_anc, divergence)
        rf = mutate(rng, right_anc, divergence)
        rev = bool(rng.integers(0, 2))
        inst = np.concatenate([lf, core, rf])
        core_start = lf.size
        if rev:
            inst = revcomp(inst)
            core_start = rf.size
        full = np.concatenate([rng.integers(0, 4, size=pad, dtype=np.uint8), inst,
                               rng.integers(0, 4, size=pad, dtype=np.uint8)])
        if n_frac > 0:
            full = full.copy()
            full[rng.random(full.size) < n_frac] = 99
            # keep the core itself clean
            full[pad + core_start: pad + core_start + core_len] = inst[core_start: core_start + core_len]
        start, end = pad + core_start, pad + core_start + core_len
        lflag = rflag = 1
        if one_sided_frac > 0 and rng.random() < one_sided_frac:
            if rng.integers(0, 2):
                lflag = 0
            else:
                rflag = 0
        # window logic of loadSequenceSubsetMinimal for a core alone on its sequence with
        # max_flanking_bp >= the sequence size (sequence.c:625-743): the window reaches the
        # sequence end on each extendable side, and stops at the core on the other side.
        if rev:   # '-' strand: left = end side, right = start side
            w_lo = 0 if rflag else start
            w_hi = full.size if lflag else end
        else:
            w_lo = 0 if lflag else start
            w_hi = full.size if rflag else end
        win = full[w_lo:w_hi]
        lower_flank = start - w_lo
        c = cores[i]
        c["lower_seq_bound"], c["upper_seq_bound"] = off, off + win.size - 1
        c["left_extendable"], c["right_extendable"], c["orient"] = lflag, rflag, int(rev)
        if rev:   # sequence.c:886-891
            c["right_seq_pos"] = off + lower_flank
            c["left_seq_pos"] = off + lower_flank + core_len - 1
        else:     # sequence.c:892-897
            c["left_seq_pos"] = off + lower_flank
            c["right_seq_pos"] = off + lower_flank + core_len - 1
        off += win.size
        windows.append(win)
        spans.append((int(w_lo), int(w_hi)))
        name = "s%07d" % i
        names.append(name)
        seqs.append(full)
        bed.append((name, start, end, lflag, rflag, "-" if rev else "+"))
    return Family(sequence=np.concatenate(windows), cores=cores, names=names, seqs=seqs, bed=bed,
                  params=dict(params or {}), spans=spans)


def config4(seed: int = 4, n: int = 10000, flank: int = 5000) -> Family:
    """BASELINE.json configs[3]: 10k instances, ~15% divergence, 5 kb flanks, bandwidth 40."""
    return make_family(seed, n, flank, 0.15, params=dict(
        bandwidth=40, matrix="14p43g", minimprovement=30, cappenalty=-90, L=20000, when_to_stop=100))


def matrix_for_divergence(div_pct: float):
    """Matrix / minimprovement pick of util/extend-stk.pl:289-302."""
    if div_pct >= 22.5:
        return "25p43g", 3 * 9 - 3  # 25p43g default minimprovement 24 (ram_extend.c:335)
    if div_pct >= 19:
        return "20p43g", 27
    if div_pct >= 16:
        return "18p43g", 27
    return "14p43g", 30


def config5_specs(seed: int = 5, n_families: int = 2000):
    """BASELINE.json configs[4]: N log-uniform in [100,10000], bandwidth in {20,40,60,80,100},
    flank uniform in [500,5000], divergence uniform in [0.05,0.25], 10% one-sided cores."""
    rng = np.random.default_rng(seed)
    specs = []
    for f in range(n_families):
        n = int(round(10 ** rng.uniform(2, 4)))
        bw = int(rng.choice([20, 40, 60, 80, 100]))
        flank = int(rng.integers(500, 5001))
        div = float(rng.uniform(0.05, 0.25))
        matrix, minimp = matrix_for_divergence(div * 100)
        specs.append(dict(seed=int(seed * 1000003 + f), n=n, flank=flank, divergence=div,
                          one_sided_frac=0.10,
                          params=dict(bandwidth=bw, matrix=matrix, minimprovement=minimp,
                                      cappenalty=-90, L=20000, when_to_stop=100)))
    return specs


def family_from_spec(spec: dict) -> Family:
    return make_family(spec["seed"], spec["n"], spec["flank"], spec["divergence"],
                       one_sided_frac=spec.get("one_sided_frac", 0.0), params=spec["params"])


# ------------------------------------------------------------------------------------------
# .2bit writer / BED writer (format: SURVEY.md Appendix D; kentsrc/twoBitNew.c:250-290, 487-640)
# ------------------------------------------------------------------------------------------
_RAM_TO_2BIT = np.array([2, 1, 3, 0], dtype=np.uint8)   # A,C,G,T -> T0 C1 A2 G3 coding


def write_twobit(path: str, names, seqs):
    """Minimal UCSC .2bit (version 0): N runs become N-blocks, no mask blocks."""
    recs = []
    for s in seqs:
        s = np.asarray(s, dtype=np.uint8)
        isn = s > 3
        d = np.diff(np.concatenate([[0], isn.astype(np.int8), [0]]))
        nstarts = np.flatnonzero(d == 1)
        nsizes = np.flatnonzero(d == -1) - nstarts
        codes = np.where(isn, 0, _RAM_TO_2BIT[np.minimum(s, 3)]).astype(np.uint8)
        padn = (-codes.size) % 4
        codes = np.concatenate([codes, np.zeros(padn, dtype=np.uint8)]).reshape(-1, 4)
        packed = (codes[:, 0] << 6) | (codes[:, 1] << 4) | (codes[:, 2] << 2) | codes[:, 3]
        rec = struct.pack("<II", s.size, nstarts.size)
        rec += nstarts.astype("<u4").tobytes() + nsizes.astype("<u4").tobytes()
        rec += struct.pack("<II", 0, 0)
        rec += packed.astype(np.uint8).tobytes()
        recs.append(rec)
    header = struct.pack("<IIII", 0x1A412743, 0, len(names), 0)
    index_size = sum(1 + len(n) + 4 for n in names)
    off = len(header) + index_size
    index = b""
    for n, rec in zip(names, recs):
        index += struct.pack("<B", len(n)) + n.encode() + struct.pack("<I", off)
        off += len(rec)
    with open(path, "wb") as fh:
        fh.write(header + index + b"".join(recs))


def write_ranges(path: str, bed):
    with open(path, "w") as fh:
        for name, start, end, lflag, rflag, strand in bed:
            fh.write(f"{name}\t{start}\t{end}\t{lflag}\t{rflag}\t{strand}\n")


def write_family_files(fam: Family, twobit_path: str, ranges_path: str):
    write_twobit(twobit_path, fam.names, fam.seqs)
    write_ranges(ranges_path, fam.bed)


def read_twobit(path: str):
    """Minimal .2bit reader (version 0, little endian): returns (names, [uint8 code arrays]) with
    A0 C1 G2 T3 and 99 inside N-blocks (soft-mask blocks are ignored, as RAMExtend upper-cases)."""
    data = open(path, "rb").read()
    sig, version, count, _ = struct.unpack_from("<IIII", data, 0)
    assert sig == 0x1A412743 and version == 0, "unsupported .2bit"
    pos = 16
    names, offs = [], []
    for _ in range(count):
        ln = data[pos]
        names.append(data[pos + 1: pos + 1 + ln].decode())
        offs.append(struct.unpack_from("<I", data, pos + 1 + ln)[0])
        pos += 1 + ln + 4
    two2ram = np.array([3, 1, 0, 2], dtype=np.uint8)   # T0 C1 A2 G3 -> A0 C1 G2 T3
    seqs = []
    for off in offs:
        size, nb = struct.unpack_from("<II", data, off)
        p = off + 8
        nstarts = np.frombuffer(data, "<u4", nb, p)
        nsizes = np.frombuffer(data, "<u4", nb, p + 4 * nb)
        p += 8 * nb
        mb = struct.unpack_from("<I", data, p)[0]
        p += 4 + 8 * mb + 4
        packed = np.frombuffer(data, np.uint8, (size + 3) // 4, p)
        codes = np.stack([(packed >> 6) & 3, (packed >> 4) & 3, (packed >> 2) & 3, packed & 3], axis=1)
        s = two2ram[codes.reshape(-1)[:size]].copy()
        for a, b in zip(nstarts, nsizes):
            s[a:a + b] = 99
        seqs.append(s)
    return names, seqs


def pack_twobit_record(seq: np.ndarray) -> np.ndarray:
    """packed DNA of one .2bit record: 4 bases per byte, first base in the top two bits, T0 C1 A2 G3
    (N positions are stored as T, as UCSC's faToTwoBit does)"""
    seq = np.asarray(seq, dtype=np.uint8)
    codes = np.where(seq > 3, 0, _RAM_TO_2BIT[np.minimum(seq, 3)]).astype(np.uint8)
    codes = np.concatenate([codes, np.zeros((-codes.size) % 4, dtype=np.uint8)]).reshape(-1, 4)
    return ((codes[:, 0] << 6) | (codes[:, 1] << 4) | (codes[:, 2] << 2) | codes[:, 3]).astype(np.uint8)


def family_twobit_windows(fam: Family):
    """What a 2bit-native loader hands to ramcu_library_from_twobit for this family: the packed DNA
    of every record, one window per core (the span loadSequenceSubsetMinimal would extract) and the
    N-blocks clipped to the windows, in library coordinates."""
    blobs, wins, n_start, n_len = [], [], [], []
    off_bytes, dst = 0, 0
    for full, (lo, hi) in zip(fam.seqs, fam.spans):
        rec = pack_twobit_record(full)
        wins.append((dst, hi - lo, off_bytes, lo))
        isn = np.asarray(full[lo:hi]) > 3
        d = np.diff(np.concatenate([[0], isn.astype(np.int8), [0]]))
        for a, b in zip(np.flatnonzero(d == 1), np.flatnonzero(d == -1)):
            n_start.append(dst + int(a))
            n_len.append(int(b - a))
        blobs.append(rec)
        off_bytes += rec.size
        dst += hi - lo
    return (np.concatenate(blobs), np.array(wins, dtype=np.uint64), np.array(n_start, dtype=np.uint64),
            np.array(n_len, dtype=np.uint64), dst)


def config5_parity_sample(specs=None, seed: int = 11):
    """Seeded, stratified sample of the config-5 family mix used to pin parity at FULL size against
    the reference (tests/golden/full_config5_sample.npz): for every bandwidth one family from each
    of four size strata (N < 300, < 1000, < 3500, >= 3500; families above 3e10 cell updates are left
    out to bound the reference's CPU time) plus, for bandwidth 40 and 100, the cheapest family with
    N >= 5000 that is not picked already. Returns indices into `specs`."""
    specs = specs if specs is not None else config5_specs()
    rng = np.random.default_rng(seed)

    def cu(i):
        return specs[i]["n"] * (2 * specs[i]["params"]["bandwidth"] + 1) * (specs[i]["flank"] + 100) * 8.0

    edges = [0, 300, 1000, 3500, 1 << 30]
    pick = []
    for bw in (20, 40, 60, 80, 100):
        for lo, hi in zip(edges[:-1], edges[1:]):
            cell = [i for i, s in enumerate(specs)
                    if s["params"]["bandwidth"] == bw and lo <= s["n"] < hi and cu(i) <= 3e10]
            pick.append(int(cell[int(rng.integers(0, len(cell)))]))
    for bw in (40, 100):
        big = sorted((i for i, s in enumerate(specs)
                      if s["n"] >= 5000 and s["params"]["bandwidth"] == bw and i not in pick), key=cu)
        pick.append(big[0])
    return sorted(pick)


@dataclass
class PackedFamily:
    """A family whose library travels as the .2bit file stores DNA (4 bases per byte, first base in the
    top bits, T0 C1 A2 G3 -- kentsrc/twoBitNew.c:571-582) instead of the loader's byte per base: one
    record holding the concatenated windows + the N runs, in library coordinates."""

    blob: np.ndarray        # uint8, ceil(total / 4) bytes
    total: int              # bases
    n_start: np.ndarray     # uint64 N runs (library coordinates)
    n_len: np.ndarray
    cores: np.ndarray       # CORE_DTYPE[n]
    params: dict = field(default_factory=dict)

    @property
    def n(self) -> int:
        return int(self.cores.size)

    @property
    def windows(self) -> np.ndarray:
        return np.array([[0, self.total, 0, 0]], dtype=np.uint64)


def pack_family(fam: Family) -> PackedFamily:
    seq = np.asarray(fam.sequence, dtype=np.uint8)
    isn = seq > 3
    if isn.any():
        d = np.diff(np.concatenate([[0], isn.astype(np.int8), [0]]))
        n_start = np.flatnonzero(d == 1).astype(np.uint64)
        n_len = (np.flatnonzero(d == -1) - np.flatnonzero(d == 1)).astype(np.uint64)
    else:
        n_start = np.zeros(0, dtype=np.uint64)
        n_len = np.zeros(0, dtype=np.uint64)
    return PackedFamily(blob=pack_twobit_record(seq), total=int(seq.size), n_start=n_start, n_len=n_len,
                        cores=fam.cores, params=dict(fam.params))


def _packed_from_spec(spec):
    return pack_family(family_from_spec(spec))


def generate_packed_families(specs, procs: int = 0):
    """family_from_spec + pack_family for a list of specs on `procs` worker processes (generation is
    a Python loop per instance: ~30 ns per base on one core)"""
    import multiprocessing as mp
    import os
    procs = procs or max(1, (os.cpu_count() or 1))
    if procs == 1 or len(specs) <= 1:
        return [_packed_from_spec(s) for s in specs]
    # largest first so that the pool's tail is short
    order = sorted(range(len(specs)), key=lambda i: -specs[i]["n"] * specs[i]["flank"])
    out = [None] * len(specs)
    with mp.get_context("fork").Pool(procs) as pool:
        for i, pf in zip(order, pool.imap(_packed_from_spec, [specs[i] for i in order], chunksize=1)):
            out[i] = pf
    return out
