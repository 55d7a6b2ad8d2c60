"""GPU parity tests: the CUDA path, called through the C-ABI (libramcu.so), against the oracle,
the committed reference outputs (tests/golden) and size-independent properties at full size.
Bit-exact everywhere: the path is int32 max-plus."""
import os
import subprocess

import numpy as np
import pytest

from repeatafterme_b200 import batch, synth
from repeatafterme_b200.abi import CORE_DTYPE, RamCuError, make_params
from tests import golden_io as gio
from tests import support

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _got(res, master, ext_len, score_add):
    return dict(res.as_dict(), master=master, ext_len=ext_len, score_add=score_add)


@pytest.mark.parametrize("name", gio.flat_case_names())
def test_golden_flat_cases_every_tiling(ramcu, oracle, name):
    """reference outputs (golden) == CUDA path for the automatic tiling and every forced one"""
    case = gio.load_flat(name)
    sc = gio.oracle_scoring_for(case, oracle).to_product()
    lib = ramcu.upload_library(case["sequence"])
    ran = 0
    try:
        for G in (0, 1, 2, 3, 4, 5, 6, 7, 8, 16, 32):
            p = gio.params_for(case, lanes_per_instance=G)
            for d in (1, 0):
                try:
                    got = _got(*ramcu.extend(lib, d, case["cores"], sc, p))
                except RamCuError as ex:
                    assert "no band tiling" in str(ex), ex
                    continue
                gio.check_against_golden(case, d, got, f"{name} G={G} dir={d}")
                ran += 1
    finally:
        ramcu.free_library(lib)
    assert ran >= 4


def test_unit_vectors_through_debug_rows(ramcu, oracle):
    """bnw_extend.c:1384-1651 replayed through the kernel (forced consensus): every band row, the
    reciprocity of the two libraries and the known-answer row 6 (bnw_extend.c:1566-1567)"""
    u = gio.load_unit()
    B = u["bandwidth"]
    sc = oracle.scoring(u["matrix"]).to_product()
    rows = {}
    for side, direction in (("left", 0), ("right", 1)):
        seq = np.array(u[f"p_{side}"], dtype=np.uint8)
        cores = np.zeros(3, dtype=CORE_DTYPE)
        for n in range(3):
            l, r, lo, hi, orient = u[f"{side}_cores"][n]
            cores[n]["left_seq_pos"], cores[n]["right_seq_pos"] = l, r
            cores[n]["lower_seq_bound"], cores[n]["upper_seq_bound"], cores[n]["orient"] = lo, hi, orient
            cores[n]["left_extendable"] = cores[n]["right_extendable"] = 1
        nrows = len(u["cons_base"])
        lib = ramcu.upload_library(seq)
        r, best, sums = ramcu.debug_rows(lib, direction, cores, sc, make_params(bandwidth=B, L=nrows),
                                         nrows, forced=u["cons_base"])
        ramcu.free_library(lib)
        rows[side] = r
        for n in range(3):
            for row in range(nrows):
                want = u["reference_rows"][side][n][row]
                assert r[row, n, :, 0].tolist() == want["sub"] and r[row, n, :, 1].tolist() == want["gap"]
                assert best[row, n, u["cons_base"][row]] == want["best"]
    assert np.array_equal(rows["left"], rows["right"])
    assert rows["right"][6, 0, :, 1].tolist() == u["final_right_gap"]
    assert rows["right"][6, 0, :, 0].tolist() == u["final_right_sub"]


@pytest.mark.parametrize("seed,n,flank,B,matrix,kw", [
    (101, 300, 400, 40, "14p43g", dict(one_sided_frac=0.1, n_frac=0.005)),
    (102, 700, 300, 14, "20p43g", dict()),
    (103, 150, 250, 100, "25p43g", dict(one_sided_frac=0.2)),
    (104, 64, 200, 7, "18p43g", dict(pad=0)),
    (105, 200, 300, 33, "14p43g", dict(n_frac=0.05)),
    (106, 257, 200, 80, "18p43g", dict(pad=5)),
])
def test_random_families_vs_oracle(ramcu, oracle, seed, n, flank, B, matrix, kw):
    fam = synth.make_family(seed, n, flank, 0.15, **kw)
    osc = oracle.scoring(matrix)
    lib = ramcu.upload_library(fam.sequence)
    try:
        for L in (20000, flank // 2):
            p = make_params(bandwidth=B, minimprovement=27, L=L)
            for d in (1, 0):
                want = oracle.extend(d, fam.cores, fam.sequence, osc, p)
                got = _got(*ramcu.extend(lib, d, fam.cores, osc.to_product(), p))
                support.compare_extend(want, got, f"seed={seed} L={L} dir={d}")
    finally:
        ramcu.free_library(lib)


@pytest.mark.gpu
@pytest.mark.parametrize("scale,B,n", [(3000, 14, 60), (1500, 14, 60), (2000, 40, 300), (700, 40, 300)])
def test_scores_beyond_int16_vs_oracle(ramcu, oracle, scale, B, n):
    """the sweep reads the candidate scores as int16 (IDP.2A); a matrix beyond that range must take the exact
    path (the ABI accepts |score| * (L + 2B + 3) <= 1.2e7): scaled 14p43g on both sides of the int16 limit,
    every tiling of the band"""
    fam = synth.make_family(93 + n, n, 120, 0.15, pad=40, one_sided_frac=0.2, n_frac=0.01)
    osc = oracle.scoring("14p43g")
    for a in range(4):
        for b in range(9):
            osc.matrix[a][b] *= scale
    osc.gapopen *= scale
    osc.gapextn *= scale
    L = 60 if B == 14 else 30   # |gapopen| + |gapext| = 40 * scale: inside the accepted range
    lib = ramcu.upload_library(fam.sequence)
    ran = 0
    try:
        for G in (0, 1, 2, 3, 4, 8):
            p = make_params(bandwidth=B, minimprovement=27 * scale, L=L, when_to_stop=30, lanes_per_instance=G)
            for d in (1, 0):
                want = oracle.extend(d, fam.cores, fam.sequence, osc, p)
                try:
                    got = _got(*ramcu.extend(lib, d, fam.cores, osc.to_product(), p))
                except RamCuError as ex:
                    assert "no band tiling" in str(ex), ex
                    continue
                support.compare_extend(want, got, f"scale={scale} B={B} G={G} dir={d}")
                ran += 1
    finally:
        ramcu.free_library(lib)
    assert ran >= 4


def test_rows_and_sums_vs_oracle_trace(ramcu, oracle):
    """per-row column sums / chosen base (the argmax) and per-candidate bests vs the oracle"""
    fam = synth.make_family(7, 40, 150, 0.15, one_sided_frac=0.2)
    osc = oracle.scoring("14p43g")
    p = make_params(bandwidth=40, minimprovement=30, L=120, when_to_stop=1000)
    lib = ramcu.upload_library(fam.sequence)
    try:
        for d in (1, 0):
            want = oracle.extend(d, fam.cores, fam.sequence, osc, p, trace=True)
            rows, best, sums = ramcu.debug_rows(lib, d, fam.cores, osc.to_product(), p, 120)
            assert np.array_equal(sums, want["trace"][:, :4])
    finally:
        ramcu.free_library(lib)


def test_edge_cases(ramcu, oracle):
    osc = oracle.scoring("20p43g")
    sc = osc.to_product()
    fam = synth.make_family(8, 12, 60, 0.1)
    lib = ramcu.upload_library(fam.sequence)
    try:
        # nothing extendable in one direction
        cores = fam.cores.copy()
        cores["right_extendable"] = 0
        p = make_params(bandwidth=14, L=500)
        for d in (1, 0):
            want = oracle.extend(d, cores, fam.sequence, osc, p)
            got = _got(*ramcu.extend(lib, d, cores, sc, p))
            support.compare_extend(want, got, f"no-right dir={d}")
        # empty family, L = 0, L = 1, when_to_stop = 0/1, a single instance
        for cs, kw in ((fam.cores[:0], dict(L=300)), (fam.cores, dict(L=0)), (fam.cores, dict(L=1)),
                       (fam.cores, dict(L=300, when_to_stop=0)), (fam.cores, dict(L=300, when_to_stop=1)),
                       (fam.cores[:1], dict(L=300)), (fam.cores, dict(L=300, minimprovement=0)),
                       (fam.cores, dict(L=300, cappenalty=0))):
            p = make_params(bandwidth=14, **kw)
            for d in (1, 0):
                want = oracle.extend(d, cs, fam.sequence, osc, p)
                got = _got(*ramcu.extend(lib, d, cs, sc, p))
                support.compare_extend(want, got, f"edge {kw} n={cs.size} dir={d}")
        with pytest.raises(RamCuError):
            ramcu.extend(lib, 1, fam.cores, sc, make_params(bandwidth=1024))
    finally:
        ramcu.free_library(lib)


def test_batch_equals_individual_runs(ramcu, oracle):
    """many families in one launch (ramcu_batch_run) == each family on its own == oracle"""
    specs = synth.config5_specs(5, 12)
    fams = []
    for s in specs:
        s = dict(s, n=max(20, s["n"] // 40), flank=max(60, s["flank"] // 12))
        fams.append(synth.family_from_spec(s))
    res = batch.extend_families(ramcu, fams)
    for f, r in zip(fams, res):
        osc = oracle.scoring(f.params["matrix"])
        p = make_params(**{k: v for k, v in f.params.items() if k != "matrix"})
        for d, got in ((1, r.right), (0, r.left)):
            want = oracle.extend(d, f.cores, f.sequence, osc, p)
            support.compare_extend(want, got, f"batch dir={d}")


def test_batch_with_and_without_between_passes_hook(ramcu):
    """extend_families: all passes of all families in one batch (no hook) == right batch, hook, left
    batch (the form a caller with an overlap-avoidance step uses)"""
    specs = synth.config5_specs(7, 6)
    fams = [synth.family_from_spec(dict(s, n=max(20, s["n"] // 30), flank=max(60, s["flank"] // 10))) for s in specs]
    seen = []
    a = batch.extend_families(ramcu, fams)
    b = batch.extend_families(ramcu, fams, between_passes=lambda f, right: seen.append(right["extended_bp"]))
    assert len(seen) == len(fams)
    for x, y in zip(a, b):
        support.compare_extend(x.right, y.right, "hook vs no hook, right")
        support.compare_extend(x.left, y.left, "hook vs no hook, left")


@pytest.mark.parametrize("seed,n,flank,B,matrix,kw", [
    (111, 300, 300, 40, "14p43g", dict(one_sided_frac=0.2, n_frac=0.005)),   # both passes in few CTAs
    (112, 90, 250, 14, "20p43g", dict()),                                    # one CTA per pass
    (113, 2500, 150, 40, "14p43g", dict(one_sided_frac=0.1)),                # many CTAs, full registers
    (114, 400, 200, 100, "25p43g", dict(pad=3)),
    (115, 500, 200, 33, "18p43g", dict()),                                   # runtime-width tiling
])
def test_pair_launch_vs_oracle(ramcu, oracle, seed, n, flank, B, matrix, kw, monkeypatch):
    """ramcu_extend_pair (right and left pass of one family sharing one launch) == oracle ==
    two separate ramcu_extend calls, including families with one-side-extendable cores; in the form
    the planner picks, as two single-pass units sharing the launch (RAMCU_NO_PAIR) and with both
    passes in the same threads (RAMCU_PAIR_ALWAYS)"""
    fam = synth.make_family(seed, n, flank, 0.15, **kw)
    osc = oracle.scoring(matrix)
    lib = ramcu.upload_library(fam.sequence)
    try:
        p = make_params(bandwidth=B, minimprovement=27, L=20000)
        want = [oracle.extend(d, fam.cores, fam.sequence, osc, p) for d in (1, 0)]
        alone = [_got(*ramcu.extend(lib, d, fam.cores, osc.to_product(), p)) for d in (1, 0)]
        for env in (None, "RAMCU_NO_PAIR", "RAMCU_PAIR_ALWAYS"):
            monkeypatch.delenv("RAMCU_NO_PAIR", raising=False)
            monkeypatch.delenv("RAMCU_PAIR_ALWAYS", raising=False)
            if env:
                monkeypatch.setenv(env, "1")
            pair = [_got(*o) for o in ramcu.extend_pair(lib, fam.cores, osc.to_product(), p)]
            for d, got, w, a in zip((1, 0), pair, want, alone):
                support.compare_extend(w, got, f"pair ({env}) seed={seed} dir={d}")
                support.compare_extend(a, got, f"pair ({env}) vs alone seed={seed} dir={d}")
    finally:
        ramcu.free_library(lib)


def test_pair_launch_edge_cases(ramcu, oracle):
    """nothing extendable on one side / on both sides / empty family through the pair entry"""
    fam = synth.make_family(121, 60, 200, 0.15)
    osc = oracle.scoring("14p43g")
    p = make_params(bandwidth=40, minimprovement=30, L=20000)
    lib = ramcu.upload_library(fam.sequence)
    try:
        for mask in ((1, 0), (0, 1), (0, 0)):
            cores = fam.cores.copy()
            cores["right_extendable"] = mask[0]
            cores["left_extendable"] = mask[1]
            pair = [_got(*o) for o in ramcu.extend_pair(lib, cores, osc.to_product(), p)]
            for d, got in zip((1, 0), pair):
                want = oracle.extend(d, cores, fam.sequence, osc, p)
                support.compare_extend(want, got, f"pair mask={mask} dir={d}")
        empty = [_got(*o) for o in ramcu.extend_pair(lib, fam.cores[:0], osc.to_product(), p)]
        assert [e["extended_bp"] for e in empty] == [0, 0]
    finally:
        ramcu.free_library(lib)


def test_low_register_tilings_vs_oracle(ramcu, oracle):
    """the tilings compiled for half the registers (cta_thread_limit=512: the capacity route of the
    pair launch) against the oracle for every bandwidth that has them"""
    ran = 0
    for seed, B, matrix in ((131, 14, "20p43g"), (132, 20, "18p43g"), (133, 40, "14p43g"), (134, 60, "14p43g"),
                            (135, 80, "25p43g"), (136, 100, "18p43g")):
        fam = synth.make_family(seed, 120, 220, 0.15, one_sided_frac=0.1)
        osc = oracle.scoring(matrix)
        lib = ramcu.upload_library(fam.sequence)
        try:
            base = make_params(bandwidth=B, minimprovement=27, L=20000)
            want = [oracle.extend(d, fam.cores, fam.sequence, osc, base) for d in (1, 0)]
            for G in range(1, 9):
                p = make_params(bandwidth=B, minimprovement=27, L=20000, lanes_per_instance=G, cta_thread_limit=512)
                try:
                    pair = [_got(*o) for o in ramcu.extend_pair(lib, fam.cores, osc.to_product(), p)]
                except RamCuError as ex:
                    assert "no band tiling" in str(ex), ex
                    continue
                for w, got in zip(want, pair):
                    support.compare_extend(w, got, f"B={B} G={G} limit=512")
                ran += 1
        finally:
            ramcu.free_library(lib)
    assert ran >= 12


def test_full_size_properties(ramcu):
    """BASELINE config 4 scale (10k instances is too slow for the oracle): independent band
    tilings must agree bit for bit, instance order must not matter, and mirroring the library
    (reverse-complement sequence[], flip every strand) must leave both extensions unchanged."""
    fam = synth.make_family(4, 10000, 1200, 0.15)
    sc = ramcu.scoring("14p43g")
    lib = ramcu.upload_library(fam.sequence)
    outs = {}
    try:
        for G in (1, 2, 3, 4):
            p = make_params(bandwidth=40, minimprovement=30, L=20000, lanes_per_instance=G)
            outs[G] = [_got(*ramcu.extend(lib, d, fam.cores, sc, p)) for d in (1, 0)]
        for G in (2, 3, 4):
            for a, b in zip(outs[1], outs[G]):
                support.compare_extend(a, b, f"G=1 vs G={G}")
        assert outs[1][0]["extended_bp"] > 1000 and outs[1][1]["extended_bp"] > 1000
        # both passes in one launch, in the form the planner picks
        pair = [_got(*o) for o in ramcu.extend_pair(lib, fam.cores, sc, make_params(bandwidth=40, minimprovement=30, L=20000))]
        for a, b in zip(outs[1], pair):
            support.compare_extend(a, b, "separate passes vs pair launch")
        # permutation invariance
        perm = np.random.default_rng(0).permutation(fam.cores.size)
        p = make_params(bandwidth=40, minimprovement=30, L=20000)
        for d in (1, 0):
            got = _got(*ramcu.extend(lib, d, fam.cores[perm], sc, p))
            assert got["extended_bp"] == outs[1][1 - d]["extended_bp"]
            assert np.array_equal(got["master"], outs[1][1 - d]["master"])
            assert np.array_equal(got["ext_len"], outs[1][1 - d]["ext_len"][perm])
            assert np.array_equal(got["score_add"], outs[1][1 - d]["score_add"][perm])
    finally:
        ramcu.free_library(lib)
    # mirror: reverse-complement sequence[] (position p -> n-1-p) and flip every strand; the
    # consensus orientation of each core is unchanged, so both extensions must be identical
    n = fam.sequence.size
    mseq = synth.revcomp(fam.sequence)
    mc = fam.cores.copy()
    mc["left_seq_pos"] = n - 1 - fam.cores["left_seq_pos"]
    mc["right_seq_pos"] = n - 1 - fam.cores["right_seq_pos"]
    mc["lower_seq_bound"] = n - 1 - fam.cores["upper_seq_bound"]
    mc["upper_seq_bound"] = n - 1 - fam.cores["lower_seq_bound"]
    mc["orient"] = 1 - fam.cores["orient"]
    lib = ramcu.upload_library(mseq)
    try:
        p = make_params(bandwidth=40, minimprovement=30, L=20000)
        for d in (1, 0):
            got = _got(*ramcu.extend(lib, d, mc, sc, p))
            support.compare_extend(outs[1][1 - d], got, f"mirror dir={d}")
    finally:
        ramcu.free_library(lib)


@pytest.mark.parametrize("name", gio.cli_case_names())
def test_cli_outputs_byte_identical(name, tmp_path):
    """the product CLI (reference main() + CUDA extend_alignment) writes byte-identical -cons,
    -outtsv, -outfa and stdout to the reference RAMExtend (golden outputs)"""
    exe = os.path.join(ROOT, "repeatafterme_b200", "bin", "RAMExtend")
    if not os.path.exists(exe):
        pytest.skip("bin/RAMExtend not built (needs the reference's host objects)")
    from tests.cases import filter_stdout
    case = gio.load_cli(name)
    tb, rg = str(tmp_path / "in.2bit"), str(tmp_path / "in.tsv")
    synth.write_twobit(tb, case["names"], case["seqs"])
    synth.write_ranges(rg, case["bed"])
    cmd = [exe, "-twobit", tb, "-ranges", rg, "-outtsv", str(tmp_path / "out.tsv"), "-outfa",
           str(tmp_path / "out.fa"), "-cons", str(tmp_path / "cons.fa")] + case["flags"]
    # default: both passes in one launch at the right call, reused at the left call if overlap
    # avoidance changed no bound; RAMCU_PAIR=0: two launches
    for pair in ("1", "0"):
        for f in ("out.tsv", "out.fa", "cons.fa"):
            if (tmp_path / f).exists():
                (tmp_path / f).unlink()
        r = subprocess.run(cmd, capture_output=True, text=True, env=dict(os.environ, RAMCU_PAIR=pair))
        assert r.returncode == 0, r.stdout + r.stderr
        assert filter_stdout(r.stdout, str(tmp_path)) == case["want"]["stdout.txt"], pair
        for f in ("out.tsv", "out.fa", "cons.fa"):
            p = tmp_path / f
            assert (p.read_text() if p.exists() else "") == case["want"][f], (f, pair)


@pytest.mark.parametrize("name", gio.cli_case_names())
def test_cli_outmat_byte_identical(name, tmp_path):
    """-outmat (per-row path strings of every extendable core, ram_extend.c:1122-1132 /
    bnw_extend.c:1027-1044), written from the kernel's path states, against the reference binary's
    file (sha256, line count, first and last lines)."""
    import hashlib
    import json
    exe = os.path.join(ROOT, "repeatafterme_b200", "bin", "RAMExtend")
    if not os.path.exists(exe):
        pytest.skip("bin/RAMExtend not built (needs the reference's host objects)")
    gpath = os.path.join(ROOT, "tests", "golden", f"cli_{name}", "outmat.json")
    if not os.path.exists(gpath):
        pytest.skip("no -outmat golden for this case")
    want = json.load(open(gpath))
    case = gio.load_cli(name)
    tb, rg, mat = str(tmp_path / "in.2bit"), str(tmp_path / "in.tsv"), str(tmp_path / "mat.txt")
    synth.write_twobit(tb, case["names"], case["seqs"])
    synth.write_ranges(rg, case["bed"])
    cmd = [exe, "-twobit", tb, "-ranges", rg, "-outmat", mat, "-cons", str(tmp_path / "cons.fa")] + case["flags"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    data = open(mat).read()
    lines = data.splitlines()
    assert len(lines) == want["lines"]
    for i, (a, b) in enumerate(zip(lines[:40], want["head"])):
        assert a == b, f"line {i}"
    assert lines[-40:] == want["tail"]
    assert hashlib.sha256(data.encode()).hexdigest() == want["sha256"]
    assert (tmp_path / "cons.fa").read_text() == case["want"]["cons.fa"]


@pytest.mark.parametrize("name,level", [("config1_default", "-vvvv"), ("config1_default", "-vvvvv"),
                                        ("config1_default", "-vvvvvv"), ("satellite", "-vvvv"),
                                        ("satellite", "-vvvvvv"), ("overlap", "-vvvv"), ("tandem", "-vvvvv")])
def test_cli_verbose_dumps_byte_identical(name, level, tmp_path):
    """-vvvv / -vvvvv / -vvvvvv (VERBOSE 10 / 12 / 20): the per-row dumps of the reference's column loop
    (ram_extend.c:949-959, 992-1090, 1177-1214; bnw_extend.c:843-868), printed by the boundary from what the
    kernel reports for every candidate base, against the reference binary's stdout (sha256, line count,
    first and last lines; tests/golden/make_golden_verbose.py)."""
    import hashlib
    import json
    exe = os.path.join(ROOT, "repeatafterme_b200", "bin", "RAMExtend")
    if not os.path.exists(exe):
        pytest.skip("bin/RAMExtend not built (needs the reference's host objects)")
    from tests.cases import filter_stdout
    want = json.load(open(os.path.join(ROOT, "tests", "golden", f"cli_{name}", "verbose.json")))[level]
    case = gio.load_cli(name)
    tb, rg = str(tmp_path / "in.2bit"), str(tmp_path / "in.tsv")
    synth.write_twobit(tb, case["names"], case["seqs"])
    synth.write_ranges(rg, case["bed"])
    cmd = [exe, "-twobit", tb, "-ranges", rg, "-cons", str(tmp_path / "cons.fa")] + case["flags"] + [level]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr
    text = filter_stdout(r.stdout, str(tmp_path))
    lines = text.splitlines()
    for i, (a, b) in enumerate(zip(lines[:60], want["head"])):
        assert a == b, f"line {i}"
    assert len(lines) == want["lines"]
    assert lines[-60:] == want["tail"]
    assert hashlib.sha256(text.encode()).hexdigest() == want["sha256"]
    assert (tmp_path / "cons.fa").read_text() == case["want"]["cons.fa"]


@pytest.mark.parametrize("bandwidth,n,lanes,with_cells", [(40, 900, 0, False), (40, 120, 8, True), (100, 260, 0, False),
                                                         (23, 500, 0, False), (23, 90, 0, True)])
def test_verbose_values_consistent_with_the_extension(ramcu, oracle, bandwidth, n, lanes, with_cells):
    """ramcu_extend_verbose on families that span several CTAs and several lanes per instance (static and
    runtime-width bands): the per-candidate row bests reproduce the column sums through the reference's clamp /
    cap rule (ram_extend.c:1042-1062) and record setter (:1140-1150), the chosen base's best and column equal
    what ramcu_extend_paths reports (-outmat, pinned against the reference binary), and the extension itself is
    the oracle's."""
    fam = synth.make_family(500 + n, n, 260, 0.15, one_sided_frac=0.1, n_frac=0.003)
    osc = oracle.scoring("14p43g")
    sc = osc.to_product()
    p = make_params(bandwidth=bandwidth, minimprovement=30, L=600, lanes_per_instance=lanes)
    lib = ramcu.upload_library(fam.sequence)
    try:
        for d in (1, 0):
            want = oracle.extend(d, fam.cores, fam.sequence, osc, p)
            rows = want["rows_processed"]
            res, master, ext_len, score_add, cand, sums, cells = ramcu.extend_verbose(lib, d, fam.cores, sc, p, rows, cells=with_cells)
            support.compare_extend(want, _got(res, master, ext_len, score_add), f"verbose dir={d}")
            r2, m2, e2, s2, state, pbest, pidx = ramcu.extend_paths(lib, d, fam.cores, sc, p)
            assert np.array_equal(master, m2)
            ext = np.flatnonzero(fam.cores["right_extendable" if d else "left_extendable"] != 0)
            high = np.zeros(fam.cores.size, dtype=np.int64)
            for r in range(rows):
                best = cand[r, ext, :, 0].astype(np.int64)                       # [ext, candidate]
                contrib = np.maximum(np.maximum(best, 0), (high[ext] + p.cappenalty)[:, None])
                got_sums = contrib.sum(axis=0).astype(np.int64)
                assert np.array_equal((got_sums & 0xffffffff), sums[r].astype(np.int64) & 0xffffffff), (d, r)
                a = int(master[r])
                assert np.array_equal(cand[r, ext, a, 0], pbest[r, ext]), (d, r)
                assert np.array_equal(cand[r, ext, a, 1], r - bandwidth + pidx[r, ext]), (d, r)
                if cells is not None:
                    W = 2 * bandwidth + 1
                    cell = np.maximum(cells[r, ext, :, :, 0], cells[r, ext, :, :, 1])      # [ext, candidate, W]
                    assert np.array_equal(cell.max(axis=2), cand[r, ext, :, 0]), (d, r)
                    first = (cell == cell.max(axis=2, keepdims=True)).argmax(axis=2)
                    assert np.array_equal(r - bandwidth + first, cand[r, ext, :, 1]), (d, r)
                    assert np.array_equal(cells[r, ext, :, 0, 1], cand[r, ext, :, 2]), (d, r)
                    assert np.array_equal(cells[r, ext, :, W - 1, 1], cand[r, ext, :, 3]), (d, r)
                high[ext] = np.maximum(high[ext], best[:, a])
    finally:
        ramcu.free_library(lib)


def test_twobit_windows_aliasing_the_blob_tail(ramcu):
    """many windows that all read the last ~200 bases of an unpadded blob (the windows of a library may alias the
    same source bytes; every 128-base group of such a window takes the base-by-base route): the bounded slow list
    does not overflow and the decoded library is right"""
    from tests.test_kernel_emul import _aliasing_windows
    for n_windows in (12, 40):
        blob, wins, total, want = _aliasing_windows(n_windows=n_windows)
        none = np.zeros(0, dtype=np.uint64)
        h = ramcu.library_from_twobit(blob, wins, none, none, total)
        try:
            a2, an = ramcu.download_library(h, total)
            assert np.array_equal(_decode_packed(a2, an, total), want)
        finally:
            ramcu.free_library(h)


@pytest.fixture(scope="module")
def ce10_standin(tmp_path_factory):
    from tests import ce10_standin as cs
    d = tmp_path_factory.mktemp("ce10")
    twobit, tsvs = cs.build(str(d))
    return str(d), twobit, tsvs


@pytest.mark.parametrize("fam", [1, 2, 3])
def test_ce10_families_cli_byte_identical(ce10_standin, fam):
    """BASELINE configs 2-3: the extend-stk.pl command line (util/extend-stk.pl:365: -L 20000
    -bandwidth 40 -matrix 14p43g -minimprovement 30 -vvv) on test/ce10-fam{1,2,3}.stk, against a
    deterministic stand-in for the missing ce10.2bit; outputs of the reference binary are golden."""
    exe = os.path.join(ROOT, "repeatafterme_b200", "bin", "RAMExtend")
    if not os.path.exists(exe):
        pytest.skip("bin/RAMExtend not built (needs the reference's host objects)")
    from tests import ce10_standin as cs
    from tests.cases import filter_stdout
    d, twobit, tsvs = ce10_standin
    want = gio.load_cli_outputs(f"ce10_fam{fam}")
    outs = {f: os.path.join(d, f"fam{fam}_{f}") for f in ("out.tsv", "out.fa", "cons.fa")}
    cmd = [exe, "-twobit", twobit, "-ranges", tsvs[fam], "-outtsv", outs["out.tsv"], "-outfa", outs["out.fa"],
           "-cons", outs["cons.fa"]] + cs.STK_FLAGS
    for pair in ("1", "0"):
        r = subprocess.run(cmd, capture_output=True, text=True, env=dict(os.environ, RAMCU_PAIR=pair))
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr
        assert filter_stdout(r.stdout, d) == want["stdout.txt"], pair
        for f, p in outs.items():
            assert (open(p).read() if os.path.exists(p) else "") == want[f], (f, pair)
            if os.path.exists(p):
                os.unlink(p)


def _decode_packed(lib2, libn, n):
    i = np.arange(n)
    c = (lib2[i >> 4] >> (2 * (i & 15)).astype(np.uint32)) & 3
    isn = (libn[i >> 5] >> (i & 31).astype(np.uint32)) & 1
    return np.where(isn == 1, 99, c).astype(np.uint8)


@pytest.mark.parametrize("seed,kw", [(201, dict(n_frac=0.02, one_sided_frac=0.3)), (202, dict(pad=3)),
                                      (203, dict(n_frac=0.2))])
def test_twobit_native_library_equals_byte_library(ramcu, oracle, seed, kw):
    """flank feed: the library decoded on the device from the .2bit records' own packed DNA
    (ramcu_library_from_twobit) is identical to the one packed from the reference loader's bytes
    (ramcu_library_upload), and extends to the same result as the oracle"""
    fam = synth.make_family(seed, 37, 211, 0.15, **kw)
    blob, wins, ns, nl, total = synth.family_twobit_windows(fam)
    assert total == fam.sequence.size
    h2 = ramcu.library_from_twobit(blob, wins, ns, nl, total)
    h1 = ramcu.upload_library(fam.sequence)
    try:
        a2, an = ramcu.download_library(h2, total)
        b2, bn = ramcu.download_library(h1, total)
        want = np.where(fam.sequence > 3, 99, fam.sequence)
        assert np.array_equal(_decode_packed(b2, bn, total), want)
        assert np.array_equal(_decode_packed(a2, an, total), want)
        osc = oracle.scoring("14p43g")
        p = make_params(bandwidth=40, minimprovement=30, L=20000)
        for d in (1, 0):
            ref = oracle.extend(d, fam.cores, fam.sequence, osc, p)
            got = _got(*ramcu.extend(h2, d, fam.cores, osc.to_product(), p))
            support.compare_extend(ref, got, f"twobit library dir={d}")
    finally:
        ramcu.free_library(h1)
        ramcu.free_library(h2)
