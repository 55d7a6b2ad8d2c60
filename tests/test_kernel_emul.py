"""The real kernel source (csrc/extend_kernel.cuh) + the real host flatten logic (csrc/flatten.h)
compiled for the CPU against a lock-step SIMT emulation (tests/emul), diffed against the oracle.
This is a logic check that runs without a GPU; the GPU parity tests are tests/test_gpu_parity.py.
"""
import numpy as np
import pytest

from repeatafterme_b200 import synth
from repeatafterme_b200.abi import CORE_DTYPE, make_params
from tests import golden_io as gio
from tests import support


@pytest.mark.parametrize("cfg,threads", [
    ((11, 1, 11), 32), ((6, 2, 11), 32), ((3, 4, 11), 64), ((4, 4, 0), 32), ((16, 32, 0), 64),
])
def test_emul_b5_family(emul, oracle, cfg, threads):
    fam = synth.make_family(3, 14, 70, 0.15, pad=0, one_sided_frac=0.3, n_frac=0.02)
    osc = oracle.scoring("18p43g")
    p = make_params(bandwidth=5, minimprovement=27, L=400, when_to_stop=25)
    for d in (1, 0):
        want = oracle.extend(d, fam.cores, fam.sequence, osc, p)
        got = emul.extend(cfg, d, fam.cores, fam.sequence, osc.to_product(), p, threads=threads)
        support.compare_extend(want, got, f"cfg={cfg} dir={d}")


@pytest.mark.parametrize("cfg,B,matrix", [
    ((8, 4, 29), 14, "20p43g"), ((29, 1, 29), 14, "20p43g"), ((21, 4, 81), 40, "14p43g"),
    ((11, 8, 81), 40, "14p43g"), ((32, 4, 0), 40, "14p43g"),
])
def test_emul_multi_cta_ragged(emul, oracle, cfg, B, matrix):
    """several CTAs per family (grid exchange), instances that run out of flank at different
    rows (per-cell bounds path, retired warps)"""
    from tests.cases import ragged_family
    fam = ragged_family(5, 20, 90)
    osc = oracle.scoring(matrix)
    p = make_params(bandwidth=B, minimprovement=27, L=300, when_to_stop=20)
    for d in (1, 0):
        want = oracle.extend(d, fam.cores, fam.sequence, osc, p)
        got = emul.extend(cfg, d, fam.cores, fam.sequence, osc.to_product(), p, threads=32)
        support.compare_extend(want, got, f"cfg={cfg} dir={d}")


@pytest.mark.parametrize("cfg", [(11, 1, 11), (3, 4, 11), (4, 4, 0)])
def test_emul_unit_vectors_forced_consensus(emul, cfg):
    """bnw_extend.c:1384-1651 through the kernel: forced consensus, every committed band row"""
    u = gio.load_unit()
    B = u["bandwidth"]
    from tests.support import Oracle
    sc = Oracle().scoring(u["matrix"]).to_product()
    for side, direction in (("left", 0), ("right", 1)):
        seq = np.array(u[f"p_{side}"], dtype=np.uint8)
        cores = np.zeros(3, dtype=CORE_DTYPE)
        for n in range(3):
            l, r, lo, hi, orient = u[f"{side}_cores"][n]
            cores[n]["left_seq_pos"], cores[n]["right_seq_pos"] = l, r
            cores[n]["lower_seq_bound"], cores[n]["upper_seq_bound"], cores[n]["orient"] = lo, hi, orient
            cores[n]["left_extendable"] = cores[n]["right_extendable"] = 1
        nrows = len(u["cons_base"])
        p = make_params(bandwidth=B, L=nrows)
        got = emul.extend(cfg, direction, cores, seq, sc, p, threads=32, dbg_rows=nrows,
                          forced=u["cons_base"])
        for n in range(3):
            for row in range(nrows):
                want = u["reference_rows"][side][n][row]
                assert got["rows"][row, n, :, 0].tolist() == want["sub"], (side, n, row)
                assert got["rows"][row, n, :, 1].tolist() == want["gap"], (side, n, row)
                assert got["best"][row, n, u["cons_base"][row]] == want["best"]
        if side == "right":
            assert got["rows"][6, 0, :, 1].tolist() == u["final_right_gap"]
            assert got["rows"][6, 0, :, 0].tolist() == u["final_right_sub"]


@pytest.mark.parametrize("name,cfg", [("config1_default", (8, 4, 29)), ("l_limited", (4, 4, 0)),
                                      ("identical", (15, 2, 29))])
def test_emul_golden(emul, oracle, name, cfg):
    case = gio.load_flat(name)
    sc = gio.oracle_scoring_for(case, oracle).to_product()
    p = gio.params_for(case)
    for d in (1, 0):
        got = emul.extend(cfg, d, case["cores"], case["sequence"], sc, p, threads=64)
        gio.check_against_golden(case, d, got, f"{name} dir={d}")


@pytest.mark.parametrize("cfg", [(8, 4, 29), (29, 1, 29), (10, 3, 29)])
def test_emul_outmat_config1(emul, oracle, cfg):
    """-outmat path strings (ram_extend.c:1122-1132, bnw_extend.c:1027-1044) from the kernel's path
    states for BASELINE config 1 == the file the reference binary writes (sha256 + head + tail)"""
    import hashlib
    import json
    import os
    if cfg not in emul.configs():
        pytest.skip("tiling not in the emulator build")
    case = gio.load_flat("config1_default")
    want = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "cli_config1_default", "outmat.json")))
    sc = gio.oracle_scoring_for(case, oracle).to_product()
    p = gio.params_for(case)
    lines = []
    for d in (1, 0):
        got = support.emul_extend_paths(emul, cfg, d, case["cores"], case["sequence"], sc, p)
        lines += support.format_outmat(d, case["cores"], case["sequence"], p.bandwidth, got["master"],
                                       got["state"], got["best"], got["idx"])
    assert len(lines) == want["lines"]
    for i, (a, b) in enumerate(zip(lines[:40], want["head"])):
        assert a == b, f"line {i}"
    assert lines[-40:] == want["tail"]
    assert hashlib.sha256(("\n".join(lines) + "\n").encode()).hexdigest() == want["sha256"]


@pytest.mark.parametrize("seed,n,flank,kw", [(301, 12, 700, dict(n_frac=0.01)), (302, 40, 90, dict(pad=3)),
                                              (303, 3, 2500, dict()), (304, 25, 40, dict(pad=0))])
def test_emul_twobit_decode(emul, seed, n, flank, kw):
    """twobit_windows_kernel + twobit_boundary_kernel (2bit packed DNA of kentsrc/twoBitNew.c:571-582 ->
    library words, 128 bases per thread with 128-bit loads, window-straddling groups base by base)
    against a numpy decode of the same windows; windows of every length and alignment"""
    import ctypes as C
    fam = synth.make_family(seed, n, flank, 0.15, **kw)
    blob, wins, ns, nl, total = synth.family_twobit_windows(fam)
    blob = np.ascontiguousarray(blob, dtype=np.uint8)
    wins = np.ascontiguousarray(wins, dtype=np.uint64)
    n_chunks = (total + 31) // 32
    lib2 = np.zeros(2 * n_chunks, dtype=np.uint32)
    emul.lib.emul_twobit.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_int, C.c_uint64, C.c_void_p]
    nslow = emul.lib.emul_twobit(blob.ctypes.data, blob.size, wins.ctypes.data, len(wins), total, lib2.ctypes.data)
    assert 0 <= nslow <= len(wins) + 4
    if flank >= 700:
        assert nslow < (total // 128) // 2   # most groups take the 128-bit path
    # expected: the library bytes of the family (A0 C1 G2 T3; N -> code 0 + mask, masks are not this kernel's)
    want = np.zeros(total, dtype=np.uint8)
    for (dst, ln, src_byte, src_base) in wins.tolist():
        idx = src_base + np.arange(ln)
        byte = blob[src_byte + (idx >> 2)]
        t = (byte >> (6 - 2 * (idx & 3)).astype(np.uint8)) & 3          # T0 C1 A2 G3
        want[dst: dst + ln] = np.array([3, 1, 0, 2], dtype=np.uint8)[t]
    i = np.arange(total)
    got = (lib2[i >> 4] >> (2 * (i & 15)).astype(np.uint32)) & 3
    assert np.array_equal(got.astype(np.uint8), want)
    seq = np.asarray(fam.sequence)
    known = seq <= 3
    assert np.array_equal(want[known], seq[known])   # and that is the reference loader's byte per base


@pytest.mark.parametrize("cfg,B,matrix", [
    ((29, 1, 29), 14, "20p43g"), ((15, 2, 29), 14, "20p43g"), ((10, 3, 29), 14, "25p43g"), ((8, 4, 29), 14, "14p43g"),
    ((5, 6, 29), 14, "18p43g"), ((4, 8, 29), 14, "20p43g"), ((27, 3, 81), 40, "14p43g"), ((41, 2, 81), 40, "14p43g"),
    ((17, 5, 81), 40, "18p43g"),
])
def test_emul_fast_rows(emul, oracle, cfg, B, matrix):
    """long flanks inside generous sequence bounds: nearly every row takes the unpredicated path
    (chain-free sweep, local commit + carry fold-in), then the flank ends and the exact path takes
    over; two CTAs, N bases, one-sided cores"""
    fam = synth.make_family(77 + cfg[0], 11, 260, 0.15, pad=3 * B, one_sided_frac=0.2, n_frac=0.01)
    osc = oracle.scoring(matrix)
    p = make_params(bandwidth=B, minimprovement=27, L=1000, when_to_stop=30)
    for d in (1, 0):
        want = oracle.extend(d, fam.cores, fam.sequence, osc, p, trace=True)
        assert want["rows_processed"] > 150
        got = emul.extend(cfg, d, fam.cores, fam.sequence, osc.to_product(), p, threads=64)
        support.compare_extend(want, got, f"cfg={cfg} dir={d}")


@pytest.mark.parametrize("cfg,gaps", [((8, 4, 29), (3, -4)), ((10, 3, 29), (-6, 1)), ((29, 1, 29), (2, 2)),
                                      ((16, 4, 0), (4, -2)), ((15, 2, 29), (0, 0))])
def test_emul_positive_gap_penalties(emul, oracle, cfg, gaps):
    """-gapopen / -gapext accept any integer (ram_extend.c:292-293): with a positive penalty the
    insertion term can exceed the cell left of it, so the chain-free sweep does not apply"""
    fam = synth.make_family(91, 9, 120, 0.2, pad=40, one_sided_frac=0.2)
    osc = oracle.scoring("20p43g")
    osc.gapopen, osc.gapextn = gaps
    p = make_params(bandwidth=14, minimprovement=27, L=90, when_to_stop=30)
    for d in (1, 0):
        want = oracle.extend(d, fam.cores, fam.sequence, osc, p)
        got = emul.extend(cfg, d, fam.cores, fam.sequence, osc.to_product(), p, threads=64)
        support.compare_extend(want, got, f"cfg={cfg} gaps={gaps} dir={d}")


@pytest.mark.parametrize("cfg,scale", [((29, 1, 29), 3000), ((10, 3, 29), 3000), ((15, 2, 29), 1500)])
def test_emul_scores_beyond_int16(emul, oracle, cfg, scale):
    """the sweep reads the candidate scores as int16 (IDP.2A); a matrix beyond that range (the ABI accepts
    |score| * (L + 2B + 3) <= 1.2e7) must take the exact path: scaled 14p43g with entries of several 10 000,
    and a scale whose entries still fit (the packed path with large scores)"""
    fam = synth.make_family(93, 11, 120, 0.15, pad=40, one_sided_frac=0.2, n_frac=0.01)
    osc = oracle.scoring("14p43g")
    for a in range(4):
        for b in range(9):
            osc.matrix[a][b] *= scale
    osc.gapopen *= scale
    osc.gapextn *= scale
    big = max(abs(int(osc.matrix[a][b])) for a in range(4) for b in range(9))
    assert (big > 32767) == (scale == 3000)
    p = make_params(bandwidth=14, minimprovement=27 * scale, L=80, when_to_stop=30)
    for d in (1, 0):
        want = oracle.extend(d, fam.cores, fam.sequence, osc, p)
        assert want["rows_processed"] > 40
        got = emul.extend(cfg, d, fam.cores, fam.sequence, osc.to_product(), p, threads=64)
        support.compare_extend(want, got, f"cfg={cfg} scale={scale} dir={d}")


@pytest.mark.parametrize("cfg,B,threads,n", [((10, 3, 29), 14, 64, 23), ((8, 4, 29), 14, 32, 9), ((29, 1, 29), 14, 64, 70),
                                           ((27, 3, 81), 40, 32, 12), ((16, 4, 0), 20, 64, 17)])
def test_emul_two_passes_in_the_same_threads(emul, oracle, cfg, B, threads, n):
    """extend_kernel<..., NP = 2>: the right and the left pass of a family take turns in the same
    threads (barrier-free column handshake); different numbers of extendable cores per direction,
    passes that stop at different rows, one CTA and several"""
    fam = synth.make_family(131 + n, n, 150, 0.15, pad=2 * B, one_sided_frac=0.35, n_frac=0.01)
    osc = oracle.scoring("14p43g")
    p = make_params(bandwidth=B, minimprovement=30, L=400, when_to_stop=20)
    got = support.emul_extend_pair(emul, cfg, fam.cores, fam.sequence, osc.to_product(), p, threads=threads)
    for d, g in zip((1, 0), got):
        want = oracle.extend(d, fam.cores, fam.sequence, osc, p)
        support.compare_extend(want, g, f"cfg={cfg} dir={d}")


@pytest.mark.parametrize("name,level,cfg", [("config1_default", "-vvvvv", (8, 4, 29)), ("config1_default", "-vvvvv", (29, 1, 29)),
                                             ("satellite", "-vvvv", (10, 3, 29))])
def test_emul_verbose_dumps(emul, oracle, name, level, cfg):
    """VERBOSE >= 10 dumps (ram_extend.c:992-1090): what the kernel reports for EVERY candidate base of every
    row (row best, its column, the gap values behind **OUT_OF_SEQ**, column sums and -- -vvvvv -- every band
    cell) == the values in the reference binary's own -vvvv / -vvvvv stdout (tests/golden/cli_*/verbose*.txt.xz,
    made by tests/golden/make_golden_verbose.py)."""
    import lzma
    import os
    if cfg not in emul.configs():
        pytest.skip("tiling not in the emulator build")
    path = os.path.join(os.path.dirname(__file__), "golden", f"cli_{name}", f"verbose{level}.txt.xz")
    want = support.parse_verbose_dump(lzma.open(path).read().decode())
    case = gio.load_flat({"config1_default": "config1_default", "satellite": "satellite"}[name])
    sc = gio.oracle_scoring_for(case, oracle).to_product()
    p = gio.params_for(case)
    cells = level != "-vvvv"
    W = 2 * p.bandwidth + 1
    # satellite: the left pass alone has every kind of line (capped from row 273, out of sequence from row 480);
    # the emulation of 64 lock-step threads is the slow part of the CPU suite
    for d in ((0,) if name == "satellite" else (1, 0)):
        rows_want = len(want[d][1]) // 4          # rows in the reference's dump (one emulated run per direction)
        got = support.emul_extend_verbose(emul, cfg, d, case["cores"], case["sequence"], sc, p, rows_want, cells=cells,
                                          threads=64)
        assert got["rows_processed"] == rows_want
        gio.check_against_golden(case, d, got, f"{name} dir={d}")
        per, sums = want[d]
        assert len(sums) == 4 * got["rows_processed"]
        n_seen = 0
        for (r, a, n), w in per.items():
            c = got["cand"][r, n, a]
            assert (int(c[0]), int(c[1])) == (w["best"], w["col"]), (d, r, a, n)
            assert (int(c[2] if d else c[3]) < -279000) == w["oos"], (d, r, a, n)
            if cells:
                assert got["cells"][r, n, a, :, 1].tolist() == w["gap"], (d, r, a, n)
                assert got["cells"][r, n, a, :, 0].tolist() == w["sub"], (d, r, a, n)
            n_seen += 1
        ext = int((case["cores"]["right_extendable" if d else "left_extendable"] != 0).sum())
        assert n_seen == 4 * ext * got["rows_processed"]
        for (r, a), s in sums.items():
            assert int(got["sums"][r, a]) == s, (d, r, a)


def _aliasing_windows(n_bases=1000, n_windows=12, seed=5):
    """One unpadded record and many windows that all read the last ~200 bases of the blob (the windows of a
    library may alias the same source bytes): every 128-base group of such a window is a slow group."""
    rng = np.random.default_rng(seed)
    codes = rng.integers(0, 4, n_bases).astype(np.uint8)          # A0 C1 G2 T3
    blob = synth.pack_twobit_record(codes)
    wins, dst = [], 0
    for i in range(n_windows):
        ln = 150 + int(rng.integers(0, 50))
        src = n_bases - ln - (i % 5)
        wins.append((dst, ln, 0, src))
        dst += ln
    want = np.concatenate([codes[s: s + ln] for (_, ln, _, s) in wins])
    return np.ascontiguousarray(blob, dtype=np.uint8), np.array(wins, dtype=np.uint64), dst, want


def test_emul_twobit_windows_aliasing_the_blob_tail(emul):
    """more slow groups than windows (all of them in the last bytes of an unpadded blob): the bounded slow list
    must not overflow and every base must still be decoded"""
    import ctypes as C
    blob, wins, total, want = _aliasing_windows()
    n_chunks = (total + 31) // 32
    lib2 = np.zeros(2 * n_chunks, dtype=np.uint32)
    emul.lib.emul_twobit.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_int, C.c_uint64, C.c_void_p]
    nslow = emul.lib.emul_twobit(blob.ctypes.data, blob.size, wins.ctypes.data, len(wins), total, lib2.ctypes.data)
    assert nslow >= 0
    i = np.arange(total)
    got = (lib2[i >> 4] >> (2 * (i & 15)).astype(np.uint32)) & 3
    assert np.array_equal(got.astype(np.uint8), want)
