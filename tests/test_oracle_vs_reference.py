"""The oracle against the reference's own object code (oracle/_ref, built from /root/reference by
oracle/Makefile). Skipped where the reference build is not present."""
import os
import subprocess

import numpy as np
import pytest

from repeatafterme_b200 import synth
from repeatafterme_b200.abi import CORE_DTYPE, make_params
from tests import support


def test_reference_own_test_suite_passes(ref):
    exe = os.path.join(support.ORACLE_DIR, "_ref", "test_suite")
    if not os.path.exists("/root/reference/test/extension-test2.2bit"):
        pytest.skip("reference fixtures not present")
    out = subprocess.run([exe], cwd="/root/reference", capture_output=True, text=True).stdout
    assert "3 tests, 650 assertions, 0 failures" in out
    assert "1 tests, 2 assertions, 0 failures" in out


@pytest.mark.parametrize("name", support.MATRICES + ("repeatscout",))
def test_scoring_tables(oracle, ref, name):
    a, b = oracle.scoring(name, (2, -3, -4)), ref.scoring(name, (2, -3, -4))
    assert np.array_equal(a.m49(), b.m49()) and (a.gapopen, a.gapextn) == (b.gapopen, b.gapextn)


def test_product_tables_match_reference(ref):
    """the tables baked into libramcu.so (ramcu_scoring_by_name) == score_system.c"""
    from repeatafterme_b200.abi import load_ramcu, scoring_by_name
    lib = load_ramcu()
    for name in support.MATRICES + ("repeatscout",):
        want = ref.scoring(name, (2, -3, -4)).to_product()
        got = scoring_by_name(lib, name, (2, -3, -4))
        assert np.array_equal(got.as_array(), want.as_array())
        assert (got.gapopen, got.gapextn) == (want.gapopen, want.gapextn)


def test_rows_random(oracle, ref):
    """compute_nw_row, cell by cell: random bandwidths, scoring systems, directions, strands,
    N bases and bounds (including exhausted flanks and bands hanging over both ends)."""
    rng = np.random.default_rng(2024)
    cells = 0
    for trial in range(400):
        B = int(rng.integers(1, 61))
        W = 2 * B + 1
        name = (support.MATRICES + ("repeatscout",))[trial % 5]
        sc = oracle.scoring(name, (1, -1, -5))
        n = int(rng.integers(30, 400))
        seq = rng.integers(0, 4, n, dtype=np.uint8)
        seq[rng.random(n) < 0.05] = 99
        core = np.zeros(1, dtype=CORE_DTYPE)
        a = int(rng.integers(0, n - 4))
        b = int(rng.integers(a, min(n - 1, a + 20)))
        orient = int(rng.integers(0, 2))
        core["orient"] = orient
        core["left_seq_pos"], core["right_seq_pos"] = (b, a) if orient else (a, b)
        lo = int(rng.integers(0, a + 1))
        hi = int(rng.integers(b, n))
        core["lower_seq_bound"], core["upper_seq_bound"] = lo, hi
        direction = int(rng.integers(0, 2))
        ps = np.array([0 if k == B else abs(k - B) * sc.gapextn + sc.gapopen for k in range(W)], dtype=np.int32)
        pg = ps.copy()
        qs, qg = ps.copy(), pg.copy()
        for row in range(int(rng.integers(5, 80))):
            cb = int(rng.integers(0, 4))
            b1, i1, cs1, cg1 = oracle.nw_row(direction, row, cb, core, ps, pg, seq, sc, B)
            b2, i2, cs2, cg2 = ref.nw_row(direction, row, cb, core, qs, qg, seq, sc, B)
            assert (b1, i1) == (b2, i2), (trial, row)
            assert np.array_equal(cs1, cs2) and np.array_equal(cg1, cg2), (trial, row)
            ps, pg, qs, qg = cs1, cg1, cs2, cg2
            cells += W
    assert cells > 500000


@pytest.mark.parametrize("seed,B,matrix,kw", [
    (1, 14, "20p43g", dict()),
    (2, 40, "14p43g", dict(one_sided_frac=0.3, n_frac=0.02)),
    (3, 5, "18p43g", dict(pad=0)),
    (4, 100, "25p43g", dict(one_sided_frac=0.1)),
    (5, 33, "14p43g", dict(pad=3, n_frac=0.1)),
])
def test_extend_random_families(oracle, ref, seed, B, matrix, kw):
    fam = synth.make_family(seed, 25, 180, 0.15, **kw)
    sc = oracle.scoring(matrix)
    for L in (20000, 70):
        p = make_params(bandwidth=B, minimprovement=27, L=L, when_to_stop=40)
        for d in (1, 0):
            a = oracle.extend(d, fam.cores, fam.sequence, sc, p)
            b = ref.extend(d, fam.cores, fam.sequence, sc, p)
            assert a["extended_bp"] == b["extended_bp"]
            n = a["rows_processed"]
            assert np.array_equal(a["master"][:n], b["master"][:n])
            assert np.array_equal(a["ext_len"], b["ext_len"]) and np.array_equal(a["score_add"], b["score_add"])


def test_synthetic_flat_form_equals_reference_loader(ref, tmp_path):
    """synth.make_family builds the post-loader form directly; the reference's own
    loadSequenceSubsetMinimal() on the .2bit/.tsv written from the same family must agree."""
    fam = synth.make_family(9, 40, 200, 0.15, one_sided_frac=0.3, n_frac=0.02)
    tb, rg = str(tmp_path / "f.2bit"), str(tmp_path / "f.tsv")
    synth.write_family_files(fam, tb, rg)
    seq, cores, offsets, bounds = ref.load(tb, rg, 20040)
    assert np.array_equal(seq, fam.sequence)
    for f in ("left_seq_pos", "right_seq_pos", "lower_seq_bound", "upper_seq_bound", "orient",
              "left_extendable", "right_extendable"):
        assert np.array_equal(cores[f], fam.cores[f]), f
