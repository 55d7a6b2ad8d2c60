"""GPU parity at the BENCHMARKED sizes: the CUDA path, through the C-ABI, against outputs of the
reference's own extend_alignment() object code on exactly the inputs bench.py times
(tests/golden/full_*.npz, produced by tests/golden/make_golden_full.py; inputs are regenerated from
their seeds and checked against the stored sha256)."""
import numpy as np
import pytest

from repeatafterme_b200 import batch
from repeatafterme_b200.abi import make_params
from tests import golden_io as gio
from tests.cases import full_case_family

pytestmark = pytest.mark.gpu


def _got(res, master, ext_len, score_add):
    return dict(res.as_dict(), master=master, ext_len=ext_len, score_add=score_add)


def _check(want, got, what):
    assert got["extended_bp"] == want["extended_bp"], f"{what}: bp {got['extended_bp']} != {want['extended_bp']}"
    assert got["rows_processed"] == want["rows_processed"], f"{what}: rows"
    n = want["rows_processed"]
    assert np.array_equal(np.asarray(got["master"])[:n], want["master"][:n]), f"{what}: consensus differs"
    assert np.array_equal(got["ext_len"], want["ext_len"]), f"{what}: extension lengths differ"
    assert np.array_equal(got["score_add"], want["score_add"]), f"{what}: scores differ"


def _params(case, **over):
    p = dict(case["params"])
    matrix = p.pop("matrix")
    p.update(over)
    return matrix, make_params(**p)


def test_config4_full_equals_reference(ramcu, monkeypatch):
    """BASELINE configs[3] at full size (10 000 instances, 5 kb flanks, 2 x 5100 columns): the launch of
    both passes that bench.py times (the form the planner picks: two single-pass units sharing the launch),
    both passes in the same threads (RAMCU_PAIR_ALWAYS), the two-call form, and a second tiling == the
    reference's outputs"""
    case = gio.load_full("config4")["seed4"]
    fam = full_case_family("config4", "seed4")
    gio.check_full_inputs(fam, case)
    matrix, p = _params(case)
    sc = ramcu.scoring(matrix)
    lib = ramcu.upload_library(fam.sequence)
    try:
        pair = ramcu.extend_pair(lib, fam.cores, sc, p)
        for d, o in zip((1, 0), pair):
            _check(case["want"][d], _got(*o), f"pair launch dir={d}")
        monkeypatch.setenv("RAMCU_PAIR_ALWAYS", "1")
        pair = ramcu.extend_pair(lib, fam.cores, sc, p)
        monkeypatch.delenv("RAMCU_PAIR_ALWAYS")
        for d, o in zip((1, 0), pair):
            _check(case["want"][d], _got(*o), f"both passes in the same threads dir={d}")
        for G in (0, 5):
            _, pg = _params(case, lanes_per_instance=G)
            for d in (1, 0):
                _check(case["want"][d], _got(*ramcu.extend(lib, d, fam.cores, sc, pg)), f"G={G} dir={d}")
    finally:
        ramcu.free_library(lib)


def test_config5_sample_full_size_through_batch_run(ramcu):
    """22 families of the configs[4] mix at FULL size (every bandwidth, N up to 8286) in ONE
    ramcu_batch_run (all 44 column loops share the device) == the reference, family by family"""
    cases = gio.load_full("config5_sample")
    keys = sorted(cases, key=int)
    fams = [full_case_family("config5_sample", k) for k in keys]
    for k, f in zip(keys, fams):
        gio.check_full_inputs(f, cases[k])
    res = batch.extend_families(ramcu, fams)
    for k, r in zip(keys, res):
        _check(cases[k]["want"][1], r.right, f"family {k} right")
        _check(cases[k]["want"][0], r.left, f"family {k} left")


@pytest.mark.parametrize("key", ["b500", "b511", "b100_many_ctas"])
def test_wide_bands_equal_reference(ramcu, key):
    """bandwidth 500 (util/davidExtendConsRAM.pl:248), 511, and a family on far more than 40 CTAs
    at bandwidth 100 (RED column exchange)"""
    case = gio.load_full("wide_bands")[key]
    fam = full_case_family("wide_bands", key)
    gio.check_full_inputs(fam, case)
    matrix, p = _params(case)
    sc = ramcu.scoring(matrix)
    lib = ramcu.upload_library(fam.sequence)
    try:
        for d in (1, 0):
            _check(case["want"][d], _got(*ramcu.extend(lib, d, fam.cores, sc, p)), f"{key} dir={d}")
        pair = ramcu.extend_pair(lib, fam.cores, sc, p)
        for d, o in zip((1, 0), pair):
            _check(case["want"][d], _got(*o), f"{key} pair dir={d}")
    finally:
        ramcu.free_library(lib)


def test_bandwidth_1000_vs_oracle(ramcu, oracle):
    """the widest band the register tilings take (2 x 1023 + 1 cells on 32 lanes x 64 cells): bandwidth 1000
    and 1023 against the oracle"""
    from repeatafterme_b200 import synth
    from tests import support
    fam = synth.make_family(71, 24, 350, 0.15, pad=30, one_sided_frac=0.2)
    osc = oracle.scoring("18p43g")
    lib = ramcu.upload_library(fam.sequence)
    try:
        for B in (1000, 1023):
            p = make_params(bandwidth=B, minimprovement=27, L=20000)
            for d in (1, 0):
                want = oracle.extend(d, fam.cores, fam.sequence, osc, p)
                got = _got(*ramcu.extend(lib, d, fam.cores, osc.to_product(), p))
                support.compare_extend(want, got, f"B={B} dir={d}")
    finally:
        ramcu.free_library(lib)
