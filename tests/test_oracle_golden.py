"""The oracle (oracle/ram_oracle.c) against the committed golden vectors: outputs of the
reference's own object code and the known-answer rows of the reference's own unit test."""
import numpy as np
import pytest

from repeatafterme_b200.abi import CORE_DTYPE
from tests import golden_io as gio


@pytest.mark.parametrize("name", gio.flat_case_names())
def test_oracle_matches_reference_outputs(oracle, name):
    case = gio.load_flat(name)
    sc = gio.oracle_scoring_for(case, oracle)
    p = gio.params_for(case)
    for d in (1, 0):
        got = oracle.extend(d, case["cores"], case["sequence"], sc, p)
        gio.check_against_golden(case, d, got, f"{name} dir={d}")


def _unit_core(u, side, n):
    l, r, lo, hi, orient = u[f"{side}_cores"][n]
    core = np.zeros(1, dtype=CORE_DTYPE)
    core["left_seq_pos"], core["right_seq_pos"] = l, r
    core["lower_seq_bound"], core["upper_seq_bound"], core["orient"] = lo, hi, orient
    core["left_extendable"] = core["right_extendable"] = 1
    return core


def test_unit_vectors_compute_nw_row(oracle):
    """bnw_extend.c:1384-1651 replayed on the oracle: every row of both reciprocal libraries equals
    what the reference's compute_nw_row produced, left/right reciprocity holds cell by cell
    (bnw_extend.c:1617-1624) and row 6 of core 0 is the known answer (bnw_extend.c:1566-1567)."""
    u = gio.load_unit()
    sc = oracle.scoring(u["matrix"])
    B, W = u["bandwidth"], 2 * u["bandwidth"] + 1
    rows = {}
    for side, direction in (("left", 0), ("right", 1)):
        seq = np.array(u[f"p_{side}"], dtype=np.uint8)
        for n in range(3):
            core = _unit_core(u, side, n)
            ps = np.array([0 if k == B else abs(k - B) * sc.gapextn + sc.gapopen for k in range(W)], dtype=np.int32)
            pg = ps.copy()
            for row, cb in enumerate(u["cons_base"]):
                best, idx, cs, cg = oracle.nw_row(direction, row, cb, core, ps, pg, seq, sc, B)
                want = u["reference_rows"][side][n][row]
                assert best == want["best"] and idx == want["idx"]
                assert cs.tolist() == want["sub"] and cg.tolist() == want["gap"]
                rows[(side, n, row)] = (cs, cg)
                ps, pg = cs, cg
    for n in range(3):
        for row in range(len(u["cons_base"])):
            assert np.array_equal(rows[("left", n, row)][0], rows[("right", n, row)][0])
            assert np.array_equal(rows[("left", n, row)][1], rows[("right", n, row)][1])
    assert rows[("right", 0, 6)][1].tolist() == u["final_right_gap"]
    assert rows[("right", 0, 6)][0].tolist() == u["final_right_sub"]


def test_config1_headline_numbers(oracle):
    """config 1 extends right 177 bp, left 131 bp, every core scores 2950."""
    case = gio.load_flat("config1_default")
    sc = gio.oracle_scoring_for(case, oracle)
    p = gio.params_for(case)
    r = oracle.extend(1, case["cores"], case["sequence"], sc, p)
    l = oracle.extend(0, case["cores"], case["sequence"], sc, p)
    assert (r["extended_bp"], l["extended_bp"]) == (177, 131)
    assert set((r["score_add"] + l["score_add"]).tolist()) == {2950}
    assert set(r["ext_len"].tolist()) == {177} and set(l["ext_len"].tolist()) == {131}
    assert r["stop_row"] == 276 and l["stop_row"] == 230


def test_scoring_tables_are_asymmetric_as_in_reference(oracle):
    sc = oracle.scoring("14p43g")
    assert sc.matrix[0][2] == -10 and sc.matrix[2][0] == -7 and (sc.gapopen, sc.gapextn) == (-33, -7)
    rs = oracle.scoring("repeatscout", (1, -1, -5))
    assert (rs.gapopen, rs.gapextn) == (0, -5) and rs.matrix[0][8] == -1 and rs.matrix[3][3] == 1
