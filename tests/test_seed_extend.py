"""Seed extension -- cons_seed_extend() / pair_seed_extend() / global_align() of the reference
(bnw_extend.c:167-649), the other callers of compute_nw_row() -- on the extension kernel.
Fixtures (tests/golden/seed_vectors.json, tests/golden/make_golden_seed.py): the reference's own unit
test data with the answers it asserts (1573 / 67-267 / 17-214, 743, 1462; global_align 2732) and seeded
random cases (N bases in query and subject, both strands, y-drop), every value from the reference's
object code. CPU: the kernel source under the SIMT emulator, row by row. GPU: ramcu_seed_extend_batch /
ramcu_global_align through the C-ABI."""
import ctypes as C
import json
import math
import os

import numpy as np
import pytest

from repeatafterme_b200.abi import CORE_DTYPE, Scoring
from tests import support

VEC = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "seed_vectors.json")))


def complexity_adjust(score, comp, lam, freqs):
    """Phil Green's adjustment as the reference applies it (bnw_extend.c:1137-1165); host-side float math"""
    t_factor = t_sum = t_counts = 0.0
    for i in range(4):
        if comp[i]:
            t_factor += comp[i] * math.log(comp[i])
            t_sum += comp[i] * math.log(freqs[i])
            t_counts += comp[i]
    if t_counts:
        t_factor -= t_counts * math.log(t_counts)
    t_sum -= t_factor
    return max(int(score + t_sum / lam + .999), 0)


def _scan(rows, y_drop, right):
    best, sidx, cidx = 0, -1, -1
    for r, (sc, pos) in enumerate(rows):
        if y_drop >= 0 and best - sc > (180 if right else y_drop):
            break
        if sc > best:
            best, sidx, cidx = sc, pos, r
    return best, sidx, cidx


def _passes(case):
    """the two extension passes of one cons_seed_extend() call as (direction, core, forced consensus)"""
    q = np.array(case["query"], dtype=np.uint8)
    s0, s1, orient = case["subj_seed_start"], case["subj_seed_end"], case["orient"]
    bounds = [0] + case["boundaries"]
    k = next(i for i in range(len(case["boundaries"])) if case["boundaries"][i] > s0)
    lo, hi = bounds[k], case["boundaries"][k] - 1
    out = []
    for side in (0, 1):
        rows = (len(q) - (case["query_seed_end"] + 1)) if side else case["query_seed_start"]
        forced = np.array([q[case["query_seed_end"] + r + 1] if side else q[case["query_seed_start"] - r - 1] for r in range(rows)],
                          dtype=np.uint8)
        forced = np.where(forced > 3, 4, forced).astype(np.uint8)
        core = np.zeros(1, dtype=CORE_DTYPE)
        core["orient"], core["left_extendable"], core["right_extendable"] = orient, 1, 1
        core["left_seq_pos"], core["right_seq_pos"] = (s1, s0) if orient else (s0, s1)
        low_side = (side == 0) != bool(orient)
        core["lower_seq_bound"], core["upper_seq_bound"] = (lo, s0 - 1) if low_side else (s1 + 1, hi)
        out.append((side, core, forced, rows))
    return out, lo, hi


@pytest.mark.parametrize("idx", [0, 2, 3, 5, 10, 14, 21, 35])
def test_emul_seed_rows_reproduce_reference(emul, oracle, idx):
    """kernel source on the CPU: per-row values of both passes -> the reference's scan -> its answers"""
    case = (VEC["unit"] + VEC["random"])[idx]
    emul.lib.emul_seed_rows.argtypes = [C.c_int32] * 3 + [C.c_void_p, C.c_uint64, C.c_int32, C.c_void_p, C.POINTER(Scoring),
                                                        C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]
    B = case["bandwidth"]
    cfg = {5: (3, 4, 11), 14: (10, 3, 29), 20: (16, 4, 0), 40: (27, 3, 81)}[B]
    sc = oracle.scoring(case["matrix"]).to_product()
    n_row = np.array(case["want"]["n_row"], dtype=np.int32)
    seq = np.array(case["sequence"], dtype=np.uint8)
    passes, lo, hi = _passes(case)
    best, sidx, cidx = {}, {}, {}
    for side, core, forced, rows in passes:
        best[side], sidx[side], cidx[side] = 0, -1, -1
        if rows <= 0:
            continue
        out = np.zeros((rows, 2), dtype=np.int32)
        rc = emul.lib.emul_seed_rows(cfg[0], cfg[1], cfg[2], seq.ctypes.data, seq.size, side, core.ctypes.data, C.byref(sc),
                                     n_row.ctypes.data, B, rows, forced.ctypes.data, out.ctypes.data)
        assert rc == 0
        best[side], sidx[side], cidx[side] = _scan(out.tolist(), case["y_drop"], side == 1)
    w = case["want"]
    assert best[0] + best[1] == w["unadjusted"]
    assert case["query_seed_start"] - (cidx[0] + 1) == w["query_start"]
    assert case["query_seed_end"] + (cidx[1] + 1) == w["query_end"]
    o = case["orient"]
    assert case["subj_seed_start"] - ((sidx[1] if o else sidx[0]) + 1) == w["subj_start"]
    assert case["subj_seed_end"] + ((sidx[0] if o else sidx[1]) + 1) == w["subj_end"]


@pytest.mark.gpu
def test_seed_extend_batch_equals_reference(ramcu, oracle):
    """all fixtures of one scoring system in ONE ramcu_seed_extend_batch call each"""
    cases = VEC["unit"] + VEC["random"]
    for matrix in sorted({c["matrix"] for c in cases}):
        for B in sorted({c["bandwidth"] for c in cases if c["matrix"] == matrix}):
            for y_drop in sorted({c["y_drop"] for c in cases if c["matrix"] == matrix and c["bandwidth"] == B}):
                group = [c for c in cases if c["matrix"] == matrix and c["bandwidth"] == B and c["y_drop"] == y_drop]
                # every case brings its own library: concatenate them, shifting the coordinates
                seq, seeds, off = [], [], 0
                for c in group:
                    _, lo, hi = _passes(c)
                    seeds.append(dict(query=np.array(c["query"], dtype=np.uint8), query_seed_start=c["query_seed_start"],
                                      query_seed_end=c["query_seed_end"], orient=c["orient"],
                                      subj_seed_start=c["subj_seed_start"] + off, subj_seed_end=c["subj_seed_end"] + off,
                                      lower_seq_bound=lo + off, upper_seq_bound=hi + off))
                    seq.append(np.array(c["sequence"], dtype=np.uint8))
                    off += len(c["sequence"])
                lib = ramcu.upload_library(np.concatenate(seq))
                try:
                    sc = oracle.scoring(matrix).to_product()
                    res = ramcu.seed_extend_batch(lib, seeds, sc, group[0]["want"]["n_row"], B, y_drop)
                finally:
                    ramcu.free_library(lib)
                off = 0
                for c, r in zip(group, res):
                    w = c["want"]
                    assert r.left_score + r.right_score == w["unadjusted"], (matrix, B)
                    assert (r.query_start, r.query_end) == (w["query_start"], w["query_end"])
                    assert (r.subj_start - off, r.subj_end - off) == (w["subj_start"], w["subj_end"])
                    off += len(c["sequence"])


@pytest.mark.gpu
def test_reference_unit_test_answers(ramcu, oracle):
    """the numbers the reference's own tests assert (bnw_extend.c:1235, 1344-1376), complexity adjustment
    (host float math, bnw_extend.c:1137-1165) included"""
    lam_freqs = {"20p43g": None}
    for c, (score, ss, se, qs, qe) in zip(VEC["unit"], [(1573, 67, 267, 17, 214), (743, 400, 498, 0, 99), (1462, 803, 1014, 0, 212)]):
        seq = np.array(c["sequence"], dtype=np.uint8)
        _, lo, hi = _passes(c)
        lib = ramcu.upload_library(seq)
        try:
            sc = oracle.scoring(c["matrix"]).to_product()
            r = ramcu.seed_extend_batch(lib, [dict(query=np.array(c["query"], dtype=np.uint8), query_seed_start=c["query_seed_start"],
                                                   query_seed_end=c["query_seed_end"], orient=c["orient"],
                                                   subj_seed_start=c["subj_seed_start"], subj_seed_end=c["subj_seed_end"],
                                                   lower_seq_bound=lo, upper_seq_bound=hi)], sc, c["want"]["n_row"],
                                        c["bandwidth"], c["y_drop"])[0]
        finally:
            ramcu.free_library(lib)
        assert (r.subj_start, r.subj_end, r.query_start, r.query_end) == (ss, se, qs, qe)
        raw = r.left_score + r.right_score
        if c["adjust"]:
            comp = [int((seq[r.subj_start: r.subj_end + 1] == b).sum()) for b in range(4)]
            # 20p43g: lambda from the reference's own calculateLambda, background 0.285/0.215/0.215/0.285
            lam = _lambda_20p43g(oracle)
            raw = complexity_adjust(raw, comp, lam, [0.285, 0.215, 0.215, 0.285])
        assert raw == score
    for g in VEC["global"]:
        seq = np.array(g["sequence"], dtype=np.uint8)
        lib = ramcu.upload_library(seq)
        try:
            sc = oracle.scoring(g["matrix"]).to_product()
            got = ramcu.global_align(lib, g["orient"], seq[g["seq1_start"]: g["seq1_end"] + 1], g["seq2_start"], g["seq2_end"],
                                     sc, [-1, -1, -1, -1, -1], g["bandwidth"])
        finally:
            ramcu.free_library(lib)
        assert got[0] == g["want"]


def _lambda_20p43g(oracle):
    """lambda of a scoring matrix: the positive root of sum_ij p_i p_j exp(lambda s_ij) = 1 (what the
    reference's calculateLambda solves, score_system.c:23-89), by bisection"""
    m = oracle.scoring("20p43g").m49()[:, :4].astype(float)
    p = np.array([0.285, 0.215, 0.215, 0.285])
    f = lambda lam: float((np.outer(p, p) * np.exp(lam * m)).sum()) - 1.0
    lo, hi = 1e-6, 1.0
    for _ in range(200):
        mid = 0.5 * (lo + hi)
        if f(mid) > 0:
            hi = mid
        else:
            lo = mid
    return 0.5 * (lo + hi)
