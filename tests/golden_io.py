"""Readers for the committed golden fixtures (tests/golden/)."""
from __future__ import annotations

import glob
import json
import os

import numpy as np

from repeatafterme_b200.abi import make_params

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def flat_case_names():
    return sorted(os.path.basename(p)[5:-4] for p in glob.glob(os.path.join(GOLDEN, "flat_*.npz")))


def cli_case_names():
    """CLI cases whose inputs are committed (the ce10 stand-in cases regenerate theirs)"""
    return sorted(os.path.basename(p)[4:] for p in glob.glob(os.path.join(GOLDEN, "cli_*"))
                  if os.path.exists(os.path.join(p, "input.npz")))


def load_cli_outputs(name):
    d = os.path.join(GOLDEN, f"cli_{name}")
    return {f: open(os.path.join(d, f)).read() for f in ("stdout.txt", "out.tsv", "out.fa", "cons.fa")}


def load_flat(name):
    z = np.load(os.path.join(GOLDEN, f"flat_{name}.npz"), allow_pickle=False)
    meta = json.loads(str(z["meta"]))
    want = {}
    for d in (1, 0):
        want[d] = dict(extended_bp=int(z[f"bp{d}"]), master=z[f"master{d}"], ext_len=z[f"ext_len{d}"],
                       score_add=z[f"score_add{d}"])
    return dict(name=name, sequence=z["sequence"], cores=z["cores"], matrix=meta["matrix"],
                rs=tuple(meta["rs"]), gap_override=meta["gap_override"], params=meta["params"], want=want)


def oracle_scoring_for(case, checker):
    """scoring system of a flat case from `checker` (Oracle or Reference)"""
    sc = checker.scoring(case["matrix"], case["rs"])
    if case["gap_override"]:
        sc.gapopen, sc.gapextn = case["gap_override"]
    return sc


def params_for(case, **over):
    p = dict(case["params"])
    p.update(over)
    return make_params(**p)


def check_against_golden(case, direction, got, what=""):
    """`got` = dict(extended_bp, rows_processed, master, ext_len, score_add) vs reference output.
    The reference returns master[] only as far as it wrote it; compare the processed rows."""
    w = case["want"][direction]
    assert got["extended_bp"] == w["extended_bp"], f"{what}: bp {got['extended_bp']} != {w['extended_bp']}"
    n = got["rows_processed"]
    assert np.array_equal(np.asarray(got["master"])[:n], w["master"][:n]), f"{what}: consensus differs"
    assert np.array_equal(got["ext_len"], w["ext_len"]), f"{what}: extension lengths differ"
    assert np.array_equal(got["score_add"], w["score_add"]), f"{what}: scores differ"


def load_unit():
    return json.load(open(os.path.join(GOLDEN, "unit_vectors.json")))


def load_cli(name):
    d = os.path.join(GOLDEN, f"cli_{name}")
    z = np.load(os.path.join(d, "input.npz"), allow_pickle=False)
    meta = json.loads(str(z["meta"]))
    seqs, off = [], 0
    for ln in z["lengths"]:
        seqs.append(z["seq"][off: off + int(ln)])
        off += int(ln)
    want = {f: open(os.path.join(d, f)).read() for f in ("stdout.txt", "out.tsv", "out.fa", "cons.fa")}
    return dict(names=meta["names"], seqs=seqs, bed=[tuple(b) for b in meta["bed"]], flags=meta["flags"],
                want=want)


def load_full(group):
    """reference outputs at the benchmarked sizes (tests/golden/make_golden_full.py): inputs are
    regenerated from seeds by tests.cases.full_case_family and checked against the stored sha256.
    -> {key: dict(n, params, seq_sha, cores_sha, want={1: {...}, 0: {...}})}"""
    z = np.load(os.path.join(GOLDEN, f"full_{group}.npz"), allow_pickle=False)
    meta = json.loads(str(z["meta"]))
    out = {}
    for k, m in meta.items():
        want = {}
        for d in (1, 0):
            want[d] = dict(extended_bp=int(m[f"bp{d}"]), rows_processed=int(m[f"rows{d}"]),
                           master=z[f"{k}_master{d}"], ext_len=z[f"{k}_ext_len{d}"],
                           score_add=z[f"{k}_score_add{d}"])
        out[k] = dict(n=m["n"], params=m["params"], seq_sha=m["seq_sha"], cores_sha=m["cores_sha"], want=want)
    return out


def check_full_inputs(fam, case):
    """the regenerated inputs are the ones the reference saw"""
    import hashlib
    assert hashlib.sha256(np.ascontiguousarray(fam.sequence).tobytes()).hexdigest() == case["seq_sha"], "generator drifted (sequence)"
    assert hashlib.sha256(np.ascontiguousarray(fam.cores).tobytes()).hexdigest() == case["cores_sha"], "generator drifted (cores)"
