"""Seed-extension fixtures (run in the build container): the data of the reference's own unit tests
test_cons_seed_extend / test_global_align (bnw_extend.c:1139-1383: sequences, seeds, asserted answers)
plus seeded random cases, every expected value produced by the reference's object code
(oracle/_ref/libramref.so) -> tests/golden/seed_vectors.json."""
import json
import os
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from repeatafterme_b200 import synth  # noqa: E402
from tests import support  # noqa: E402

CODE = {c: i for i, c in enumerate("ACGT")}


def enc(s):
    return [CODE.get(c, 99) for c in s]


def c_string(src, name):
    """concatenated string literals of `char name[] = "..." "...";` in C source text"""
    m = re.search(r"char\s+" + name + r"\[\]\s*=(.*?);", src, re.S)
    body = re.sub(r"//[^\n]*", "", m.group(1))
    return "".join(re.findall(r'"([^"]*)"', body))


def main():
    ref = support.Reference.try_load()
    src = open("/root/reference/bnw_extend.c").read()
    t_seed = src[src.index("MU_TEST(test_cons_seed_extend)"):]
    t_glob = src[src.index("MU_TEST(test_global_align)"):src.index("MU_TEST(test_cons_seed_extend)")]
    cons, alu = c_string(t_seed, "cons_seq"), c_string(t_seed, "alu_seqs")
    out = {"unit": [], "random": [], "global": []}
    # (subject seed start/end, orient, adjust, asserted by the reference's test)
    for (s0, s1, orient, adjust, asserted) in [(100, 110, 0, 0, dict(score=1573, subj_start=67, subj_end=267, query_start=17, query_end=214)),
                                               (449, 459, 0, 1, dict(score=743, subj_start=400, subj_end=498, query_start=0, query_end=99)),
                                               (956, 966, 1, 1, dict(score=1462, subj_start=803, subj_end=1014, query_start=0, query_end=212))]:
        r = ref.cons_seed_extend("20p43g", enc(cons), 50, 60, enc(alu), [1050], s0, s1, orient, 5, adjust, -1)
        for k, v in asserted.items():
            assert r[k] == v, (k, r[k], v)
        out["unit"].append(dict(matrix="20p43g", query=enc(cons), query_seed_start=50, query_seed_end=60, sequence=enc(alu),
                                boundaries=[1050], subj_seed_start=s0, subj_seed_end=s1, orient=orient, bandwidth=5,
                                adjust=adjust, y_drop=-1, want={k: (v.tolist() if hasattr(v, "tolist") else v) for k, v in r.items()}))
    galu = c_string(t_glob, "alu_seqs")
    g = ref.global_align("20p43g", 0, 50, 367, 500, 815, enc(galu), 5)
    assert g == 2732
    out["global"].append(dict(matrix="20p43g", sequence=enc(galu), orient=0, seq1_start=50, seq1_end=367, seq2_start=500,
                              seq2_end=815, bandwidth=5, want=g))
    rng = np.random.default_rng(12)
    for t in range(40):
        matrix = ["14p43g", "18p43g", "20p43g", "25p43g"][t % 4]
        B = int(rng.choice([5, 14, 20, 40]))
        qlen = int(rng.integers(60, 500))
        anc = rng.integers(0, 4, qlen, dtype=np.uint8)
        q = anc.copy()
        if t % 5 == 0:
            q[rng.integers(0, qlen, 3)] = 99          # N bases in the query
        seeds_at = int(rng.integers(12, qlen - 24))
        l = int(rng.integers(8, 12))
        copy = synth.mutate(rng, anc, 0.12)
        # keep the seed itself exact inside the copy: rebuild around it
        left, right = synth.mutate(rng, anc[:seeds_at], 0.12), synth.mutate(rng, anc[seeds_at + l:], 0.12)
        copy = np.concatenate([left, anc[seeds_at:seeds_at + l], right])
        s_at = left.size
        orient = int(t % 3 == 0)
        pre, post = rng.integers(0, 4, int(rng.integers(0, 80)), dtype=np.uint8), rng.integers(0, 4, int(rng.integers(0, 80)), dtype=np.uint8)
        other = rng.integers(0, 4, 150, dtype=np.uint8)
        if orient:
            copy = synth.revcomp(copy)
            s_at = copy.size - (left.size + l)
        if t % 7 == 0:
            copy = copy.copy()
            copy[rng.integers(0, copy.size, 4)] = 99   # N bases in the subject
            copy[s_at:s_at + l] = (synth.revcomp(anc[seeds_at:seeds_at + l]) if orient else anc[seeds_at:seeds_at + l])
        seq = np.concatenate([other, pre, copy, post, other])
        bounds = [other.size, other.size + pre.size + copy.size + post.size, seq.size]
        s0 = other.size + pre.size + s_at
        y_drop = int(rng.choice([-1, -1, 60, 150]))
        adjust = int(t % 2)
        r = ref.cons_seed_extend(matrix, q, seeds_at, seeds_at + l - 1, seq, bounds, s0, s0 + l - 1, orient, B, adjust, y_drop)
        out["random"].append(dict(matrix=matrix, query=q.tolist(), query_seed_start=seeds_at, query_seed_end=seeds_at + l - 1,
                                  sequence=seq.tolist(), boundaries=bounds, subj_seed_start=s0, subj_seed_end=s0 + l - 1,
                                  orient=orient, bandwidth=B, adjust=adjust, y_drop=y_drop,
                                  want={k: (v.tolist() if hasattr(v, "tolist") else v) for k, v in r.items()}))
    for t in range(6):
        a = rng.integers(0, 4, int(rng.integers(80, 300)), dtype=np.uint8)
        b = synth.mutate(rng, a, 0.1)
        gap = rng.integers(0, 4, 40, dtype=np.uint8)
        seq = np.concatenate([gap, a, gap, b, gap])
        s1, s2 = gap.size, 2 * gap.size + a.size
        B = int(rng.choice([5, 14, 30]))
        g = ref.global_align("14p43g", 0, s1, s1 + a.size - 1, s2, s2 + b.size - 1, seq, B)
        out["global"].append(dict(matrix="14p43g", sequence=seq.tolist(), orient=0, seq1_start=s1, seq1_end=s1 + a.size - 1,
                                  seq2_start=s2, seq2_end=s2 + b.size - 1, bandwidth=B, want=int(g)))
    json.dump(out, open(os.path.join(HERE, "seed_vectors.json"), "w"))
    print("unit", [u["want"]["score"] for u in out["unit"]], "random", len(out["random"]), "global", [g["want"] for g in out["global"]])


if __name__ == "__main__":
    main()
