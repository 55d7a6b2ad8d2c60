"""Test support: ctypes bindings for the checkers (oracle port, reference objects, kernel
emulator). TEST INFRASTRUCTURE ONLY -- nothing in repeatafterme_b200/ imports this."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from repeatafterme_b200.abi import CORE_DTYPE, ExtendResult, Params, Scoring, make_params

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
EMUL_DIR = os.path.join(ROOT, "tests", "emul")
vp, i32, u64 = C.c_void_p, C.c_int32, C.c_uint64
P = C.POINTER

MATRICES = ("14p43g", "18p43g", "20p43g", "25p43g")


class OracleScoring(C.Structure):
    _fields_ = [("matrix", (C.c_int32 * 9) * 4), ("gapopen", C.c_int32), ("gapextn", C.c_int32)]

    def to_product(self) -> Scoring:
        """fold the oracle's 9 columns (ACGT acgt N) to the product's 5 (ACGT N)"""
        m = [[self.matrix[a][b] for b in (0, 1, 2, 3, 8)] for a in range(4)]
        return Scoring.from_array(m, self.gapopen, self.gapextn)

    def m49(self) -> np.ndarray:
        return np.array([[self.matrix[a][b] for b in range(9)] for a in range(4)], dtype=np.int32)


class OracleResult(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("extended_bp", "rows_processed", "stop_row", "hit_limit",
                                          "max_score", "n_extendable")]


def _build(target_dir, target, cmd):
    path = os.path.join(target_dir, target)
    if not os.path.exists(path):
        subprocess.check_call(cmd, cwd=target_dir)
    return path


class Oracle:
    """oracle/liboracle.so -- this repo's CPU restatement (the checker)."""

    def __init__(self):
        path = _build(ORACLE_DIR, "liboracle.so", ["make", "-s", "oracle"])
        self.lib = lib = C.CDLL(path)
        lib.oracle_get_scoring.argtypes = [C.c_char_p, i32, i32, i32, P(OracleScoring)]
        lib.oracle_nw_row.argtypes = [i32, i32, i32, vp, vp, vp, vp, vp, vp, P(OracleScoring), i32, P(i32)]
        lib.oracle_extend.argtypes = [i32, i32, vp, vp, P(OracleScoring), i32, i32, i32, i32, i32,
                                      vp, vp, vp, P(OracleResult), vp]
        lib.oracle_cell_updates.restype = C.c_double
        lib.oracle_cell_updates.argtypes = [P(OracleResult), i32]

    def scoring(self, name, rs=(1, -1, -5)) -> OracleScoring:
        s = OracleScoring()
        assert self.lib.oracle_get_scoring(name.encode(), rs[0], rs[1], rs[2], C.byref(s)) == 0
        return s

    def nw_row(self, direction, row_idx, cons_base, core, prev_sub, prev_gap, seq, sc, B):
        core = np.ascontiguousarray(core, dtype=CORE_DTYPE).reshape(1)
        ps = np.ascontiguousarray(prev_sub, dtype=np.int32)
        pg = np.ascontiguousarray(prev_gap, dtype=np.int32)
        cs, cg = np.zeros_like(ps), np.zeros_like(pg)
        idx = i32(0)
        seq = np.ascontiguousarray(seq, dtype=np.uint8)
        best = self.lib.oracle_nw_row(direction, row_idx, cons_base, core.ctypes.data, ps.ctypes.data,
                                      pg.ctypes.data, cs.ctypes.data, cg.ctypes.data,
                                      seq.ctypes.data, C.byref(sc), B, C.byref(idx))
        return best, idx.value, cs, cg

    def extend(self, direction, cores, seq, sc: OracleScoring, p: Params, trace=False):
        cores = np.ascontiguousarray(cores, dtype=CORE_DTYPE)
        seq = np.ascontiguousarray(seq, dtype=np.uint8)
        n = cores.size
        master = np.zeros(max(p.L, 1), dtype=np.uint8)
        ext_len = np.zeros(n, dtype=np.int32)
        score_add = np.zeros(n, dtype=np.int32)
        res = OracleResult()
        tr = np.zeros((max(p.L, 1), 6), dtype=np.int32) if trace else None
        self.lib.oracle_extend(direction, n, cores.ctypes.data, seq.ctypes.data, C.byref(sc),
                               p.bandwidth, p.cappenalty, p.minimprovement, p.L, p.when_to_stop,
                               master.ctypes.data, ext_len.ctypes.data, score_add.ctypes.data,
                               C.byref(res), tr.ctypes.data if trace else None)
        out = dict(extended_bp=res.extended_bp, rows_processed=res.rows_processed,
                   stop_row=res.stop_row, hit_limit=res.hit_limit, max_score=res.max_score,
                   n_extendable=res.n_extendable, master=master[: res.rows_processed].copy(),
                   ext_len=ext_len, score_add=score_add)
        if trace:
            out["trace"] = tr[: res.rows_processed].copy()
        return out


class Reference:
    """oracle/_ref/libramref.so -- the reference's own object code behind flat entry points."""

    @staticmethod
    def try_load():
        path = os.path.join(ORACLE_DIR, "_ref", "libramref.so")
        if not os.path.exists(path):
            if os.path.exists("/root/reference/ram_extend.c"):
                subprocess.check_call(["make", "-s", "ref"], cwd=ORACLE_DIR)
            else:
                return None
        return Reference(path)

    def __init__(self, path):
        self.lib = lib = C.CDLL(path)
        lib.ramref_get_scoring.argtypes = [C.c_char_p, i32, i32, i32, vp, P(i32), P(i32)]
        lib.ramref_extend.argtypes = [i32, i32, vp, vp, u64, vp, i32, i32, i32, i32, i32, i32, i32, i32,
                                      vp, vp, vp, P(C.c_double)]
        lib.ramref_nw_row.argtypes = [i32, i32, i32, vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, P(i32)]
        i64 = C.c_int64
        lib.ramref_cons_seed_extend.argtypes = [C.c_char_p, vp, i32, i32, i32, vp, u64, vp, u64, u64, i32, i32, i32, i32,
                                                vp, vp]
        lib.ramref_global_align.argtypes = [C.c_char_p, i32, u64, u64, u64, u64, vp, u64, i32]
        lib.ramref_load.argtypes = [C.c_char_p, C.c_char_p, i32, P(i32), P(u64), vp, vp, vp, vp]

    def scoring(self, name, rs=(1, -1, -5)) -> OracleScoring:
        m = np.zeros((4, 9), dtype=np.int32)
        go, ge = i32(), i32()
        assert self.lib.ramref_get_scoring(name.encode(), rs[0], rs[1], rs[2], m.ctypes.data,
                                           C.byref(go), C.byref(ge)) == 0
        s = OracleScoring()
        for a in range(4):
            for b in range(9):
                s.matrix[a][b] = int(m[a, b])
        s.gapopen, s.gapextn = go.value, ge.value
        return s

    def nw_row(self, direction, row_idx, cons_base, core, prev_sub, prev_gap, seq, sc, B):
        core = np.ascontiguousarray(core, dtype=CORE_DTYPE).reshape(1)
        ps = np.ascontiguousarray(prev_sub, dtype=np.int32)
        pg = np.ascontiguousarray(prev_gap, dtype=np.int32)
        cs, cg = np.zeros_like(ps), np.zeros_like(pg)
        idx = i32(0)
        seq = np.ascontiguousarray(seq, dtype=np.uint8)
        m = sc.m49()
        best = self.lib.ramref_nw_row(direction, row_idx, cons_base, core.ctypes.data, ps.ctypes.data,
                                      pg.ctypes.data, cs.ctypes.data, cg.ctypes.data, seq.ctypes.data,
                                      m.ctypes.data, sc.gapopen, sc.gapextn, B, C.byref(idx))
        return best, idx.value, cs, cg

    def extend(self, direction, cores, seq, sc: OracleScoring, p: Params, verbose=0):
        cores = np.ascontiguousarray(cores, dtype=CORE_DTYPE)
        seq = np.ascontiguousarray(seq, dtype=np.uint8)
        n = cores.size
        master = np.zeros(max(p.L, 1), dtype=np.uint8)
        ext_len = np.zeros(n, dtype=np.int32)
        score_add = np.zeros(n, dtype=np.int32)
        secs = C.c_double()
        m = sc.m49()
        bp = self.lib.ramref_extend(direction, n, cores.ctypes.data, seq.ctypes.data, seq.size,
                                    m.ctypes.data, sc.gapopen, sc.gapextn, p.bandwidth, p.cappenalty,
                                    p.minimprovement, p.L, p.when_to_stop, verbose, master.ctypes.data,
                                    ext_len.ctypes.data, score_add.ctypes.data, C.byref(secs))
        rows = int((master != 100).sum())   # the shim pre-fills master[] with 100
        return dict(extended_bp=bp, rows_processed=rows, master=master, ext_len=ext_len,
                    score_add=score_add, seconds=secs.value)

    def cons_seed_extend(self, name, query, q_seed_start, q_seed_end, seq, boundaries, s_seed_start, s_seed_end,
                         orient, bandwidth, adjust=0, y_drop=-1):
        """the reference's cons_seed_extend() object code -> dict(score, query_start, query_end, subj_start,
        subj_end, unadjusted, n_row)"""
        query = np.ascontiguousarray(query, dtype=np.uint8)
        seq = np.ascontiguousarray(seq, dtype=np.uint8)
        b = np.ascontiguousarray(list(boundaries) + [0], dtype=np.uint64)
        out = np.zeros(7, dtype=np.int64)
        nrow = np.zeros(5, dtype=np.int32)
        self.lib.ramref_cons_seed_extend(name.encode(), query.ctypes.data, query.size, q_seed_start, q_seed_end,
                                         seq.ctypes.data, seq.size, b.ctypes.data, s_seed_start, s_seed_end, orient,
                                         bandwidth, adjust, y_drop, out.ctypes.data, nrow.ctypes.data)
        return dict(score=int(out[0]), query_start=int(out[1]), query_end=int(out[2]), subj_start=int(out[3]),
                    subj_end=int(out[4]), unadjusted=int(out[5]), n_row=nrow)

    def global_align(self, name, orient, seq1_start, seq1_end, seq2_start, seq2_end, seq, bandwidth):
        seq = np.ascontiguousarray(seq, dtype=np.uint8)
        return self.lib.ramref_global_align(name.encode(), orient, seq1_start, seq1_end, seq2_start, seq2_end,
                                            seq.ctypes.data, seq.size, bandwidth)

    def load(self, twobit, ranges, max_flanking_bp):
        n, slen = i32(), u64()
        self.lib.ramref_load(twobit.encode(), ranges.encode(), max_flanking_bp, C.byref(n),
                             C.byref(slen), None, None, None, None)
        cores = np.zeros(n.value, dtype=CORE_DTYPE)
        seq = np.zeros(slen.value, dtype=np.uint8)
        offsets = np.zeros(n.value, dtype=np.uint64)
        bounds = np.zeros(n.value, dtype=np.uint64)
        self.lib.ramref_load(twobit.encode(), ranges.encode(), max_flanking_bp, C.byref(n),
                             C.byref(slen), cores.ctypes.data, seq.ctypes.data, offsets.ctypes.data,
                             bounds.ctypes.data)
        return seq, cores, offsets, bounds


class Emulator:
    """tests/emul/libramemul.so -- the real kernel source compiled for the CPU (lock-step SIMT
    emulation). Checks the kernel's logic without a GPU; it is not a product path."""

    def __init__(self):
        src = [os.path.join(EMUL_DIR, "emul_driver.cpp"), os.path.join(EMUL_DIR, "cuda_emul.h"),
               os.path.join(ROOT, "repeatafterme_b200", "csrc", "extend_kernel.cuh"),
               os.path.join(ROOT, "repeatafterme_b200", "csrc", "aux_kernels.cuh"),
               os.path.join(ROOT, "repeatafterme_b200", "csrc", "flatten.h")]
        path = os.path.join(EMUL_DIR, "libramemul.so")
        if not os.path.exists(path) or any(os.path.getmtime(s) > os.path.getmtime(path) for s in src):
            subprocess.check_call(["g++", "-std=c++17", "-O1", "-DRAMCU_EMUL", "-DRAMCU_CHECK", "-I.", "-fPIC", "-shared",
                                   "-pthread", "-o", "libramemul.so", "emul_driver.cpp"], cwd=EMUL_DIR)
        self.lib = lib = C.CDLL(path)
        lib.emul_extend.argtypes = [i32, i32, i32, i32, vp, u64, i32, i32, vp, P(Scoring), P(Params),
                                    vp, vp, vp, P(ExtendResult), i32, vp, vp, vp, vp]
        lib.emul_cfg.argtypes = [i32, P(i32), P(i32), P(i32)]
        lib.emul_extend_pair.argtypes = [i32, i32, i32, i32, vp, u64, i32, vp, P(Scoring), P(Params),
                                         vp, vp, vp, P(ExtendResult), vp, vp, vp, P(ExtendResult)]
        lib.emul_extend_paths.argtypes = [i32, i32, i32, i32, vp, u64, i32, i32, vp, P(Scoring), P(Params),
                                          vp, vp, vp, P(ExtendResult), vp, vp, vp]
        lib.emul_extend_verbose.argtypes = [i32, i32, i32, i32, vp, u64, i32, i32, vp, P(Scoring), P(Params),
                                            vp, vp, vp, P(ExtendResult), i32, vp, vp, vp]

    def configs(self):
        out = []
        for i in range(self.lib.emul_num_cfgs()):
            c, g, w = i32(), i32(), i32()
            self.lib.emul_cfg(i, C.byref(c), C.byref(g), C.byref(w))
            out.append((c.value, g.value, w.value))
        return out

    def configs_for(self, B):
        W = 2 * B + 1
        return [(c, g, w) for (c, g, w) in self.configs() if (w == W) or (w == 0 and c * g >= W)]

    def extend(self, cfg, direction, cores, seq, sc: Scoring, p: Params, threads=128,
               dbg_rows=0, forced=None):
        cores = np.ascontiguousarray(cores, dtype=CORE_DTYPE)
        seq = np.ascontiguousarray(seq, dtype=np.uint8)
        n, W = cores.size, 2 * p.bandwidth + 1
        master = np.zeros(max(p.L, dbg_rows, 1) + 1, dtype=np.uint8)
        ext_len = np.zeros(n, dtype=np.int32)
        score_add = np.zeros(n, dtype=np.int32)
        res = ExtendResult()
        rows = best = sums = None
        rp = bp = sp = fp = None
        if dbg_rows:
            rows = np.zeros((dbg_rows, n, W, 2), dtype=np.int32)
            best = np.zeros((dbg_rows, n, 4), dtype=np.int32)
            sums = np.zeros((dbg_rows, 4), dtype=np.int32)
            rp, bp, sp = rows.ctypes.data, best.ctypes.data, sums.ctypes.data
            if forced is not None:
                forced = np.ascontiguousarray(forced, dtype=np.uint8)
                fp = forced.ctypes.data
        rc = self.lib.emul_extend(cfg[0], cfg[1], cfg[2], threads, seq.ctypes.data, seq.size, direction,
                                  n, cores.ctypes.data, C.byref(sc), C.byref(p), master.ctypes.data,
                                  ext_len.ctypes.data, score_add.ctypes.data, C.byref(res), dbg_rows,
                                  fp, rp, bp, sp)
        assert rc == 0, f"emul_extend rc={rc}"
        return dict(extended_bp=res.extended_bp, rows_processed=res.rows_processed,
                    stop_row=res.stop_row, hit_limit=res.hit_limit, max_score=res.max_score,
                    n_extendable=res.n_extendable, master=master[: res.rows_processed].copy(),
                    ext_len=ext_len, score_add=score_add, rows=rows, best=best, sums=sums)


def emul_extend_pair(emul, cfg, cores, seq, sc, p, threads=128):
    """ramcu_extend_pair through the CPU emulation: both passes in the same threads (NP = 2)"""
    cores = np.ascontiguousarray(cores, dtype=CORE_DTYPE)
    seq = np.ascontiguousarray(seq, dtype=np.uint8)
    n = cores.size
    outs = []
    bufs = []
    for _ in range(2):
        bufs.append((np.zeros(max(p.L, 1) + 1, dtype=np.uint8), np.zeros(n, dtype=np.int32),
                     np.zeros(n, dtype=np.int32), ExtendResult()))
    rc = emul.lib.emul_extend_pair(cfg[0], cfg[1], cfg[2], threads, seq.ctypes.data, seq.size, n, cores.ctypes.data,
                                   C.byref(sc), C.byref(p), bufs[0][0].ctypes.data, bufs[0][1].ctypes.data,
                                   bufs[0][2].ctypes.data, C.byref(bufs[0][3]), bufs[1][0].ctypes.data,
                                   bufs[1][1].ctypes.data, bufs[1][2].ctypes.data, C.byref(bufs[1][3]))
    assert rc == 0, f"emul_extend_pair rc={rc}"
    for master, ext_len, score_add, res in bufs:
        outs.append(dict(extended_bp=res.extended_bp, rows_processed=res.rows_processed, stop_row=res.stop_row,
                         hit_limit=res.hit_limit, max_score=res.max_score, n_extendable=res.n_extendable,
                         master=master[: res.rows_processed].copy(), ext_len=ext_len, score_add=score_add))
    return outs   # [right, left]


def emul_extend_paths(emul, cfg, direction, cores, seq, sc, p, threads=128):
    """ramcu_extend_paths through the CPU emulation of the kernel source"""
    cores = np.ascontiguousarray(cores, dtype=CORE_DTYPE)
    seq = np.ascontiguousarray(seq, dtype=np.uint8)
    n, W, L = cores.size, 2 * p.bandwidth + 1, max(p.L, 1)
    master = np.zeros(L + 1, dtype=np.uint8)
    ext_len = np.zeros(n, dtype=np.int32)
    score_add = np.zeros(n, dtype=np.int32)
    state = np.zeros((L, n, W), dtype=np.uint8)
    best = np.zeros((L, n), dtype=np.int32)
    idx = np.zeros((L, n), dtype=np.int32)
    res = ExtendResult()
    rc = emul.lib.emul_extend_paths(cfg[0], cfg[1], cfg[2], threads, seq.ctypes.data, seq.size, direction, n,
                                    cores.ctypes.data, C.byref(sc), C.byref(p), master.ctypes.data,
                                    ext_len.ctypes.data, score_add.ctypes.data, C.byref(res),
                                    state.ctypes.data, best.ctypes.data, idx.ctypes.data)
    assert rc == 0, f"emul_extend_paths rc={rc}"
    rp = res.rows_processed
    return dict(rows_processed=rp, extended_bp=res.extended_bp, master=master[:rp].copy(), state=state[:rp],
                best=best[:rp], idx=idx[:rp])


def emul_extend_verbose(emul, cfg, direction, cores, seq, sc, p, n_rows, cells=False, threads=128):
    """ramcu_extend_verbose through the CPU emulation of the kernel source"""
    cores = np.ascontiguousarray(cores, dtype=CORE_DTYPE)
    seq = np.ascontiguousarray(seq, dtype=np.uint8)
    n, W, L = cores.size, 2 * p.bandwidth + 1, max(p.L, 1)
    master = np.zeros(L + 1, dtype=np.uint8)
    ext_len = np.zeros(n, dtype=np.int32)
    score_add = np.zeros(n, dtype=np.int32)
    cand = np.zeros((n_rows, n, 4, 4), dtype=np.int32)
    sums = np.zeros((n_rows, 4), dtype=np.int32)
    cel = np.zeros((n_rows, n, 4, W, 2), dtype=np.int32) if cells else None
    res = ExtendResult()
    rc = emul.lib.emul_extend_verbose(cfg[0], cfg[1], cfg[2], threads, seq.ctypes.data, seq.size, direction, n,
                                      cores.ctypes.data, C.byref(sc), C.byref(p), master.ctypes.data,
                                      ext_len.ctypes.data, score_add.ctypes.data, C.byref(res), n_rows,
                                      cand.ctypes.data, sums.ctypes.data, cel.ctypes.data if cells else None)
    assert rc == 0, f"emul_extend_verbose rc={rc}"
    rp = res.rows_processed
    return dict(rows_processed=rp, extended_bp=res.extended_bp, master=master[:rp].copy(), ext_len=ext_len,
                score_add=score_add, cand=cand[:rp], sums=sums[:rp], cells=(cel[:rp] if cells else None))


def parse_verbose_dump(text):
    """The per-candidate values in a RAMExtend -vvvv / -vvvvv stdout (ram_extend.c:992-1090): for each direction
    a dict (row, base, n) -> dict(best, col, oos, gap, sub) and a dict (row, base) -> column sum."""
    import re
    out = {1: ({}, {}), 0: ({}, {})}
    cur = None
    row_re = re.compile(r"^(RIGHT|LEFT) ROW (\d+) with '([ACGT])': n = (\d+)$")
    best_re = re.compile(r"^  best score = (-?\d+) @ column (-?\d+)")
    tot_re = re.compile(r"^  Total Score for '([ACGT])' = (-?\d+)$")
    d = None
    for line in text.splitlines():
        m = row_re.match(line)
        if m:
            d = 1 if m.group(1) == "RIGHT" else 0
            cur = dict(oos=False)
            out[d][0][(int(m.group(2)), "ACGT".index(m.group(3)), int(m.group(4)))] = cur
            last_row = int(m.group(2))
            continue
        if cur is not None:
            if line.startswith("    Gap: "):
                cur["gap"] = [int(x) for x in line[9:].split()]
                continue
            if line.startswith("    Sub: "):
                cur["sub"] = [int(x) for x in line[9:].split()]
                continue
            m = best_re.match(line)
            if m:
                cur["best"], cur["col"] = int(m.group(1)), int(m.group(2))
                continue
            if "**OUT_OF_SEQ**" in line:
                cur["oos"] = True
        m = tot_re.match(line)
        if m and d is not None:
            out[d][1][(last_row, "ACGT".index(m.group(1)))] = int(m.group(2))
    return out


def format_outmat(direction, cores, seq, bandwidth, master, state, best, idx):
    """The -outmat lines of ram_extend.c:1122-1132 from path states (0 substitution, 1 deletion,
    2 insertion): test-side restatement of what csrc/ramx_boundary.c prints."""
    out = []
    B = bandwidth
    for r in range(len(master)):
        for i, c in enumerate(cores):
            if not (c["right_extendable"] if direction else c["left_extendable"]):
                continue
            if direction:
                start = int(c["right_seq_pos"]) + (-1 if c["orient"] else 1)
            else:
                start = int(c["left_seq_pos"]) + (1 if c["orient"] else -1)
            chars = []
            for off in range(-B, B):
                so = (-off - r) if (bool(direction) == bool(c["orient"])) else (off + r)
                p = start + so
                if p < int(c["lower_seq_bound"]) or p > int(c["upper_seq_bound"]) or p < 0:
                    chars.append("X")
                    continue
                v = int(seq[p])
                base = "ACGTacgt"[v] if v < 8 else "N"
                st = int(state[r, i, off + B])
                chars.append(base if st == 0 else "-" if st == 1 else base.lower())
            out.append("dir=%s:n=%d:row=%d: %s best:score=%d:offset=%d:cons=%s" % (
                "right" if direction else "left", i, r, "".join(chars), best[r, i], idx[r, i], "ACGT"[master[r] & 3]))
    return out


def compare_extend(a: dict, b: dict, what=""):
    """bit-exact comparison of two extend results (oracle-style dicts)."""
    for k in ("extended_bp", "rows_processed", "stop_row", "hit_limit", "max_score", "n_extendable"):
        if k in a and k in b:
            assert a[k] == b[k], f"{what}: {k}: {a[k]} != {b[k]}"
    n = min(len(a["master"]), len(b["master"]))
    if "rows_processed" in a:
        n = a["rows_processed"]
    assert np.array_equal(a["master"][:n], b["master"][:n]), f"{what}: consensus differs"
    assert np.array_equal(a["ext_len"], b["ext_len"]), f"{what}: extension lengths differ"
    assert np.array_equal(a["score_add"], b["score_add"]), f"{what}: score increments differ"
