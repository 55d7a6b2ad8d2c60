"""ctypes mirror of include/ramcu.h (the product C-ABI) -- no torch types cross this boundary.

The product path FAILS LOUDLY when ``csrc/libramcu.so`` is missing or no CUDA device is usable:
there is no CPU fallback anywhere in this package.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RAMCU_LIB") or os.path.join(HERE, "csrc", "libramcu.so")   # RAMCU_LIB: kernel experiments

#: struct coreAlignment (reference common.h:80-98) flattened == ramcu_core_t
CORE_DTYPE = np.dtype(
    [
        ("left_seq_pos", "<i8"),
        ("right_seq_pos", "<i8"),
        ("lower_seq_bound", "<i8"),
        ("upper_seq_bound", "<i8"),
        ("orient", "i1"),
        ("left_extendable", "i1"),
        ("right_extendable", "i1"),
        ("pad_", "i1", (5,)),
    ]
)
assert CORE_DTYPE.itemsize == 40

NCODES = 5  # A,C,G,T,N

#: every symbol include/ramcu.h declares
ABI_SYMBOLS = (
    "ramcu_last_error", "ramcu_abi_version", "ramcu_init", "ramcu_destroy", "ramcu_set_stream",
    "ramcu_device_info", "ramcu_int32_peak", "ramcu_scoring_by_name", "ramcu_library_upload",
    "ramcu_library_from_device", "ramcu_library_from_twobit", "ramcu_library_download", "ramcu_library_free", "ramcu_extend", "ramcu_job_create",
    "ramcu_job_run", "ramcu_job_fetch", "ramcu_job_free", "ramcu_batch_run", "ramcu_debug_rows",
    "ramcu_job_create_pair", "ramcu_extend_pair", "ramcu_extend_paths", "ramcu_seed_extend_batch", "ramcu_global_align",
    "ramcu_extend_verbose",
)


class Scoring(C.Structure):
    """ramcu_scoring_t: matrix[cons 0..3][A,C,G,T,N] + gap penalties (score_system.h:7-17)."""

    _fields_ = [("matrix", (C.c_int32 * NCODES) * 4), ("gapopen", C.c_int32), ("gapextn", C.c_int32)]

    def as_array(self) -> np.ndarray:
        return np.array([[self.matrix[a][b] for b in range(NCODES)] for a in range(4)], dtype=np.int32)

    @staticmethod
    def from_array(m, gapopen: int, gapextn: int) -> "Scoring":
        s = Scoring()
        for a in range(4):
            for b in range(NCODES):
                s.matrix[a][b] = int(m[a][b])
        s.gapopen = int(gapopen)
        s.gapextn = int(gapextn)
        return s


class Params(C.Structure):
    """ramcu_params_t: the scalar arguments/globals of extend_alignment() (ram_extend.h:9-13)."""

    _fields_ = [
        ("bandwidth", C.c_int32),
        ("cappenalty", C.c_int32),
        ("minimprovement", C.c_int32),
        ("L", C.c_int32),
        ("when_to_stop", C.c_int32),
        ("lanes_per_instance", C.c_int32),
        ("concurrent_instances", C.c_int32),
        ("cta_thread_limit", C.c_int32),
    ]


class ExtendResult(C.Structure):
    """ramcu_result_t."""

    _fields_ = [
        ("extended_bp", C.c_int32),
        ("rows_processed", C.c_int32),
        ("stop_row", C.c_int32),
        ("hit_limit", C.c_int32),
        ("max_score", C.c_int32),
        ("n_extendable", C.c_int32),
        ("kernel_ms", C.c_float),
        ("lanes_per_instance", C.c_int32),
        ("passes_per_thread", C.c_int32),
    ]

    def as_dict(self) -> dict:
        return {f: getattr(self, f) for f, _ in self._fields_}


def make_params(bandwidth=14, cappenalty=-90, minimprovement=27, L=10000, when_to_stop=100,
                lanes_per_instance=0, concurrent_instances=0, cta_thread_limit=0) -> Params:
    """Defaults are those of the RAMExtend CLI (ram_extend.c:246-249, 301-344)."""
    p = Params()
    p.bandwidth, p.cappenalty, p.minimprovement, p.L, p.when_to_stop, p.lanes_per_instance = (
        bandwidth, cappenalty, minimprovement, L, when_to_stop, lanes_per_instance)
    p.concurrent_instances = concurrent_instances
    p.cta_thread_limit = cta_thread_limit
    return p


class Seed(C.Structure):
    """ramcu_seed_t: one cons_seed_extend() call (reference bnw_extend.h:38-43)."""

    _fields_ = [("query", C.c_void_p), ("query_len", C.c_int32), ("query_seed_start", C.c_int32),
                ("query_seed_end", C.c_int32), ("orient", C.c_int32), ("subj_seed_start", C.c_int64),
                ("subj_seed_end", C.c_int64), ("lower_seq_bound", C.c_int64), ("upper_seq_bound", C.c_int64)]


class SeedResult(C.Structure):
    """ramcu_seed_result_t."""

    _fields_ = [("left_score", C.c_int32), ("right_score", C.c_int32), ("query_start", C.c_int32),
                ("query_end", C.c_int32), ("subj_start", C.c_int64), ("subj_end", C.c_int64),
                ("left_rows", C.c_int32), ("right_rows", C.c_int32)]


class RamCuError(RuntimeError):
    pass


def load_ramcu(path: str = LIB_PATH) -> C.CDLL:
    if not os.path.exists(path):
        raise RamCuError(
            f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). There is no CPU fallback for the product path."
        )
    lib = C.CDLL(path)
    vp, i32, u64 = C.c_void_p, C.c_int32, C.c_uint64
    P = C.POINTER
    lib.ramcu_last_error.restype = C.c_char_p
    lib.ramcu_last_error.argtypes = []
    lib.ramcu_abi_version.restype = i32
    lib.ramcu_init.argtypes = [i32, P(vp)]
    lib.ramcu_destroy.argtypes = [vp]
    lib.ramcu_destroy.restype = None
    lib.ramcu_set_stream.argtypes = [vp, vp]
    lib.ramcu_device_info.argtypes = [vp, P(i32), P(i32), P(i32)]
    lib.ramcu_scoring_by_name.argtypes = [C.c_char_p, i32, i32, i32, P(Scoring)]
    lib.ramcu_int32_peak.argtypes = [vp, P(C.c_double), P(C.c_double)]
    lib.ramcu_library_upload.argtypes = [vp, vp, u64, P(vp)]
    lib.ramcu_library_from_device.argtypes = [vp, vp, u64, P(vp)]
    lib.ramcu_library_from_twobit.argtypes = [vp, vp, u64, i32, vp, i32, vp, vp, u64, P(vp)]
    lib.ramcu_library_download.argtypes = [vp, vp, vp, vp]
    lib.ramcu_library_free.argtypes = [vp, vp]
    lib.ramcu_extend.argtypes = [vp, vp, i32, i32, vp, P(Scoring), P(Params), vp, vp, vp, P(ExtendResult)]
    lib.ramcu_job_create.argtypes = [vp, vp, i32, i32, vp, P(Scoring), P(Params), P(vp)]
    lib.ramcu_job_run.argtypes = [vp, vp]
    lib.ramcu_job_fetch.argtypes = [vp, vp, vp, vp, vp, P(ExtendResult)]
    lib.ramcu_job_free.argtypes = [vp, vp]
    lib.ramcu_batch_run.argtypes = [vp, i32, P(vp)]
    lib.ramcu_job_create_pair.argtypes = [vp, vp, i32, vp, P(Scoring), P(Params), P(vp), P(vp)]
    lib.ramcu_extend_pair.argtypes = [vp, vp, i32, vp, P(Scoring), P(Params), vp, vp, vp, P(ExtendResult),
                                      vp, vp, vp, P(ExtendResult)]
    lib.ramcu_extend_paths.argtypes = [vp, vp, i32, i32, vp, P(Scoring), P(Params), vp, vp, vp, P(ExtendResult),
                                       i32, vp, vp, vp]
    lib.ramcu_extend_verbose.argtypes = [vp, vp, i32, i32, vp, P(Scoring), P(Params), vp, vp, vp, P(ExtendResult),
                                         i32, vp, vp, vp]
    lib.ramcu_seed_extend_batch.argtypes = [vp, vp, i32, vp, P(Scoring), vp, i32, i32, vp]
    lib.ramcu_global_align.argtypes = [vp, vp, i32, vp, i32, C.c_int64, C.c_int64, P(Scoring), vp, i32, P(i32), P(i32), P(i32)]
    lib.ramcu_debug_rows.argtypes = [vp, vp, i32, i32, vp, P(Scoring), P(Params), i32, vp, vp, vp, vp]
    for name in ABI_SYMBOLS:
        if name not in ("ramcu_last_error", "ramcu_destroy"):
            getattr(lib, name).restype = i32
    return lib


def scoring_by_name(lib: C.CDLL, name: str, rs=(1, -1, -5)) -> Scoring:
    s = Scoring()
    rc = lib.ramcu_scoring_by_name(name.encode(), rs[0], rs[1], rs[2], C.byref(s))
    if rc:
        raise RamCuError(lib.ramcu_last_error().decode())
    return s


class RamCu:
    """Host-side mirror of the reference's extension interface, over the CUDA C-ABI.

    ``extend(direction, ...)`` has the argument meaning of ``extend_alignment()``
    (ram_extend.h:9-13): direction 1 = right, 0 = left; it returns the extended bp and the same
    per-core side effects (extension length, score increment) as arrays.
    """

    def __init__(self, device: int = 0, path: str = LIB_PATH):
        self.lib = load_ramcu(path)
        self.ctx = C.c_void_p()
        self._check(self.lib.ramcu_init(device, C.byref(self.ctx)))

    def _check(self, rc: int):
        if rc:
            raise RamCuError(f"ramcu error {rc}: {self.lib.ramcu_last_error().decode()}")

    def close(self):
        if self.ctx:
            self.lib.ramcu_destroy(self.ctx)
            self.ctx = C.c_void_p()

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    def device_info(self) -> dict:
        a, b, c = C.c_int32(), C.c_int32(), C.c_int32()
        self._check(self.lib.ramcu_device_info(self.ctx, C.byref(a), C.byref(b), C.byref(c)))
        return {"sm_count": a.value, "cc": (b.value, c.value)}

    def int32_peak_gops(self) -> float:
        g, ms = C.c_double(), C.c_double()
        self._check(self.lib.ramcu_int32_peak(self.ctx, C.byref(g), C.byref(ms)))
        return g.value

    def set_stream(self, cuda_stream: int | None):
        self._check(self.lib.ramcu_set_stream(self.ctx, C.c_void_p(cuda_stream or 0)))

    def scoring(self, name: str, rs=(1, -1, -5)) -> Scoring:
        return scoring_by_name(self.lib, name, rs)

    # -- library ---------------------------------------------------------------------------
    def upload_library(self, seq: np.ndarray) -> C.c_void_p:
        seq = np.ascontiguousarray(seq, dtype=np.uint8)
        h = C.c_void_p()
        self._check(self.lib.ramcu_library_upload(self.ctx, seq.ctypes.data, seq.size, C.byref(h)))
        return h

    def library_from_device(self, dev_ptr: int, length: int) -> C.c_void_p:
        h = C.c_void_p()
        self._check(self.lib.ramcu_library_from_device(self.ctx, C.c_void_p(dev_ptr), length, C.byref(h)))
        return h

    def library_from_twobit(self, blob: np.ndarray, windows: np.ndarray, n_start, n_len, total: int) -> C.c_void_p:
        """windows: uint64 array [n][4] = (dst, len, src_byte, src_base); see ramcu_library_from_twobit"""
        blob = np.ascontiguousarray(blob, dtype=np.uint8)
        windows = np.ascontiguousarray(windows, dtype=np.uint64).reshape(-1, 4)
        n_start = np.ascontiguousarray(n_start, dtype=np.uint64)
        n_len = np.ascontiguousarray(n_len, dtype=np.uint64)
        h = C.c_void_p()
        self._check(self.lib.ramcu_library_from_twobit(self.ctx, blob.ctypes.data, blob.size, windows.shape[0],
                                                       windows.ctypes.data, n_start.size, n_start.ctypes.data,
                                                       n_len.ctypes.data, total, C.byref(h)))
        return h

    def download_library(self, h, length: int):
        n = (length + 31) // 32
        lib2 = np.zeros(2 * n, dtype=np.uint32)
        libn = np.zeros(n, dtype=np.uint32)
        self._check(self.lib.ramcu_library_download(self.ctx, h, lib2.ctypes.data, libn.ctypes.data))
        return lib2, libn

    def free_library(self, h):
        self._check(self.lib.ramcu_library_free(self.ctx, h))

    # -- one extend_alignment() call ----------------------------------------------------------
    def extend(self, lib_h, direction: int, cores: np.ndarray, scoring: Scoring, params: Params):
        cores = np.ascontiguousarray(cores, dtype=CORE_DTYPE)
        n = cores.size
        master = np.zeros(max(params.L, 1), dtype=np.uint8)
        ext_len = np.zeros(n, dtype=np.int32)
        score_add = np.zeros(n, dtype=np.int32)
        res = ExtendResult()
        self._check(self.lib.ramcu_extend(self.ctx, lib_h, direction, n, cores.ctypes.data,
                                          C.byref(scoring), C.byref(params), master.ctypes.data,
                                          ext_len.ctypes.data, score_add.ctypes.data, C.byref(res)))
        return res, master[: res.rows_processed].copy(), ext_len, score_add

    def extend_paths(self, lib_h, direction: int, cores: np.ndarray, scoring: Scoring, params: Params):
        """ramcu_extend_paths: extend() plus the -outmat data (path state per band cell, best score and
        first best band index per row and core)."""
        cores = np.ascontiguousarray(cores, dtype=CORE_DTYPE)
        n, W, L = cores.size, 2 * params.bandwidth + 1, max(params.L, 1)
        master = np.zeros(L, dtype=np.uint8)
        ext_len = np.zeros(n, dtype=np.int32)
        score_add = np.zeros(n, dtype=np.int32)
        state = np.zeros((L, n, W), dtype=np.uint8)
        best = np.zeros((L, n), dtype=np.int32)
        idx = np.zeros((L, n), dtype=np.int32)
        res = ExtendResult()
        self._check(self.lib.ramcu_extend_paths(self.ctx, lib_h, direction, n, cores.ctypes.data, C.byref(scoring),
                                                C.byref(params), master.ctypes.data, ext_len.ctypes.data,
                                                score_add.ctypes.data, C.byref(res), L, state.ctypes.data,
                                                best.ctypes.data, idx.ctypes.data))
        rp = res.rows_processed
        return res, master[:rp].copy(), ext_len, score_add, state[:rp], best[:rp], idx[:rp]

    def extend_verbose(self, lib_h, direction: int, cores: np.ndarray, scoring: Scoring, params: Params,
                       n_rows: int, cells: bool = False):
        """ramcu_extend_verbose: extend() plus, for the first n_rows rows, what compute_nw_row() returned and left
        behind for every candidate base (the data of the reference's VERBOSE >= 10 dumps): cand[r, core, a] =
        (best, best column index, gap of band cell 0, gap of band cell 2B), sums[r, a], and with cells=True the
        (sub, gap) pair of every band cell of every candidate's row."""
        cores = np.ascontiguousarray(cores, dtype=CORE_DTYPE)
        n, W, L = cores.size, 2 * params.bandwidth + 1, max(params.L, 1)
        master = np.zeros(L, dtype=np.uint8)
        ext_len = np.zeros(n, dtype=np.int32)
        score_add = np.zeros(n, dtype=np.int32)
        cand = np.zeros((max(n_rows, 1), n, 4, 4), dtype=np.int32)
        sums = np.zeros((max(n_rows, 1), 4), dtype=np.int32)
        cel = np.zeros((max(n_rows, 1), n, 4, W, 2), dtype=np.int32) if cells else None
        res = ExtendResult()
        self._check(self.lib.ramcu_extend_verbose(self.ctx, lib_h, direction, n, cores.ctypes.data, C.byref(scoring),
                                                  C.byref(params), master.ctypes.data, ext_len.ctypes.data,
                                                  score_add.ctypes.data, C.byref(res), n_rows, cand.ctypes.data,
                                                  sums.ctypes.data, cel.ctypes.data if cells else None))
        rp = min(res.rows_processed, n_rows)
        return res, master[:res.rows_processed].copy(), ext_len, score_add, cand[:rp], sums[:rp], (cel[:rp] if cells else None)

    # -- split form -------------------------------------------------------------------------
    def job_create(self, lib_h, direction: int, cores: np.ndarray, scoring: Scoring, params: Params):
        cores = np.ascontiguousarray(cores, dtype=CORE_DTYPE)
        h = C.c_void_p()
        self._check(self.lib.ramcu_job_create(self.ctx, lib_h, direction, cores.size, cores.ctypes.data,
                                              C.byref(scoring), C.byref(params), C.byref(h)))
        return h

    def job_create_pair(self, lib_h, cores: np.ndarray, scoring: Scoring, params: Params):
        """Right and left pass of one family laid out to share one launch (ramcu_job_create_pair);
        run them with batch_run([right, left])."""
        cores = np.ascontiguousarray(cores, dtype=CORE_DTYPE)
        r, l = C.c_void_p(), C.c_void_p()
        self._check(self.lib.ramcu_job_create_pair(self.ctx, lib_h, cores.size, cores.ctypes.data,
                                                   C.byref(scoring), C.byref(params), C.byref(r), C.byref(l)))
        return r, l

    def extend_pair(self, lib_h, cores: np.ndarray, scoring: Scoring, params: Params):
        """Both extend_alignment() calls of main() in one launch; valid as the left pass only when the
        overlap-avoidance step between them changes no bound (include/ramcu.h)."""
        jobs = self.job_create_pair(lib_h, cores, scoring, params)
        try:
            self.batch_run(list(jobs))
            return [self.job_fetch(j, cores.size, params.L) for j in jobs]
        finally:
            for j in jobs:
                self.job_free(j)

    def job_run(self, job):
        self._check(self.lib.ramcu_job_run(self.ctx, job))

    def batch_run(self, jobs):
        arr = (C.c_void_p * len(jobs))(*[j.value for j in jobs])
        self._check(self.lib.ramcu_batch_run(self.ctx, len(jobs), arr))

    def job_fetch(self, job, n_cores: int, L: int):
        master = np.zeros(max(L, 1), dtype=np.uint8)
        ext_len = np.zeros(n_cores, dtype=np.int32)
        score_add = np.zeros(n_cores, dtype=np.int32)
        res = ExtendResult()
        self._check(self.lib.ramcu_job_fetch(self.ctx, job, master.ctypes.data, ext_len.ctypes.data,
                                             score_add.ctypes.data, C.byref(res)))
        return res, master[: res.rows_processed].copy(), ext_len, score_add

    def job_free(self, job):
        self._check(self.lib.ramcu_job_free(self.ctx, job))

    # -- seed extension (cons_seed_extend / pair_seed_extend / global_align, bnw_extend.c:167-649) ------
    def seed_extend_batch(self, lib_h, seeds, scoring: Scoring, n_row, bandwidth: int, y_drop: int = -1):
        """seeds: list of dicts(query=uint8 codes, query_seed_start, query_seed_end, subj_seed_start,
        subj_seed_end, lower_seq_bound, upper_seq_bound, orient). Returns a list of SeedResult; the
        reference's return value before complexity adjustment is left_score + right_score."""
        n = len(seeds)
        arr = (Seed * max(n, 1))()
        keep = []
        for i, sd in enumerate(seeds):
            q = np.ascontiguousarray(sd["query"], dtype=np.uint8)
            keep.append(q)
            arr[i].query = q.ctypes.data
            arr[i].query_len = q.size
            for k in ("query_seed_start", "query_seed_end", "orient", "subj_seed_start", "subj_seed_end",
                      "lower_seq_bound", "upper_seq_bound"):
                setattr(arr[i], k, int(sd[k]))
        res = (SeedResult * max(n, 1))()
        nr = np.ascontiguousarray(n_row, dtype=np.int32)
        self._check(self.lib.ramcu_seed_extend_batch(self.ctx, lib_h, n, C.byref(arr), C.byref(scoring), nr.ctypes.data,
                                                     bandwidth, y_drop, C.byref(res)))
        return [res[i] for i in range(n)]

    def global_align(self, lib_h, orient: int, seq1, seq2_start: int, seq2_end: int, scoring: Scoring, n_row,
                     bandwidth: int):
        q = np.ascontiguousarray(seq1, dtype=np.uint8)
        nr = np.ascontiguousarray(n_row, dtype=np.int32)
        g, l, p = C.c_int32(), C.c_int32(), C.c_int32()
        self._check(self.lib.ramcu_global_align(self.ctx, lib_h, orient, q.ctypes.data, q.size, seq2_start, seq2_end,
                                                C.byref(scoring), nr.ctypes.data, bandwidth, C.byref(g), C.byref(l),
                                                C.byref(p)))
        return g.value, l.value, p.value

    # -- test-only: compute_nw_row boundary -----------------------------------------------------
    def debug_rows(self, lib_h, direction: int, cores: np.ndarray, scoring: Scoring, params: Params,
                   n_rows: int, forced=None):
        cores = np.ascontiguousarray(cores, dtype=CORE_DTYPE)
        n, W = cores.size, 2 * params.bandwidth + 1
        rows = np.zeros((n_rows, n, W, 2), dtype=np.int32)
        best = np.zeros((n_rows, n, 4), dtype=np.int32)
        sums = np.zeros((n_rows, 4), dtype=np.int32)
        fptr = None
        if forced is not None:
            forced = np.ascontiguousarray(forced, dtype=np.uint8)
            assert forced.size >= n_rows
            fptr = forced.ctypes.data
        self._check(self.lib.ramcu_debug_rows(self.ctx, lib_h, direction, n, cores.ctypes.data,
                                              C.byref(scoring), C.byref(params), n_rows, fptr,
                                              rows.ctypes.data, best.ctypes.data, sums.ctypes.data))
        return rows, best, sums
