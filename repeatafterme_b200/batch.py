"""Family batches: what util/extend-stk.pl does with one RAMExtend process per family
(extend-stk.pl:242-371) -- here many independent families share one CUDA context, and families
are sharded across the GPUs of a box with no communication between them.

Per family this runs exactly the two calls the reference's main() makes (ram_extend.c:424-427,
506-509): the right pass, then the left pass. The overlap-avoidance step between the passes
(ram_extend.c:445-499) only narrows bounds of cores that share a 2bit sequence; callers that have
such families pass ``between_passes`` (the unchanged CLI keeps running the reference's own loop).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from .abi import RamCu, Scoring, make_params


def family_cost(n: int, bandwidth: int, flank: int) -> float:
    """LPT weight N * (2B+1) * flank."""
    return float(n) * (2 * bandwidth + 1) * max(flank, 1)


def lpt_shard(costs, world_size: int):
    """Longest-processing-time-first partition of families over ranks. Returns a list of index
    lists, one per rank; every family appears exactly once."""
    order = sorted(range(len(costs)), key=lambda i: (-costs[i], i))
    loads = [0.0] * world_size
    shards = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda k: (loads[k], k))
        shards[r].append(i)
        loads[r] += costs[i]
    return shards


@dataclass
class FamilyResult:
    right: dict = field(default_factory=dict)
    left: dict = field(default_factory=dict)

    @property
    def cell_updates(self) -> float:
        cu = 0.0
        for d in (self.right, self.left):
            if d:
                cu += d["rows_processed"] * d["n_extendable"] * (2 * d["bandwidth"] + 1) * 4.0
        return cu


def _as_dict(res, master, ext_len, score_add, bandwidth):
    d = res.as_dict()
    d.update(master=master, ext_len=ext_len, score_add=score_add, bandwidth=bandwidth)
    return d


def extend_families(rc: RamCu, fams, between_passes=None, concurrent_hint=True, timing=None):
    """Extend a list of families (objects with .sequence or .blob, .cores, .params: synth.Family or
    synth.PackedFamily).
    Host buffers in, host results out: library upload, every pass of every family as one batch
    (or right batch, `between_passes` hook, left batch), fetch.
    timing: a dict that receives the host wall clock of the phases (seconds): upload, create, run_fetch."""
    import time
    t_begin = time.perf_counter()
    libs, scs, prs = [], [], []
    total = sum(int(f.cores.size) for f in fams)
    for f in fams:
        p = dict(f.params)
        matrix = p.pop("matrix", "20p43g")
        rs = p.pop("rs", (1, -1, -5))
        sc = rc.scoring(matrix, rs)
        if concurrent_hint and len(fams) > 1:
            p["concurrent_instances"] = total - int(f.cores.size)
        scs.append(sc)
        prs.append(make_params(**p))
        if hasattr(f, "blob"):   # synth.PackedFamily: 2 bits per base, decoded on the device
            libs.append(rc.library_from_twobit(f.blob, f.windows, f.n_start, f.n_len, f.total))
        else:
            libs.append(rc.upload_library(f.sequence))
    results = [FamilyResult() for _ in fams]
    t_upload = time.perf_counter()
    t_create = t_upload
    try:
        # Without a between-passes step (the overlap avoidance of ram_extend.c:445-499 belongs to the
        # caller) the left passes do not depend on the right ones: all 2 x len(fams) column loops go
        # into one batch and share the SMs. With one, right batch -> hook -> left batch.
        if between_passes is None:
            # the two passes of a family are created as a linked pair (ramcu_job_create_pair): where both
            # row states fit the register file, every thread carries a band of both and the passes take turns
            jobs = []
            for i, f in enumerate(fams):
                r, l = rc.job_create_pair(libs[i], f.cores, scs[i], prs[i])
                jobs += [(i, 1, r), (i, 0, l)]
            rounds = [jobs]
            t_create = time.perf_counter()
        else:
            rounds = [(1,), (0,)]
        for rnd in rounds:
            jobs = rnd if isinstance(rnd, list) else [(i, d, rc.job_create(libs[i], d, f.cores, scs[i], prs[i]))
                                                      for d in rnd for i, f in enumerate(fams)]
            if len(jobs) == 1:
                rc.job_run(jobs[0][2])
            else:
                rc.batch_run([j for _, _, j in jobs])
            for i, d, j in jobs:
                out = _as_dict(*rc.job_fetch(j, fams[i].cores.size, prs[i].L), prs[i].bandwidth)
                if d:
                    results[i].right = out
                else:
                    results[i].left = out
                rc.job_free(j)
            if rnd == (1,):
                for i, f in enumerate(fams):
                    between_passes(f, results[i].right)
    finally:
        for h in libs:
            rc.free_library(h)
    if timing is not None:
        t_end = time.perf_counter()
        timing["upload"] = timing.get("upload", 0.0) + (t_upload - t_begin)
        timing["create"] = timing.get("create", 0.0) + (t_create - t_upload)
        timing["run_fetch"] = timing.get("run_fetch", 0.0) + (t_end - t_create)
    return results
