#!/usr/bin/env python
"""bench.py -- banded-DP extension throughput (GCUPS) of the CUDA path, with the reference's own
CPU implementation timed beside it.

  python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path
  python bench.py --impl reference [--steps K] [--warmup W]      # reference RAMExtend on host cores

Workload (BASELINE.json configs[3], SURVEY.md 8d "config 4"): one synthetic family of 10 000
instances at ~15 % divergence with 5 kb flanks, bandwidth 40, matrix 14p43g, minimprovement 30,
L 20000. One STEP = one full extension of the family = the two extend_alignment() calls the
reference's main() makes (right pass, then left pass), ~3.3e10 cell updates. With N GPUs every
rank extends its own family of that shape (families are independent: no collective on the data
path, weak scaling); `value` = cell updates of all ranks / max-over-ranks device time.

CU (cell update) = one band cell x one instance x one row x one candidate base:
    CU = sum_dir rows_processed * n_extendable * (2*bandwidth+1) * 4          
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = dict(n=10000, flank=5000, divergence=0.15, bandwidth=40, matrix="14p43g", minimprovement=30,
                cappenalty=-90, L=20000, when_to_stop=100)
OPS_PER_CU = 8            # algorithmic int32 add/max per cell update
NOMINAL_INT32_GOPS = 148 * 64 * 1.965   # 148 SM x 64 INT32 lanes x 1.965 GHz


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# ---------------------------------------------------------------------------------------------
# clocks during the timed region
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.samples, self.stop_flag, self.thread = index, [], threading.Event(), None

    def _run(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True,
                                     timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.1)

    def start(self):
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()

    def stop(self) -> dict:
        self.stop_flag.set()
        if self.thread:
            self.thread.join(timeout=6)
        sm, mx, reasons = [], 0, set()
        for s in self.samples:
            try:
                sm.append(float(s[0]))
                mx = max(mx, float(s[1]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------
# CPU baseline: the reference's own extend_alignment() object code (oracle/_ref/libramref.so)
# ---------------------------------------------------------------------------------------------
CPU_SAMPLE = dict(n=1000, rows=300)   # instances / columns per direction of the CPU sample (config-4 shape)
# the reference's extend_alignment() on the WHOLE config-4 family (N=10000, 2 x 5100 columns), one process, in the
# build container while tests/golden/full_config4.npz was produced (tests/golden/make_golden_full.py)
FULL_CONFIG4_CPU = {"seconds": 879.0, "gcups": 0.0376, "where": "build container (Xeon, 8 vCPU), one core per direction"}


def _cpu_worker(args):
    seed, n, flank, kind, rows, steps = args
    from repeatafterme_b200 import synth
    from repeatafterme_b200.abi import make_params
    from tests import support
    fam = synth.make_family(seed, n, flank, WORKLOAD["divergence"])
    p = make_params(bandwidth=WORKLOAD["bandwidth"], minimprovement=WORKLOAD["minimprovement"],
                    cappenalty=WORKLOAD["cappenalty"], L=min(WORKLOAD["L"], rows), when_to_stop=WORKLOAD["when_to_stop"])
    ora = support.Oracle()
    sc = ora.scoring(WORKLOAD["matrix"])
    ref = support.Reference.try_load() if kind == "reference" else None
    out = []
    for _ in range(steps):
        secs, cu = 0.0, 0.0
        for d in (1, 0):
            if ref is not None:
                o = ref.extend(d, fam.cores, fam.sequence, sc, p)
                secs += o["seconds"]
            else:
                t0 = time.perf_counter()
                o = ora.extend(d, fam.cores, fam.sequence, sc, p)
                secs += time.perf_counter() - t0
            next_ = int(((fam.cores["right_extendable"] if d else fam.cores["left_extendable"]) != 0).sum())
            cu += o["rows_processed"] * next_ * (2 * WORKLOAD["bandwidth"] + 1) * 4.0
        out.append((secs, cu))
    return out


def cpu_baseline(procs: int, steps: int = 1, seed0: int = 900, n_inst: int = CPU_SAMPLE["n"],
                 rows: int = CPU_SAMPLE["rows"]):
    """The reference's own extend_alignment() object code (oracle/_ref/libramref.so; the oracle port if
    the reference could not be built) on `procs` host cores at once, one process per core as
    util/extend-stk.pl runs one RAMExtend per family: every process extends its own synthetic family
    of the config-4 shape (n_inst instances x 5 kb flanks, not cache resident) for the first `rows`
    columns of both directions, `steps` times. Returns per step (aggregate GCUPS = cell updates of all
    processes / slowest process, single-process GCUPS), the kind and the wall seconds."""
    import multiprocessing as mp
    kind = "reference" if os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libramref.so")) else "port"
    t0 = time.perf_counter()
    with mp.get_context("spawn").Pool(procs) as pool:
        outs = pool.map(_cpu_worker, [(seed0 + i, n_inst, WORKLOAD["flank"], kind, rows, steps) for i in range(procs)])
    wall = time.perf_counter() - t0
    per_step = []
    for k in range(steps):
        cu = sum(o[k][1] for o in outs)
        slowest = max(o[k][0] for o in outs)
        per_step.append((cu / slowest / 1e9, outs[0][k][1] / outs[0][k][0] / 1e9, slowest))
    return per_step, kind, wall


def cpu_sample_text(cores: int) -> str:
    return (f"{cores} processes at once (one per host core, as extend-stk.pl runs one RAMExtend per family), each the "
            f"two extend_alignment() calls of one synthetic family of {CPU_SAMPLE['n']} instances x {WORKLOAD['flank']} bp "
            f"flanks (config-4 shape) limited to the first {CPU_SAMPLE['rows']} columns per direction; DP time only "
            f"(clock_gettime around the calls); aggregate = cell updates of all processes / slowest process")


# ---------------------------------------------------------------------------------------------
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    per_step, kind, wall = cpu_baseline(cores, steps=args.warmup + args.steps)
    vals = per_step[args.warmup:]
    agg = float(np.mean([v[0] for v in vals]))
    single = float(np.mean([v[1] for v in vals]))
    line = {"impl": "reference", "metric": "banded_dp_gcups", "value": round(agg, 4), "unit": "GCUPS",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(1e3 * float(np.mean([v[2] for v in vals])), 1), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": workload_config(args.gpus),
            "cpu_baseline": {"value": round(agg, 4), "unit": "GCUPS", "cores": cores, "kind": kind,
                             "sample": cpu_sample_text(cores), "single_process_value": round(single, 4),
                             "full_config4_single_process": FULL_CONFIG4_CPU},
            "e2e": {"value": round(agg, 4), "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def workload_config(n_gpus: int) -> dict:
    return {"workload": "synthetic single family per GPU: 10k instances, ~15% divergence, 5 kb flanks, "
                        "bandwidth 40, 14p43g, minimprovement 30, L 20000 (BASELINE configs[3])",
            "instances": WORKLOAD["n"], "flank_bp": WORKLOAD["flank"], "bandwidth": WORKLOAD["bandwidth"],
            "matrix": WORKLOAD["matrix"], "minimprovement": WORKLOAD["minimprovement"], "L": WORKLOAD["L"],
            "families_per_step": n_gpus, "parallelism": f"families x{n_gpus}",
            "l2": "GPU arm: L2 flushed between timed steps (256 MiB write); CPU arm: working set larger than the caches"}


# ---------------------------------------------------------------------------------------------
# the product CLI on the same family: what a RAMExtend user sees
# ---------------------------------------------------------------------------------------------
def cli_leg(fam, runs=2):
    """bin/RAMExtend (the reference's unmodified main() linked with this repo's extend_alignment() and flank
    loader) on the config-4 family written as .2bit + ranges files: wall seconds of the whole process (CUDA context
    creation, .2bit loading, both passes, overlap avoidance, report, -cons / -outtsv / -outfa), with the breakdown
    the boundary prints under RAMX_TIMING=1. The first run warms the page cache and is not reported."""
    import tempfile
    from repeatafterme_b200 import synth
    exe = os.path.join(ROOT, "repeatafterme_b200", "bin", "RAMExtend")
    if not os.path.exists(exe):
        return {"unavailable": "bin/RAMExtend not built (needs the reference sources at build time: host/_refobj)"}
    with tempfile.TemporaryDirectory() as tmp:
        tb, rg = os.path.join(tmp, "f.2bit"), os.path.join(tmp, "f.tsv")
        synth.write_family_files(fam, tb, rg)
        cmd = [exe, "-twobit", tb, "-ranges", rg, "-outtsv", os.path.join(tmp, "o.tsv"), "-outfa", os.path.join(tmp, "o.fa"),
               "-cons", os.path.join(tmp, "c.fa"), "-bandwidth", str(WORKLOAD["bandwidth"]), "-matrix", WORKLOAD["matrix"],
               "-minimprovement", str(WORKLOAD["minimprovement"]), "-L", str(WORKLOAD["L"])]
        walls, r = [], None
        for _ in range(runs + 1):
            t0 = time.perf_counter()
            r = subprocess.run(cmd, capture_output=True, text=True, env=dict(os.environ, RAMX_TIMING="1"))
            walls.append(time.perf_counter() - t0)
            if r.returncode != 0:
                return {"unavailable": "bin/RAMExtend exited %d: %s" % (r.returncode, (r.stdout + r.stderr)[-300:])}
        ext = [l.strip() for l in r.stdout.splitlines() if l.startswith("Extended")]
        return {"wall_s": round(min(walls[1:]), 3), "runs": [round(w, 3) for w in walls[1:]],
                "breakdown": [l[len("[ramx timing] "):] for l in r.stderr.splitlines() if l.startswith("[ramx timing]")],
                "extended": ext, "outputs": "-cons -outtsv -outfa written",
                "what": "wall clock of the whole bin/RAMExtend process on BASELINE config 4 as .2bit + ranges files "
                        "(context creation + loader + both passes + overlap avoidance + report + three output files)"}


# ---------------------------------------------------------------------------------------------
# BASELINE configs[4]: synthetic genome-scale round
# ---------------------------------------------------------------------------------------------
def config5_round(args, rc, stream, world, rank, local, barrier, peak_gops):
    """2 000 seeded families (N 100-10k log-uniform, bandwidth 20-100, flank 0.5-5 kb, 10 % one-sided
    cores; synth.config5_specs) -- the SAME set at every GPU count, sharded by family over the ranks
    (longest-processing-time-first on N x band x flank, no communication between ranks: strong
    scaling). `--families K` (K > 0) takes the first K families of a seeded permutation instead.
    Per rank: the shard's families are generated as 2-bit packed libraries (4 bases per byte, as a
    .2bit stores them), then
      device leg : libraries decoded into HBM and jobs created beforehand; CUDA events around ONE
                   ramcu_batch_run of all 2 x F column loops (right+left pass of a family in the same
                   threads where both row states fit); median of the timed runs
      e2e leg    : host blobs in, host results out: ramcu_library_from_twobit + ramcu_job_create_pair
                   + ramcu_batch_run + ramcu_job_fetch + frees, wall clock
    and the families of the full-size parity sample that fall into the shard are compared with the
    reference's outputs (tests/golden/full_config5_sample.npz)."""
    import torch
    from repeatafterme_b200 import batch, synth
    specs = synth.config5_specs(5, 2000)
    ids = list(range(len(specs)))
    if args.families > 0:
        ids = sorted(np.random.default_rng(0).permutation(len(specs))[: args.families].tolist())
    costs = [batch.family_cost(specs[i]["n"], specs[i]["params"]["bandwidth"], specs[i]["flank"]) for i in ids]
    mine = [ids[k] for k in batch.lpt_shard(costs, world)[rank]]
    local_world = int(os.environ.get("LOCAL_WORLD_SIZE", str(world)))
    procs = max(1, (os.cpu_count() or 1) // max(local_world, 1))
    t0 = time.perf_counter()
    fams = synth.generate_packed_families([specs[i] for i in mine], procs=procs)
    gen_s = time.perf_counter() - t0
    log(f"[rank {rank}] config 5: {len(fams)} of {len(ids)} families, {sum(f.total for f in fams) / 1e9:.2f} Gbases "
        f"generated in {gen_s:.0f} s on {procs} processes")
    # the end-to-end leg uploads from PINNED host memory: one pinned arena holds the packed libraries of the shard
    t0 = time.perf_counter()
    arena = torch.empty(sum((f.blob.size + 255) & ~255 for f in fams) + 256, dtype=torch.uint8, pin_memory=True)
    off = 0
    for f in fams:
        view = arena[off: off + f.blob.size].numpy()
        view[:] = f.blob
        f.blob = view
        off += (view.size + 255) & ~255
    log(f"[rank {rank}] config 5: {off / 1e9:.2f} GB of packed libraries in pinned host memory ({time.perf_counter() - t0:.1f} s)")

    # ---- device-resident leg --------------------------------------------------------------------------
    total_inst = sum(f.n for f in fams)
    libs, jobs, prs = [], [], []
    for f in fams:
        pr = dict(f.params)
        sc = rc.scoring(pr.pop("matrix"))
        pr["concurrent_instances"] = total_inst - f.n
        from repeatafterme_b200.abi import make_params
        prm = make_params(**pr)
        h = rc.library_from_twobit(f.blob, f.windows, f.n_start, f.n_len, f.total)
        libs.append(h)
        prs.append(prm)
        jobs += list(rc.job_create_pair(h, f.cores, sc, prm))
    rc.batch_run(jobs)               # warm-up: first use of every tiling loads its code
    torch.cuda.synchronize()
    dev = []
    for _ in range(max(1, args.family_runs)):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        rc.batch_run(jobs)
        e1.record(stream)
        torch.cuda.synchronize()
        dev.append(e0.elapsed_time(e1) * 1e-3)
    cu, outs = 0.0, {}
    for k, f in enumerate(fams):
        for d, j in zip((1, 0), jobs[2 * k: 2 * k + 2]):
            o = rc.job_fetch(j, f.n, prs[k].L)
            cu += o[0].rows_processed * o[0].n_extendable * (2 * prs[k].bandwidth + 1) * 4.0
            outs[(mine[k], d)] = o
    for j in jobs:
        rc.job_free(j)
    for h in libs:
        rc.free_library(h)

    # parity at full size against the reference, for the sample families of this shard
    checked = 0
    golden = os.path.join(ROOT, "tests", "golden", "full_config5_sample.npz")
    if os.path.exists(golden):
        z = np.load(golden, allow_pickle=False)
        meta = json.loads(str(z["meta"]))
        for key, m in meta.items():
            if int(key) not in mine:
                continue
            for d in (1, 0):
                o, rows = outs[(int(key), d)], m[f"rows{d}"]
                if not (o[0].extended_bp == m[f"bp{d}"] and o[0].rows_processed == rows and
                        np.array_equal(o[1][:rows], z[f"{key}_master{d}"][:rows]) and
                        np.array_equal(o[2], z[f"{key}_ext_len{d}"]) and np.array_equal(o[3], z[f"{key}_score_add{d}"])):
                    raise SystemExit(f"bench.py: config-5 family {key} direction {d} differs from the reference golden")
            checked += 1
    del outs

    # ---- end-to-end leg -----------------------------------------------------------------------------------
    e2e = []
    for _ in range(max(1, args.family_runs)):
        barrier()
        t0 = time.perf_counter()
        phases = {}
        batch.extend_families(rc, fams, timing=phases)
        torch.cuda.synchronize()
        e2e.append(time.perf_counter() - t0)
        log(f"[rank {rank}] config 5 e2e {e2e[-1]:.3f} s: " + ", ".join(f"{k} {v:.3f}" for k, v in phases.items()))
    h2d = int(sum(f.blob.nbytes + 2 * f.cores.nbytes for f in fams))
    stats = (float(np.median(dev)), float(np.median(e2e)), float(len(fams)), float(cu), float(total_inst), gen_s)
    chk = torch.tensor([float(checked)], dtype=torch.float64, device="cuda")
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(chk, op=dist.ReduceOp.SUM)
    extra = {"what": "BASELINE configs[4]: %d seeded families (N 100-10k log-uniform, bandwidth 20-100, flank 0.5-5 kb, 10%% "
                     "one-sided cores), the same set at every GPU count, sharded by family (LPT) over the ranks with no "
                     "communication; device = CUDA events around one ramcu_batch_run of all column loops with libraries "
                     "and jobs resident, e2e = 2-bit packed libraries in pinned host memory in / host results out (upload + decode, job "
                     "creation, run, fetch); per rank the median of %d runs after a warm-up, max over ranks"
                     % (len(ids), max(1, args.family_runs)),
             "families_checked_against_reference_golden": int(chk.item()),
             "h2d_bytes_rank0": h2d}
    return stats, extra


# ---------------------------------------------------------------------------------------------
def run_cuda(args):
    import torch
    import torch.distributed as dist
    from repeatafterme_b200 import RamCu, batch, synth
    from repeatafterme_b200.abi import make_params

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback "
                         "(use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    # stdout carries exactly one JSON line: libraries that write to fd 1 (NCCL prints its version
    # banner there) are pointed at stderr, and the line goes to the saved descriptor at the end
    sys.stdout.flush()
    json_out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/null")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rc = RamCu(local)
    stream = torch.cuda.Stream()          # a real (non-default) stream, shared with the C-ABI
    torch.cuda.set_stream(stream)
    rc.set_stream(stream.cuda_stream)

    n, flank = args.instances, args.flank
    t0 = time.time()
    fam = synth.make_family(4 + 17 * rank, n, flank, WORKLOAD["divergence"])
    log(f"[rank {rank}] generated family N={n} flank={flank} ({fam.sequence.size / 1e6:.1f} Mbases) in {time.time() - t0:.1f}s")
    sc = rc.scoring(WORKLOAD["matrix"])
    p = make_params(bandwidth=WORKLOAD["bandwidth"], minimprovement=WORKLOAD["minimprovement"],
                    cappenalty=WORKLOAD["cappenalty"], L=WORKLOAD["L"], when_to_stop=WORKLOAD["when_to_stop"],
                    lanes_per_instance=args.lanes)

    # ---- device-resident leg: library + descriptors in HBM before the timed region -------------
    host_seq = torch.from_numpy(fam.sequence).pin_memory()
    lib = rc.upload_library(host_seq.numpy())
    # Right and left pass share one launch (ramcu_job_create_pair): every instance of the synthetic
    # family is its own sequence, so the overlap-avoidance step main() runs between the passes
    # (ram_extend.c:445-499) cannot change a bound and the left pass reads the same cores. The planner
    # decides between two single-pass layouts side by side and both passes in the same threads. Checked
    # below against the reference golden. --sequential times the two-launch form.
    if args.sequential:
        jobs = [rc.job_create(lib, d, fam.cores, sc, p) for d in (1, 0)]
    else:
        jobs = list(rc.job_create_pair(lib, fam.cores, sc, p))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def step_device():
        if args.sequential:
            for j in jobs:
                rc.job_run(j)
        else:
            rc.batch_run(jobs)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step_device()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    evs = []
    for _ in range(args.steps):
        flush.zero_()                       # evict the library / state from L2 between steps
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        step_device()
        e1.record(stream)
        evs.append((e0, e1))
    barrier()
    dev_ms = sum(a.elapsed_time(b) for a, b in evs)
    clocks = sampler.stop()
    outs = [rc.job_fetch(j, fam.cores.size, p.L) for j in jobs]
    cu_step = sum(o[0].rows_processed * o[0].n_extendable * (2 * p.bandwidth + 1) * 4.0 for o in outs)
    lanes = outs[0][0].lanes_per_instance
    passes_per_thread = outs[0][0].passes_per_thread
    bp = (outs[0][0].extended_bp, outs[1][0].extended_bp)
    # the timed job's outputs against the reference: rank 0's family is the one whose reference outputs are
    # committed (tests/golden/full_config4.npz: the reference's extend_alignment() object code on exactly
    # this input, 2 x 440 s of one core); other ranks / sizes compare with two separate ramcu_extend calls
    checked = "two separate ramcu_extend calls"
    golden = os.path.join(ROOT, "tests", "golden", "full_config4.npz")
    if rank == 0 and n == WORKLOAD["n"] and flank == WORKLOAD["flank"] and os.path.exists(golden):
        import hashlib
        z = np.load(golden, allow_pickle=False)
        meta = json.loads(str(z["meta"]))["seed4"]
        if hashlib.sha256(fam.sequence.tobytes()).hexdigest() != meta["seq_sha"]:
            raise SystemExit("bench.py: the generated family is not the one the reference golden was made from")
        for d, o in zip((1, 0), outs):
            rows = meta[f"rows{d}"]
            if not (o[0].extended_bp == meta[f"bp{d}"] and o[0].rows_processed == rows and
                    np.array_equal(o[1][:rows], z[f"seed4_master{d}"][:rows]) and
                    np.array_equal(o[2], z[f"seed4_ext_len{d}"]) and np.array_equal(o[3], z[f"seed4_score_add{d}"])):
                raise SystemExit(f"bench.py: outputs of the timed job differ from the reference golden (direction {d})")
        checked = "outputs == reference golden (tests/golden/full_config4.npz)"
    else:
        for d, o in zip((1, 0), outs):
            alone = rc.extend(lib, d, fam.cores, sc, p)
            same = (alone[0].rows_processed == o[0].rows_processed and alone[0].extended_bp == o[0].extended_bp and
                    np.array_equal(alone[1], o[1]) and np.array_equal(alone[2], o[2]) and np.array_equal(alone[3], o[3]))
            if not same:
                raise SystemExit(f"bench.py: pair launch differs from a separate ramcu_extend call (direction {d})")

    # ---- end-to-end leg: host buffers in, host results out, every step ---------------------------
    def step_e2e(verbose=False):
        t = [time.perf_counter()]
        h = rc.upload_library(host_seq.numpy())
        t.append(time.perf_counter())
        if args.sequential:
            res = [rc.extend(h, d, fam.cores, sc, p) for d in (1, 0)]
        else:
            res = rc.extend_pair(h, fam.cores, sc, p)
        t.append(time.perf_counter())
        rc.free_library(h)
        t.append(time.perf_counter())
        if verbose:
            log("[e2e breakdown ms] upload %.1f both passes %.1f free %.1f" %
                tuple(1e3 * (b - a) for a, b in zip(t[:-1], t[1:])))
        return res

    step_e2e(verbose=True)
    step_e2e(verbose=True)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        r = step_e2e()
    barrier()
    e2e_s = time.perf_counter() - t0
    h2d = int(fam.sequence.nbytes + 2 * fam.cores.nbytes)
    d2h = int(sum(x[1].nbytes + x[2].nbytes + x[3].nbytes + 32 for x in r))

    # same leg, but the flanks are uploaded as the .2bit file stores them (2 bits per base) and
    # decoded on the device (ramcu_library_from_twobit) instead of 1 byte per base
    blob, wins, ns, nl, total = synth.family_twobit_windows(fam)
    blob_pinned = torch.from_numpy(blob).pin_memory()

    def step_e2e_twobit():
        h = rc.library_from_twobit(blob_pinned.numpy(), wins, ns, nl, total)
        if args.sequential:
            res = [rc.extend(h, d, fam.cores, sc, p) for d in (1, 0)]
        else:
            res = rc.extend_pair(h, fam.cores, sc, p)
        rc.free_library(h)
        return res

    step_e2e_twobit()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e_twobit()
    barrier()
    e2e2_s = time.perf_counter() - t0
    h2d2 = int(blob.nbytes + wins.nbytes + ns.nbytes + nl.nbytes + 2 * fam.cores.nbytes)

    peak_gops = rc.int32_peak_gops()

    # ---- flank feed in isolation: byte codes already in HBM -> packed library (HBM-bound kernel) ----
    dev_seq = host_seq.to(device="cuda", non_blocking=False)
    feed_ms = []
    for i in range(5):
        flush.zero_()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record(stream)
        hfeed = rc.library_from_device(dev_seq.data_ptr(), dev_seq.numel())
        f1.record(stream)
        torch.cuda.synchronize()
        if i > 0:
            feed_ms.append(f0.elapsed_time(f1))
        rc.free_library(hfeed)
    feed_bytes = dev_seq.numel() * (1.0 + 0.25 + 0.125)        # read 1 B/base, write 2 bit + 1 bit
    feed_alloc = 0.0
    del dev_seq

    n_bases = int(fam.sequence.size)
    cli = None
    if rank == 0 and world == 1 and not args.no_cli and n == WORKLOAD["n"] and flank == WORKLOAD["flank"]:
        torch.cuda.synchronize()
        try:
            cli = cli_leg(fam)
        except Exception as e:          # the leg is informative: never let it take the bench line down
            cli = {"unavailable": f"{type(e).__name__}: {e}"}
        log(f"[cli] {cli}")
    # the single-family leg is done: release it before the genome-scale round
    for j in jobs:
        rc.job_free(j)
    rc.free_library(lib)
    jobs, lib = [], None
    del host_seq, blob_pinned, fam

    # ---- BASELINE configs[4]: the genome-scale round, sharded by family over the ranks -------------------
    fam_info = None
    fam_stats = None
    if args.families != 0:
        fam_stats, fam_extra = config5_round(args, rc, stream, world, rank, local, barrier, peak_gops)

    # ---- max over ranks -------------------------------------------------------------------------
    if fam_stats is not None:
        # [device s, e2e s, families, cell updates, instances, generation s]
        fs = torch.tensor(fam_stats, dtype=torch.float64, device="cuda")
        fmx, fsum = fs.clone(), fs.clone()
        if world > 1:
            dist.all_reduce(fmx, op=dist.ReduceOp.MAX)
            dist.all_reduce(fsum, op=dist.ReduceOp.SUM)
        dev5, e2e5, gen5 = float(fmx[0]), float(fmx[1]), float(fmx[5])
        nf5, cu5, ni5 = float(fsum[2]), float(fsum[3]), float(fsum[4])
        fam_info = dict(fam_extra, n_families=int(nf5), instances=int(ni5), cell_updates=cu5,
                        seconds_device=round(dev5, 3), seconds_e2e=round(e2e5, 3),
                        families_per_s=round(nf5 / e2e5, 2), families_per_s_device=round(nf5 / dev5, 2),
                        gcups=round(cu5 / dev5 / 1e9, 1), gcups_e2e=round(cu5 / e2e5 / 1e9, 1),
                        # LPT balance: busy time of the ranks (device seconds), max over mean
                        rank_busy_max_over_mean=round(dev5 / (float(fsum[0]) / world), 4),
                        roofline_frac=round(cu5 / dev5 / 1e9 * OPS_PER_CU / world / peak_gops, 4),
                        generation_seconds=round(gen5, 1), scaling="strong")
    stats = torch.tensor([dev_ms, e2e_s, cu_step, e2e2_s], dtype=torch.float64, device="cuda")
    if world > 1:
        mx = stats.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = stats.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        dev_ms, e2e_s, cu_total, e2e2_s = float(mx[0]), float(mx[1]), float(sm[2]), float(mx[3])
    else:
        cu_total = cu_step
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    gcups = cu_total * args.steps / (dev_ms * 1e-3) / 1e9
    e2e_gcups = cu_total * args.steps / e2e_s / 1e9
    per_gpu_gops = cu_step * args.steps / (dev_ms * 1e-3) / 1e9 * OPS_PER_CU
    flank_bytes = sum(o[0].rows_processed * o[0].n_extendable for o in outs) / 4.0   # 2 bits/base/row
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    traffic = None
    try:   # DRAM bytes per launch of the dominant kernel, from the committed ncu --set full capture
        tf = "r2_extend_kernel_traffic.json" if args.sequential else "r2_final_extend_kernel_duo_traffic.json"
        traffic = json.load(open(os.path.join(ROOT, "profiles", tf)))["dram_bytes_per_launch"]
    except Exception:
        pass

    cpu = None
    if world == 1 or rank == 0:
        cores = os.cpu_count() or 1
        per_step, kind, wall = cpu_baseline(cores, steps=2)
        agg, single = per_step[-1][0], per_step[-1][1]
        cpu = {"value": round(agg, 4), "unit": "GCUPS", "cores": cores, "kind": kind,
               "single_process_value": round(single, 4), "sample": cpu_sample_text(cores) + f" ({wall:.0f} s wall)",
               "full_config4_single_process": FULL_CONFIG4_CPU}

    line = {"metric": "banded_dp_gcups", "value": round(gcups, 1), "unit": "GCUPS", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(dev_ms / args.steps, 3),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32",
            "data": "synthetic", "config": dict(workload_config(world), instances=n, flank_bp=flank),
            "details": {"lanes_per_instance": lanes, "extended_bp": list(bp), "parity": checked,
                        "passes": "two launches, right then left" if args.sequential else
                                  ("right and left pass of the family in ONE launch (independent here: every instance is "
                                   "its own sequence): " +
                                   ("both passes in the same threads, taking turns row by row" if passes_per_thread == 2 else
                                    "each pass on its own CTAs with the layout it would have alone, the two sets of CTAs "
                                    "co-resident, so one pass computes while the other's column sums travel")),
                        "passes_per_thread": passes_per_thread,
                        "timing": "CUDA events on the launching stream around each step, L2 flushed between steps "
                                  "(256 MiB write), max over ranks"},
            "families_per_s": round(world * args.steps / (dev_ms * 1e-3), 2),
            "roofline": {"bound": "int32_alu", "achieved": round(per_gpu_gops, 1), "peak": round(peak_gops, 1),
                         "unit": "Gop/s", "frac": round(per_gpu_gops / peak_gops, 4),
                         "peak_source": "measured in this run (ramcu_int32_peak: independent fused add+max chains "
                                        "on all SMs, one op per instruction-lane)",
                         "nominal_peak": round(NOMINAL_INT32_GOPS, 1),
                         "frac_of_nominal": round(per_gpu_gops / NOMINAL_INT32_GOPS, 4),
                         "ops_per_cell_update": OPS_PER_CU, "traffic": traffic,
                         "traffic_unit": "DRAM bytes per extend_kernel launch (ncu, config 4)",
                         "hbm": {"achieved": round(flank_bytes * args.steps / (dev_ms * 1e-3) / 1e9, 3),
                                 "peak": hbm_peak, "unit": "GB/s",
                                 "note": "algorithmic flank bytes (2 bits per base per instance-row); the path is "
                                         "ALU bound, HBM is a sanity figure"}},
            "flank_feed": {"kernel": "pack_library_kernel (1 byte/base codes in HBM -> 2 bit/base + N mask), with the "
                                     "initialisation of the slack around its two output buffers", "bases": n_bases,
                           "ms": round(float(np.median(feed_ms)), 4),
                           "achieved": round((feed_bytes + feed_alloc) / (float(np.median(feed_ms)) * 1e-3) / 1e9, 1),
                           "peak": hbm_peak, "unit": "GB/s",
                           "frac": round((feed_bytes + feed_alloc) / (float(np.median(feed_ms)) * 1e-3) / 1e9 / hbm_peak, 3),
                           "note": "CUDA events around ramcu_library_from_device, L2 flushed before each call; "
                                   "the 2bit-native decode kernels are timed in profiles/ (ncu)"},
            "cpu_baseline": cpu,
            "e2e": {"value": round(e2e_gcups, 1), "unit": "GCUPS", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": round(1e3 * e2e_s / args.steps, 2),
                    "what": "ramcu_library_upload (pinned host bytes -> HBM, packed on device) + both passes "
                            "(ramcu_extend_pair, or two ramcu_extend calls with --sequential) with host outputs, per step"},
            "e2e_twobit": {"value": round(cu_total * args.steps / e2e2_s / 1e9, 1), "unit": "GCUPS",
                           "h2d_bytes_per_step": h2d2, "d2h_bytes_per_step": d2h,
                           "ms_per_step": round(1e3 * e2e2_s / args.steps, 2),
                           "what": "same, but flanks uploaded as 2bit packed DNA (ramcu_library_from_twobit, decoded "
                                   "on the device) instead of the reference loader's 1 byte per base"},
            "gpu_launches": (2 if args.sequential else 1) * args.steps,
            "clocks": clocks}
    if cli:
        line["cli"] = cli
    if fam_info:
        line["family_batch"] = fam_info
    print(json.dumps(line), file=json_out, flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--instances", type=int, default=WORKLOAD["n"])
    ap.add_argument("--flank", type=int, default=WORKLOAD["flank"])
    ap.add_argument("--lanes", type=int, default=0, help="force lanes per instance (0 = auto)")
    ap.add_argument("--sequential", action="store_true",
                    help="time the right and the left pass as two launches (the form the unchanged CLI uses when "
                         "overlap avoidance may apply) instead of the pair launch")
    ap.add_argument("--families", type=int, default=-1,
                    help="genome-scale round (BASELINE configs[4]): -1 = all 2000 families, K > 0 = a seeded subset of K, "
                         "0 = skip")
    ap.add_argument("--no-cli", action="store_true", help="skip the product CLI wall-clock leg (N=1 only)")
    ap.add_argument("--family-runs", type=int, default=2, help="timed runs of the genome-scale round per leg")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_cuda(args)


if __name__ == "__main__":
    main()
