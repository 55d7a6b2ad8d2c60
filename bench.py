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
def _cpu_worker(args):
    seed, n, flank, kind = args
    from repeatafterme_b200 import synth
    from repeatafterme_b200.abi import make_params
    from tests import support
    fam = synth.make_family(seed, n, flank, WORKLOAD["divergence"])
    p = make_params(bandwidth=WORKLOAD["bandwidth"], minimprovement=WORKLOAD["minimprovement"],
                    cappenalty=WORKLOAD["cappenalty"], L=WORKLOAD["L"], when_to_stop=WORKLOAD["when_to_stop"])
    ora = support.Oracle()
    sc = ora.scoring(WORKLOAD["matrix"])
    ref = support.Reference.try_load() if kind == "reference" else None
    secs, cu = 0.0, 0.0
    for d in (1, 0):
        if ref is not None:
            o = ref.extend(d, fam.cores, fam.sequence, sc, p)
            secs += o["seconds"]
            rows = o["rows_processed"]
        else:
            t0 = time.perf_counter()
            o = ora.extend(d, fam.cores, fam.sequence, sc, p)
            secs += time.perf_counter() - t0
            rows = o["rows_processed"]
        next_ = int(((fam.cores["right_extendable"] if d else fam.cores["left_extendable"]) != 0).sum())
        cu += rows * next_ * (2 * WORKLOAD["bandwidth"] + 1) * 4.0
    return secs, cu


def cpu_baseline(n_inst: int, flank: int, procs: int, seed0: int = 900):
    """one sample family per process on `procs` host cores at once (the reference is single
    threaded; util/extend-stk.pl runs one RAMExtend process per family). Returns
    (aggregate GCUPS, single-process GCUPS, kind)."""
    import multiprocessing as mp
    kind = "reference" if os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libramref.so")) else "port"
    ctx = mp.get_context("spawn")
    t0 = time.perf_counter()
    with ctx.Pool(procs) as pool:
        outs = pool.map(_cpu_worker, [(seed0 + i, n_inst, flank, kind) for i in range(procs)])
    wall = time.perf_counter() - t0
    cu = sum(o[1] for o in outs)
    slowest = max(o[0] for o in outs)
    single = outs[0][1] / outs[0][0] / 1e9
    return cu / slowest / 1e9, single, kind, wall


# ---------------------------------------------------------------------------------------------
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    n_inst, flank = 32, WORKLOAD["flank"]
    vals = []
    for step in range(args.warmup + args.steps):
        agg, single, kind, wall = cpu_baseline(n_inst, flank, cores, seed0=900 + 1000 * step)
        if step >= args.warmup:
            vals.append((agg, single, wall))
    agg = float(np.mean([v[0] for v in vals]))
    single = float(np.mean([v[1] for v in vals]))
    sample = (f"{cores} processes, each the two extend_alignment() calls of one synthetic family of "
              f"{n_inst} instances x {flank} bp flanks (config-4 shape, DP time only via clock_gettime)")
    line = {"impl": "reference", "metric": "banded_dp_gcups", "value": round(agg, 4), "unit": "GCUPS",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(1e3 * float(np.mean([v[2] for v in vals])), 1), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": workload_config(args.gpus),
            "cpu_baseline": {"value": round(agg, 4), "unit": "GCUPS", "cores": cores, "kind": kind,
                             "sample": sample, "single_process_value": round(single, 4)},
            "e2e": {"value": round(agg, 4), "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def workload_config(n_gpus: int) -> dict:
    return {"workload": "synthetic single family per GPU: 10k instances, ~15% divergence, 5 kb flanks, "
                        "bandwidth 40, 14p43g, minimprovement 30, L 20000 (BASELINE configs[3])",
            "instances": WORKLOAD["n"], "flank_bp": WORKLOAD["flank"], "bandwidth": WORKLOAD["bandwidth"],
            "matrix": WORKLOAD["matrix"], "families_per_step": n_gpus, "parallelism": f"families x{n_gpus}",
            "l2": "L2 flushed between timed steps (256 MiB write)"}


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
    # (ram_extend.c:445-499) cannot change a bound and the left pass reads the same cores. Checked
    # below against two separate ramcu_extend calls. --sequential times the two-launch form.
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
    bp = (outs[0][0].extended_bp, outs[1][0].extended_bp)
    for d, o in zip((1, 0), outs):   # the pair launch must equal the reference's two-call form
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
    feed_alloc = dev_seq.numel() * (0.25 + 0.125)             # the two output buffers are zero-filled first
    del dev_seq

    # ---- config-5 style batch: many independent families sharing the GPU (families/s) -------------
    fam_info = None
    fam_stats = None
    if args.families > 0:
        # BASELINE configs[4] shape: a seeded sample of the 2 000-family mix, `--families` per GPU, sharded by
        # family over the ranks (longest-processing-time-first, no communication between ranks)
        specs = synth.config5_specs(5, 2000)
        pick = sorted(np.random.default_rng(0).choice(len(specs), size=args.families * world, replace=False))
        costs = [batch.family_cost(specs[i]["n"], specs[i]["params"]["bandwidth"], specs[i]["flank"]) for i in pick]
        mine = batch.lpt_shard(costs, world)[rank]
        fams5 = [synth.family_from_spec(specs[pick[i]]) for i in mine]
        batch.extend_families(rc, fams5)          # warm-up: first use of every tiling in the mix loads its code
        times5 = []
        for _ in range(3):
            barrier()
            t0 = time.perf_counter()
            res5 = batch.extend_families(rc, fams5)
            torch.cuda.synchronize()
            times5.append(time.perf_counter() - t0)
        cu5 = sum(x.cell_updates for x in res5)
        fam_stats = (float(np.median(times5)), float(len(fams5)), float(cu5), float(sum(f.n for f in fams5)))

    for j in jobs:
        rc.job_free(j)
    rc.free_library(lib)

    # ---- max over ranks -------------------------------------------------------------------------
    if fam_stats is not None:
        fs = torch.tensor(fam_stats, dtype=torch.float64, device="cuda")
        if world > 1:
            fmx = fs.clone()
            dist.all_reduce(fmx, op=dist.ReduceOp.MAX)
            dist.all_reduce(fs, op=dist.ReduceOp.SUM)
            fs[0] = fmx[0]
        dt5, nf5, cu5, ni5 = (float(x) for x in fs)
        fam_info = {"n_families": int(nf5), "instances": int(ni5), "families_per_s": round(nf5 / dt5, 2),
                    "gcups_e2e": round(cu5 / dt5 / 1e9, 1), "seconds": round(dt5, 3),
                    "what": "seeded sample of the BASELINE configs[4] family mix (N 100-10k log-uniform, "
                            "bandwidth 20-100, flank 0.5-5 kb), %d families per GPU sharded by family (LPT) over "
                            "the ranks, host buffers in / host results out; per rank the median of three runs "
                            "after one warm-up run, max over ranks" % args.families}
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
        tf = "r1_extend_kernel_traffic.json" if args.sequential else "r1_extend_kernel_pair_traffic.json"
        traffic = json.load(open(os.path.join(ROOT, "profiles", tf)))["dram_bytes_per_launch"]
    except Exception:
        pass

    cpu = None
    if world == 1 or rank == 0:
        cores = os.cpu_count() or 1
        agg, single, kind, wall = cpu_baseline(args.cpu_instances, WORKLOAD["flank"], cores)
        cpu = {"value": round(agg, 4), "unit": "GCUPS", "cores": cores, "kind": kind,
               "single_process_value": round(single, 4),
               "sample": f"{cores} processes at once, each one synthetic family of {args.cpu_instances} instances x "
                         f"{WORKLOAD['flank']} bp flanks (config-4 shape), reference extend_alignment() object "
                         f"code timed with clock_gettime around the two calls ({wall:.1f}s wall)"}

    line = {"metric": "banded_dp_gcups", "value": round(gcups, 1), "unit": "GCUPS", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(dev_ms / args.steps, 3),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32",
            "data": "synthetic", "config": dict(workload_config(world), instances=n, flank_bp=flank,
                                                lanes_per_instance=lanes, extended_bp=list(bp),
                                                passes="two launches, right then left" if args.sequential else
                                                "right and left pass in one launch (independent here: every "
                                                "instance is its own sequence; checked against separate calls)"),
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
                                     "zero-fill of its two output buffers", "bases": int(fam.sequence.size),
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
    ap.add_argument("--families", type=int, default=12, help="size of the config-5 family-batch sample (0 = skip)")
    ap.add_argument("--cpu-instances", type=int, default=48, help="instances of the CPU-baseline sample family")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_cuda(args)


if __name__ == "__main__":
    main()
